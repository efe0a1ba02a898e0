"""Turns ncu exports into the markdown summaries kept under profiles/.
  python tools/ncu_report.py launches <launches.csv> "<title / command>"      -> launch list table
  python tools/ncu_report.py full <raw.csv> "<title / command>"               -> per-kernel metric tables
(raw.csv = `ncu -i X.ncu-rep --page raw --csv`)"""
import csv, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(fn, title):
    rows = [r for r in csv.reader(open(fn)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, ig, ib, iu = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Grid Size', 'Block Size', 'Metric Unit'))
    L = []
    for r in rows[1:]:
        v = float(r[iv].replace(',', ''))
        if r[iu] == 'ns':
            v /= 1000.0
        L.append((r[ik], r[ig], r[ib], v))
    tot = sum(x[3] for x in L)
    print(f"# {title}\n")
    print("Per-launch times are cold-cache and serialised under the profiler (compare shares, not absolutes).\n")
    print("| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|")
    for i, (k, g, b, v) in enumerate(L):
        print(f"| {i} | `{k[:88]}` | {g} | {b} | {v:.2f} | {100 * v / tot:.1f}% |")
    own = sum(1 for x in L if 'ag::' in x[0])
    print(f"\nTotal {tot:.1f} us over {len(L)} launches ({own} own kernels, the rest CUB radix sort / select).")


def full(fn, title):
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        grid = vals[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''
        block = vals[hdr.index('Block Size')] if 'Block Size' in hdr else ''
        print(f"### `{name[:100]}`  grid {grid} block {block}\n")
        print("| metric | value | unit |\n|---|---|---|")
        d = dict(zip(hdr, zip(vals, units)))
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k][0]} | {d[k][1]} |")
        st = sorted(((float(v[0] or 0), h) for h, v in d.items() if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio')), reverse=True)[:5]
        print("\nTop warp stall reasons (per issue): " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in st) + "\n")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2], sys.argv[3])
