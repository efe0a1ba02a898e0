"""Writes the round-2 summaries under profiles/ from what tools/r02_profiles.sh and tools/r02_scaling.sh left in gpurun_out/.
  python tools/r02_collect.py"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def ncu_rows(fn):
    lines = [l for l in open(fn) if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    return rows[0], rows[1:]


def short(name):
    m = re.search(r"(k_\w+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name[:70]


def json_line(fn):
    for l in open(fn):
        if l.startswith("{"):
            return json.loads(l)
    return None


# ---- launch list of the bench command
fn = os.path.join(O, "r02_launches_bench.csv")
if os.path.exists(fn):
    hdr, rows = ncu_rows(fn)
    ik, iv, ig, ib = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) <= iv:
            continue
        a = agg.setdefault(short(r[ik]), dict(n=0, us=0.0, grid=r[ig], block=r[ib], full=r[ik]))
        a["n"] += 1
        a["us"] += float(r[iv].replace(",", "")) / 1000.0
        a["grid"] = r[ig]
    tot = sum(a["us"] for a in agg.values())
    with open(os.path.join(P, "r02_launches_bench.md"), "w") as f:
        f.write("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-extras` (round 2, final)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600` (first 600 launches of the command: untimed "
                "priming, warm-up and timed steps of both passes; the kernel nodes of the replayed CUDA graphs are profiled one by "
                "one). Per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes. Raw "
                "list: `r02_launches_bench.csv`.\n\n| kernel | launches | grid (last) | block | total us | mean us | share |\n"
                "|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"| `{k}` | {a['n']} | {a['grid']} | {a['block']} | {a['us']:.1f} | {a['us'] / a['n']:.2f} | {100 * a['us'] / tot:.1f}% |\n")
        other = [a["full"][:80] for k, a in agg.items() if not k.startswith("k_")]
        f.write(f"\nTotal {tot:.0f} us over {sum(a['n'] for a in agg.values())} launches. Kernels that are not this library's: {other} "
                "(the L2 flush of bench.py).\n")
    open(os.path.join(P, "r02_launches_bench.csv"), "w").write("".join(l for l in open(fn) if not l.startswith("==")))

# ---- one step, every kernel
fn = os.path.join(O, "r02_step_kernels.csv")
if os.path.exists(fn):
    hdr, rows = ncu_rows(fn)
    iid, ik, im, iv, ig, ib = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Grid Size", "Block Size"))
    L = collections.OrderedDict()
    for r in rows:
        if len(r) <= iv:
            continue
        d = L.setdefault(r[iid], dict(name=short(r[ik]), grid=r[ig], block=r[ib]))
        d[r[im]] = float(r[iv].replace(",", ""))
    tot = sum(d.get("gpu__time_duration.sum", 0) for d in L.values()) / 1000.0
    with open(os.path.join(P, "r02_step_kernels.md"), "w") as f:
        f.write("# One step of config 2, production normal mode, every kernel (round 2, final)\n\n"
                "`AG_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,"
                "dram__bytes_write.sum --clock-control none --profile-from-start off python tools/profile_step.py 2 3 0` (one "
                "launch per kernel in pipeline order; serialised and cold-cache under the profiler — `k_rank_picks` and "
                "`k_compact_slots` overlap their neighbours on the side stream in the real step).\n\n"
                "| # | kernel | grid | block | us | share | warp instructions | DRAM read | DRAM written |\n|---|---|---|---|---|---|---|---|---|\n")
        for i, d in enumerate(L.values()):
            us = d.get("gpu__time_duration.sum", 0) / 1000.0
            f.write(f"| {i} | `{d['name']}` | {d['grid']} | {d['block']} | {us:.2f} | {100 * us / tot:.1f}% | "
                    f"{int(d.get('smsp__inst_executed.sum', 0)):,} | {d.get('dram__bytes_read.sum', 0) / 1e6:.2f} MB | "
                    f"{d.get('dram__bytes_write.sum', 0) / 1e6:.2f} MB |\n")
        f.write(f"\nTotal {tot:.1f} us over {len(L)} launches, all of them this library's kernels.\n")

# ---- full capture: metric tables + hot source lines, DRAM traffic of the roofline kernel
fn = os.path.join(O, "r02_full_raw.csv")
if os.path.exists(fn):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_report.py"), "full", fn,
                          "ncu --set full, config 2 production mode, one step (AG_NO_GRAPH=1), round 2 (final)"],
                         capture_output=True, text=True).stdout
    for k, obj in (("k_hand_sweep", "sweep.o"), ("k_hog_svm", "hog_svm.o"), ("k_taubin_solve", "quadric.o"),
                   ("k_rank_picks", "quadric.o"), ("k_ball_moments", "quadric.o"), ("k_emit_bitmap", "preprocess.o")):
        src = os.path.join(O, f"r02_full_src_{k}.csv")
        if os.path.exists(src) and os.path.getsize(src) > 1000:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src,
                                os.path.join(ROOT, "agile_grasp_b200", "csrc", obj), k, "14"], capture_output=True, text=True)
            out += f"## {k} — source lines by stall samples\n```\n{r.stdout}```\n"
    open(os.path.join(P, "r02_ncu_full_config2.md"), "w").write(out)
    rows = list(csv.reader(open(fn)))
    hdr = rows[0]
    for vals in rows[2:]:
        if "k_ball_moments" in vals[hdr.index("Kernel Name")]:
            rd, wr = (vals[hdr.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            ur, uw = (rows[1][hdr.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            total = float(rd.replace(",", "")) * mul.get(ur, 1) + float(wr.replace(",", "")) * mul.get(uw, 1)
            json.dump({"kernel": "k_ball_moments", "dram_bytes_per_launch": total,
                       "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch of the bench workload "
                                 "(config 2, 2000 samples; the 1.3 MB voxel cloud and the neighbour lists stay in L2)"},
                      open(os.path.join(P, "taubin_traffic.json"), "w"), indent=1)
            break

# ---- bench lines, parity report, scaling
for src, dst in (("r02_final_bench.json", "r02_bench_default.json"), ("r02_final_ref.json", "r02_bench_reference_arm.json")):
    if os.path.exists(os.path.join(O, src)):
        d = json_line(os.path.join(O, src))
        if d:
            open(os.path.join(P, dst), "w").write(json.dumps(d) + "\n")
if os.path.exists(os.path.join(O, "parity_report.json")):
    open(os.path.join(P, "r02_parity_report.json"), "w").write(open(os.path.join(O, "parity_report.json")).read())
scal = {}
d1 = json_line(os.path.join(O, "r02_final_bench.json")) if os.path.exists(os.path.join(O, "r02_final_bench.json")) else None
if d1:
    scal["1"] = d1
for n in (2, 4, 8):
    fn = os.path.join(O, f"r02_scale_{n}.json")
    if os.path.exists(fn):
        d = json_line(fn)
        if d:
            scal[str(n)] = d
if scal:
    keep = {n: {k: d[k] for k in ("n_gpus", "value", "ms_per_step", "e2e", "stages_ms", "strong_scaling", "gathered_last_step",
                                   "clocks", "config") if k in d} for n, d in scal.items()}
    t1 = scal.get("1", {}).get("ms_per_step")
    summ = {}
    for n, d in scal.items():
        s = d.get("strong_scaling", {})
        summ[n] = {"weak_ms_per_step": d["ms_per_step"], "weak_time_efficiency": (t1 / d["ms_per_step"]) if t1 else None,
                   "weak_hyp_per_s": d["value"], "e2e_ms_per_cloud": d["e2e"]["ms_per_cloud"],
                   "config5_ms_per_cloud": s.get("config5", {}).get("ms_per_cloud"),
                   "config4_ms_per_cloud": s.get("config4", {}).get("ms_per_cloud")}
    c51 = summ.get("1", {}).get("config5_ms_per_cloud")
    for n in summ:
        if c51 and summ[n]["config5_ms_per_cloud"]:
            summ[n]["config5_speedup"] = c51 / summ[n]["config5_ms_per_cloud"]
    json.dump({"summary": summ, "runs": keep}, open(os.path.join(P, "r02_scaling.json"), "w"), indent=1)
    print(json.dumps(summ, indent=1))
