"""Per-source-line totals of an ncu SASS source page:  ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv
   python tools/ncu_lines.py src.csv <object.o> <kernel substring> [top]
Correlates the SASS addresses with the -lineinfo line table of the object file (nvdisasm -g) and prints the source
lines that execute the most warp instructions / collect the most stall samples."""
import csv, re, subprocess, sys, tempfile, os
src_csv, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
# the export holds one table per profiled launch ("Kernel Name" row, header row, SASS rows): take the first whose name matches
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
tab = next((a, b) for a, b in zip(starts, starts[1:]) if kname in rows[a][1])
hdr = rows[tab[0] + 1]; data = rows[tab[0] + 2:tab[1]]
ia, ie, ism = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout
# walk the listing: track current function, current line marker, instruction offsets
line_of = {}
cur_fn, cur_line, infn = None, None, False
for ln in txt.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        cur_fn = m.group(1); infn = kname in cur_fn; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = 'inlined' in m.group(3)
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m and infn:
        line_of[int(m.group(1), 16)] = cur_line
base = min(int(r[ia], 16) for r in data)
agg = {}
tot_e = tot_s = 0
for r in data:
    off = int(r[ia], 16) - base
    e, s = int(r[ie]), int(r[ism])
    k = line_of.get(off, ('?', 0))
    a = agg.setdefault(k, [0, 0]); a[0] += e; a[1] += s
    tot_e += e; tot_s += s
print('total warp inst', tot_e, 'samples', tot_s)
srcs = {}
for (f, l), (e, s) in sorted(agg.items(), key=lambda x: (-x[1][1] if os.environ.get("BY","s")=="s" else -x[1][0]))[:top]:
    if f not in srcs:
        p = [os.path.join(os.path.dirname(os.path.abspath(obj)), f), os.path.join(os.path.dirname(os.path.abspath(obj)), '..', '..', 'include', f)]
        srcs[f] = next((open(q).read().splitlines() for q in p if os.path.exists(q)), [])
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ''
    print(f"{f}:{l:5d} inst%={100*e/tot_e:5.1f} samp%={100*s/tot_s:5.1f}  {text}")
