"""SASS evidence table for profiles/: counts of the mnemonics that show TMA bulk copies (UBLKCP), mbarrier traffic (SYNCS),
binary64 arithmetic and the absence of tensor-core / tensor-map instructions, per kernel of every object file.
  python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "agile_grasp_b200", "csrc")
print("# SASS evidence (round 2)\n")
print("`cuobjdump -sass` of the objects linked into `agile_grasp_b200/libag_b200.so` (sm_100a). TMA bulk copies appear as "
      "`UBLKCP`, their mbarrier as `SYNCS`; no `UTMALDG` (no tensor-map copies: every run is a 1-D contiguous range), no "
      "`UTC*MMA` / `HMMA` (nothing on this path is a contraction at a precision an MMA reproduces — the POLY SVM product "
      "must keep OpenCV's rounding order).\n")
print("| object | kernel | UBLKCP | SYNCS | DFMA/DADD/DMUL | ATOMS/REDS | MATCH | UTMALDG | UTCMMA/HMMA |\n|---|---|---|---|---|---|---|---|---|")
for obj in ("quadric.o", "sweep.o", "preprocess.o", "hog_svm.o", "api.o", "handles.o"):
    path = os.path.join(root, obj)
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, counts, order = None, {}, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = {}
            order.append(cur)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1).split(".")[0]
            counts[cur][op] = counts[cur].get(op, 0) + 1
    for fn in order:
        c = counts[fn]
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        m2 = re.search(r"(k_\w+(?:<[^>]*>)?)", name)
        name = m2.group(1) if m2 else name[:60]
        g = lambda *ops: sum(c.get(o, 0) for o in ops)
        print(f"| {obj} | `{name}` | {g('UBLKCP')} | {g('SYNCS')} | {g('DFMA', 'DADD', 'DMUL')} | {g('ATOMS', 'REDS', 'RED', 'ATOMG')} | "
              f"{g('MATCH')} | {g('UTMALDG')} | {g('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCMMA', 'HMMA')} |")
