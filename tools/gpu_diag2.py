"""Quadric-frame parity statistics: GPU vs oracle(dggev from OpenBLAS A) vs oracle(OpenBLAS B) vs summation permutation."""
import os, sys, time, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
from oracle import oracle as O

small = (640, 480, 2000) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
pts, size_left, P, S = scenes.config_cloud(2, small=small)
P.num_threads = os.cpu_count()
print("cpus", os.cpu_count())
ctx = api.Context(0, P)
xyz, cam = ctx.preprocess(pts, size_left)
n = len(xyz); idx = O.draw_samples(n, S, P.seed)
tree = O.Tree(xyz)
L = O.lib()
provs = O.lapack_providers()
print(provs)
res = {}
for path, sym in provs:
    assert O.set_lapack(path, sym), L.ago_last_error()
    t = time.time(); res[sym] = O.fit_quadrics(tree, cam, idx, 0.03, P)["frames"]; print(sym, "oracle quadrics", time.time() - t)
perm = O.fit_quadrics(tree, cam, idx, 0.03, P, sum_perm=1)["frames"]
t = time.time(); fg = ctx.fit_quadrics(idx, 0.03); print("gpu", time.time() - t)
keys = list(res)
A = res[keys[0]]; B = res[keys[-1]]
def d(x, y, nm): return np.linalg.norm(x[nm] - y[nm], axis=1)
q = [0.5, 0.9, 0.99, 0.999, 1.0]
for nm in ("normal", "axis"):
    print(nm, "gpu vs A      ", np.quantile(d(fg, A, nm), q))
    print(nm, "gpu vs B      ", np.quantile(d(fg, B, nm), q))
    print(nm, "A vs B        ", np.quantile(d(A, B, nm), q))
    print(nm, "B vs B-perm   ", np.quantile(d(B, perm, nm), q))
dn = d(fg, A, "normal"); ab = d(A, B, "normal")
for thr in (1e-5, 1e-6):
    print("frac gpu-vs-A normal <=", thr, (dn <= thr).mean(), "| frac A-vs-B <=", thr, (ab <= thr).mean())
big = dn > 1e-5
print("of the", big.sum(), "samples with gpu-vs-A > 1e-5: A-vs-B quantiles", np.quantile(ab[big], [0, 0.5, 1.0]) if big.any() else None)
cond = ab < 1e-7
print("conditioned subset (A-vs-B < 1e-7):", cond.sum(), "max gpu-vs-A there", dn[cond].max() if cond.any() else None, "quantiles", np.quantile(dn[cond], q))
# full path timing
g = ctx.localize(pts, size_left, idx); 
for _ in range(3): g = ctx.localize(pts, size_left, idx)
print("gpu localize", len(g), ctx.timings())
svm = api.Svm("tests/golden/svm_032015_linear_20_20_same")
gg, keep = ctx.classify(svm, g); print("classify", keep.sum(), ctx.timings()["hog_svm_ms"])
t = time.time(); Hf, tm, nv = O.localize(pts, size_left, P, idx, 0, O.Svm("tests/golden/svm_032015_linear_20_20_same"), True); print("oracle localize all cores", time.time() - t, tm, len(Hf))
go = Hf.grasps
# match hypotheses by (sample_slot, orientation)
ko = set(zip(go["sample_slot"].tolist(), go["orientation"].tolist())); kg = set(zip(gg["sample_slot"].tolist(), gg["orientation"].tolist()))
print("hyp oracle", len(ko), "gpu", len(kg), "common", len(ko & kg))
