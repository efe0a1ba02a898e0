"""First-light diagnostic: every GPU stage against the CPU oracle on a small scene (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
from agile_grasp_b200.ctypes_defs import *
from oracle import oracle as O

small = (320, 240, 200) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pts, size_left, P, S = scenes.config_cloud(cfg, small=small)
print("cloud", pts.shape, "size_left", size_left, "samples", S)
ctx = api.Context(0, P)
t = time.time(); xyz_g, cam_g = ctx.preprocess(pts, size_left); print("gpu preprocess", time.time() - t, xyz_g.shape)
xyz_o, cam_o = O.preprocess(pts, size_left, P, False)
print("preprocess equal:", xyz_g.shape == xyz_o.shape and (xyz_g.view(np.uint32) == xyz_o.view(np.uint32)).all() and (cam_g == cam_o).all())
tree = O.Tree(xyz_o)
n = len(xyz_o)
idx = O.draw_samples(n, S, P.seed)
bad = 0
for i in idx[:50]:
    a = ctx.radius_search(xyz_o[i], 0.03); b, _ = tree.radius_search(xyz_o[i], 0.03, 0)
    a8 = ctx.radius_search(xyz_o[i], 0.08); b8, _ = tree.radius_search(xyz_o[i], 0.08, 0)
    if not (np.array_equal(a, np.sort(b)) and np.array_equal(a8, np.sort(b8))): bad += 1
print("radius search mismatches:", bad, "of 50")
t = time.time(); fr_g = ctx.fit_quadrics(idx, 0.03); print("gpu quadrics", time.time() - t)
t = time.time(); ro = O.fit_quadrics(tree, cam_o, idx, 0.03, P); fr_o = ro["frames"]; print("oracle quadrics", time.time() - t)
ro2 = O.fit_quadrics(tree, cam_o, idx, 0.03, P, sum_perm=1)["frames"]
print("nn equal:", (fr_g["num_neighbors"] == fr_o["num_neighbors"]).all(), "major equal:", (fr_g["majority_cam"] == fr_o["majority_cam"]).all())
for nm in ("normal", "axis", "binormal"):
    d = np.linalg.norm(fr_g[nm] - fr_o[nm], axis=1); d2 = np.linalg.norm(ro2[nm] - fr_o[nm], axis=1)
    print(nm, "gpu-vs-oracle quantiles", np.quantile(d, [0.5, 0.9, 0.99, 1.0]), "| oracle self-sensitivity", np.quantile(d2, [0.5, 0.9, 0.99, 1.0]))
# sweep with ORACLE frames -> must be bit exact
scam = cam_o[idx]
normals = np.zeros((n, 3)); normals[idx] = fr_o["normal"]
t = time.time(); Ho = O.find_hands(tree, cam_o, idx, fr_o, scam, normals, P); print("oracle hands", time.time() - t, len(Ho))
t = time.time(); g_g = ctx.hand_sweep(idx, fr_o, normals); print("gpu sweep", time.time() - t, len(g_g))
g_o = Ho.grasps
dbg_o = Ho.debug(len(idx)); dbg_g = ctx.sweep_debug(len(idx))
print("slab counts equal:", (dbg_o["num_slab"] == dbg_g["num_slab"]).all(), "status equal:", (dbg_o["status"] == dbg_g["status"]).all())
ok2 = dbg_o["status"] == 2
print("hand idx equal:", (dbg_o["hand_idx"][ok2] == dbg_g["hand_idx"][ok2]).all(), "depth equal:", (dbg_o["depth_steps"][ok2] == dbg_g["depth_steps"][ok2]).all(),
      "fingers equal:", (dbg_o["finger_mask"][ok2] == dbg_g["finger_mask"][ok2]).all())
if len(g_g) == len(g_o):
    for nm in ("axis", "approach", "binormal", "bottom", "surface", "width"):
        print(" ", nm, "max abs diff", np.abs(g_g[nm] - g_o[nm]).max(), "bit-equal", (g_g[nm].view(np.uint64) == g_o[nm].view(np.uint64)).all())
    for nm in ("sample_index", "orientation", "cam_source", "num_points", "half_antipodal", "full_antipodal"):
        print(" ", nm, "equal", (g_g[nm] == g_o[nm]).all())
    imgs_g = api.unpack_images(ctx.images())
    nbad = sum(int((imgs_g[k] != Ho.image(k, P)).any()) for k in range(len(g_o)))
    print("image mismatches:", nbad, "of", len(g_o))
    svm_g = api.Svm("tests/golden/svm_032015_linear_20_20_same"); svm_o = O.Svm("tests/golden/svm_032015_linear_20_20_same")
    keep_o = Ho.classify(svm_o, P); sc_o = Ho.grasps["score"]
    gg, keep_g = ctx.classify(svm_g, g_g)
    print("scores max abs diff", np.abs(gg["score"] - sc_o).max(), "bit-equal", (gg["score"].view(np.uint32) == sc_o.view(np.uint32)).all(), "labels equal", (keep_g == keep_o).all(), "kept", keep_o.sum())
    sc2, desc = ctx.hog_svm(svm_g, ctx.images(), want_descriptors=True)
    d_o = np.stack([O.hog(Ho.image(k, P)) for k in range(len(g_o))])
    print("descriptor bit mismatches", int((desc.view(np.uint32) != d_o.view(np.uint32)).sum()), "max abs", np.abs(desc - d_o).max())
# full path
t = time.time(); g_full = ctx.localize(pts, size_left, idx); print("gpu localize", time.time() - t, len(g_full), ctx.timings())
t = time.time(); g_full = ctx.localize(pts, size_left, idx); print("gpu localize (2nd)", time.time() - t, len(g_full), ctx.timings())
Hf, tm, nv = O.localize(pts, size_left, P, idx, 0, None, True); print("oracle localize", tm, len(Hf))
