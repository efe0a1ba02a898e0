import copy, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from agile_grasp_b200 import api, scenes
from oracle import oracle as O
pts, size_left, P, S = scenes.config_cloud(2, small=(320, 240, 150))
xyz, cam = O.preprocess(pts, size_left, P, False)
tree = O.Tree(xyz)
P.deterministic_normals = 0
c = api.Context(0, P)
c.set_cloud(xyz, cam)
idx = np.arange(0, len(xyz), 3, dtype=np.int32)
fg = c.fit_quadrics(idx, 0.03)
ex = O.fit_quadrics(tree, cam, idx, 0.03, P, sum_perm=-1)["frames"]
d = np.linalg.norm(fg["normal"] - ex["normal"], axis=1)
da = np.linalg.norm(fg["axis"] - ex["axis"], axis=1)
ok = fg["num_neighbors"] >= 10
print("quantiles normal", np.quantile(d[ok], [0.5, 0.9, 0.99, 0.999, 1.0]))
print("quantiles axis", np.quantile(da[ok], [0.5, 0.9, 0.99, 0.999, 1.0]))
bad = np.nonzero(ok & (d > 1e-9))[0]
print("bad", len(bad), "of", ok.sum())
P1 = copy.copy(P); P1.deterministic_normals = 1
c.set_params(P1); c.set_cloud(xyz, cam)
fd = c.fit_quadrics(idx, 0.03)
exd = O.fit_quadrics(tree, cam, idx, 0.03, P1, sum_perm=-1)["frames"]
for b in bad[:12]:
    print(b, idx[b], fg["num_neighbors"][b], "d", d[b], "da", da[b], "gpu n", fg["normal"][b], "ora n", ex["normal"][b], "gpu ax", fg["axis"][b], "ora ax", ex["axis"][b],
          "det-mode diff", np.linalg.norm(fd["normal"][b]-exd["normal"][b]), "dot", float(fg["normal"][b] @ ex["normal"][b]))
# small-scene e2e stats for threshold setting (tests/test_gpu_parity.py)
svm = "tests/golden/svm_032015_linear_20_20_same"
c.set_params(P1)
s_idx = O.draw_samples(len(xyz), S, P1.seed)
g = c.localize(pts, size_left, s_idx)
gg, keep = c.classify(api.Svm(svm), g)
H, tm, nv = O.localize(pts, size_left, P1, s_idx, 0, O.Svm(svm), False)
go = H.grasps
ko = {(a, b): i for i, (a, b) in enumerate(zip(go["sample_index"].tolist(), go["orientation"].tolist()))}
kg = {(a, b): i for i, (a, b) in enumerate(zip(gg["sample_index"].tolist(), gg["orientation"].tolist()))}
common = sorted(set(ko) & set(kg))
io = np.array([ko[k] for k in common]); ig = np.array([kg[k] for k in common])
print("small e2e: common", len(common), len(ko), len(kg), "same_img", (gg["num_points"][ig] == go["num_points"][io]).mean(),
      "labels", (gg["label"][ig] == go["label"][io]).mean())
for s_ in ("small", "two"):
    pts2, sl2, P2, S2 = scenes.config_cloud(2 if s_ == "small" else 3, small=(320, 240, 150) if s_ == "small" else (200, 150, 120))
    x2, c2 = O.preprocess(pts2, sl2, P2, False); t2 = O.Tree(x2); i2 = O.draw_samples(len(x2), S2, P2.seed)
    c.set_params(P2); c.set_cloud(x2, c2)
    f_g = c.fit_quadrics(i2, 0.03); f_r = O.fit_quadrics(t2, c2, i2, 0.03, P2)["frames"]
    dr = np.linalg.norm(f_g["normal"] - f_r["normal"], axis=1)
    print(s_, "frac<=1e-5 vs dggev", (dr <= 1e-5).mean(), "median", np.median(dr))
