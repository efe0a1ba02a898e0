import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
from oracle import oracle as O
for cfg, small in ((2, (320, 240, 150)), (3, (200, 150, 120)), (2, (640, 480, 2000))):
    pts, size_left, P, S = scenes.config_cloud(cfg, small=small)
    xyz, cam = O.preprocess(pts, size_left, P)
    tree = O.Tree(xyz); idx = O.draw_samples(len(xyz), S, P.seed)
    P.num_threads = 16
    ctx = api.Context(0, P); ctx.set_cloud(xyz, cam)
    fg = ctx.fit_quadrics(idx, 0.03)
    ex = O.fit_quadrics(tree, cam, idx, 0.03, P, sum_perm=-1, want_params=True)
    fe = ex["frames"]
    dn = np.linalg.norm(fg["normal"] - fe["normal"], axis=1); da = np.linalg.norm(fg["axis"] - fe["axis"], axis=1)
    print("cfg", cfg, small, "normal quantiles", np.quantile(dn, [0.5, 0.9, 0.99, 1.0]), "axis", np.quantile(da, [0.5, 0.9, 0.99, 1.0]))
    for k in np.nonzero((dn > 1e-9) | (da > 1e-7))[0][:12]:
        print("  k", k, "idx", idx[k], "nn", fe["num_neighbors"][k], "dn", dn[k], "da", da[k], "gpu n", fg["normal"][k], "ex n", fe["normal"][k], "gpu a", fg["axis"][k], "ex a", fe["axis"][k])
    ctx.close()
