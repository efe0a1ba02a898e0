import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from agile_grasp_b200 import api, scenes
from oracle import oracle as O
for cfg in (1, 2):
    pts, size_left, P, S = scenes.config_cloud(cfg)
    xo, co = O.preprocess(pts, size_left, P, False)
    idx = O.draw_samples(len(xo), S, P.seed)
    tree = O.Tree(xo)
    c = api.Context(0, P)
    c.set_cloud(xo, co)
    fg = c.fit_quadrics(idx, 0.03)
    fx = O.fit_quadrics(tree, co, idx, 0.03, P, sum_perm=-1, want_params=True)
    ex = fx["frames"]
    d = np.linalg.norm(fg["normal"] - ex["normal"], axis=1)
    ok = ex["num_neighbors"] >= 10
    print("cfg", cfg, "forced" if os.environ.get("AG_FORCE_JACOBI") else "fast", "quantiles", np.quantile(d[ok], [0.5, 0.9, 0.99, 0.999, 1.0]), "n>1e-9:", int((d[ok] > 1e-9).sum()))
    for b in np.argsort(-np.where(ok, d, 0))[:6]:
        print("   ", b, idx[b], ex["num_neighbors"][b], d[b], "eig", fx["eigvals"][b][:4])
    c.close()
