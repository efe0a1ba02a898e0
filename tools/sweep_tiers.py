"""Hand-sweep stage time and slab-size distribution for configs 2 and 5 (slab = points of the r = 0.08 ball inside
|z_hand| < hand_height, the shared-memory working set of k_hand_sweep; its capacity tiers are sized from this)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
svm_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/svm_032015_linear_20_20_same")
for cfg in (2, 5):
    cache = f"/tmp/ag_cfg{cfg}.npz"
    if not os.path.exists(cache):
        pts, size_left, P, S = scenes.config_cloud(cfg)
        np.savez(cache, pts=pts, size_left=size_left)
    z = np.load(cache)
    pts, size_left = z["pts"], int(z["size_left"])
    _, _, P, _ = scenes.config_cloud(2, small=(32, 24, 8))
    P.num_samples = scenes.CONFIGS[cfg]["samples"]
    ctx = api.Context(0, P)
    svm = api.Svm(svm_path)
    ctx.set_svm(svm)
    ts = []
    for i in range(10):
        g = ctx.localize(pts, size_left)
        t = ctx.timings()
        if i >= 4:
            ts.append((t["sweep_ms"], t["quadric_ms"], t["hog_svm_ms"], t["preprocess_ms"]))
    slab = ctx.sweep_debug(P.num_samples)["num_slab"]
    q = np.quantile(slab, [0.5, 0.9, 0.99, 0.999, 1.0]).astype(int).tolist()
    m = np.median(np.array(ts), axis=0)
    print("config", cfg, "sweep_ms", round(float(m[0]), 4), "quadric_ms", round(float(m[1]), 4), "hog_svm_ms", round(float(m[2]), 4),
          "preprocess_ms", round(float(m[3]), 4), "hyp", len(g), "slab 50/90/99/99.9/max", q,
          "frac>1280", round(float((slab > 1280).mean()), 4), "frac>3072", round(float((slab > 3072).mean()), 5),
          "score checksum", float(np.nansum(g["score"])))
    ctx.set_svm(None)
    ctx.close()
