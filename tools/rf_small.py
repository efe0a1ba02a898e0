import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes
pts, size_left, P, S = scenes.config_cloud(2)
ctx = api.Context(0, P)
xyz, cam = ctx.preprocess(pts, size_left)
n = len(xyz)
for mode in ("samples", "all"):
    idx = np.arange(n, dtype=np.int32) if mode == "all" else np.sort(np.random.default_rng(1).choice(n, 2000, replace=False)).astype(np.int32)
    best = None
    for rep in range(5):
        ctx.fit_quadrics(idx, 0.03); t = ctx.timings()
        if best is None or t["moments_ms"] < best["moments_ms"]: best = t
    b = 16 * best["taubin_neighbor_points"] + 292 * len(idx)
    print(mode, "search_ms", round(best["search_ms"], 4), "moments_ms", round(best["moments_ms"], 4), "moments GB/s", round(b / best["moments_ms"] / 1e6, 1), "search+moments GB/s", round(b / (best["moments_ms"] + best["search_ms"]) / 1e6, 1), "axes_ms", round(best["axes_ms"], 4))
