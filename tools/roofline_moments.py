"""Roofline of k_taubin_moments at scale: the kernel on ALL voxels of a cloud (the size of the reference's
calculates_antipodal pass, but at the r = 0.03 Taubin radius) and on the big configs.  Algorithmic bytes =
16 B per neighbour + 292 B per sample; duration = CUDA events around the kernel (ag_timings.moments_ms)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes

peak = 6650.0  # fallback of B200_PROFILING.md when MEASURED_PEAKS.json is absent
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for cfg, radius, mode in ((2, 0.03, "samples"), (2, 0.03, "all"), (2, 0.01, "all"), (3, 0.03, "all"), (5, 0.03, "samples"), (5, 0.03, "all")):
    pts, size_left, P, S = scenes.config_cloud(cfg)
    ctx = api.Context(0, P)
    xyz, cam = ctx.preprocess(pts, size_left)
    n = len(xyz)
    if mode == "all":
        idx = np.arange(n, dtype=np.int32)
    else:
        rng = np.random.default_rng(1)
        idx = np.sort(rng.choice(n, min(S, n), replace=False)).astype(np.int32)
    best = None
    for rep in range(4):
        ctx.fit_quadrics(idx, radius)
        t = ctx.timings()
        if best is None or t["moments_ms"] < best["moments_ms"]:
            best = t
    bytes_alg = 16 * best["taubin_neighbor_points"] + 292 * len(idx)
    gbs = bytes_alg / (best["moments_ms"] * 1e-3) / 1e9
    row = dict(config=cfg, radius=radius, mode=mode, n_voxels=n, samples=len(idx), neighbours=best["taubin_neighbor_points"],
               candidates=best["taubin_candidates"], algorithmic_MB=bytes_alg / 1e6, search_ms=best["search_ms"],
               moments_ms=best["moments_ms"], axes_ms=best["axes_ms"], achieved_GBs=gbs, frac_of_peak=gbs / peak,
               achieved_GBs_incl_search=bytes_alg / ((best["moments_ms"] + best["search_ms"]) * 1e-3) / 1e9, peak_GBs=peak)
    rows.append(row)
    print(json.dumps(row), flush=True)
    ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/roofline_moments.json", "w"), indent=1)
