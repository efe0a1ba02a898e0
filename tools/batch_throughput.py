"""BASELINE config 4 on one GPU: a batch of VGA clouds (307,200 points, 2000 samples each, fused linear-SVM scoring)
through ag_localize_batch (4 lanes: clouds in flight on separate streams, each lane replaying its CUDA graph) against
the same clouds through ag_localize one after the other.  Host buffers in, host grasp lists out (wall clock)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agile_grasp_b200 import api, scenes

n_clouds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
distinct = min(n_clouds, 8)
base = []
for k in range(distinct):
    pts, size_left, P, S = scenes.config_cloud(2, scene_offset=k)
    if os.environ.get("AG_PAGEABLE") is None:  # pinned host buffers (what a camera driver / bench.py hands over)
        import torch
        pts = torch.from_numpy(pts).pin_memory().numpy()
    base.append((pts, size_left))
clouds = [base[i % distinct][0] for i in range(n_clouds)]
sls = [base[i % distinct][1] for i in range(n_clouds)]
ctx = api.Context(0, P)
svm = api.Svm(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/svm_032015_linear_20_20_same"))
ctx.set_svm(svm)
for _ in range(3):
    ctx.localize_batch(clouds[:8], sls[:8])
    [ctx.localize(p, s) for p, s in zip(clouds[:2], sls[:2])]
res = {}
for mode in ("sequential", "batch"):
    best = None
    for rep in range(5):
        t0 = time.perf_counter()
        out = ctx.localize_batch(clouds, sls) if mode == "batch" else [ctx.localize(p, s) for p, s in zip(clouds, sls)]
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    hyp = sum(len(g) for g in out)
    res[mode] = dict(clouds=n_clouds, hypotheses=hyp, wall_ms=best * 1e3, ms_per_cloud=best * 1e3 / n_clouds, hyp_per_s=hyp / best)
    print(mode, json.dumps(res[mode]), flush=True)
res["speedup"] = res["batch"]["hyp_per_s"] / res["sequential"]["hyp_per_s"]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/batch_throughput.json", "w"), indent=1)
print("batch / sequential:", round(res["speedup"], 3))
