"""BASELINE config 5 (dense fused scene, 20,000 samples) with the SAMPLES of one cloud sharded over the GPUs of a
box (strong scaling): every rank voxelises the cloud, handles its contiguous share of the samples
(ag_params.shard_index / shard_count) and stores its grasp list into every rank's gather buffer over NVLink
(ag_gather_*).  Launch: python -m torch.distributed.run --nproc-per-node N tools/config5_sharded.py [config] [steps]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from agile_grasp_b200 import api, scenes, shard

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cache = f"/tmp/ag_cfg{cfg}.npz"
if rank == 0 and not os.path.exists(cache):  # one rank renders the scene, the others read it
    pts, size_left, P, S = scenes.config_cloud(cfg)
    np.savez(cache, pts=pts, size_left=size_left)
if world > 1:
    dist.barrier()
_, _, P, S = scenes.config_cloud(2, small=(32, 24, 8))  # parameters only
z = np.load(cache)
pts, size_left = z["pts"], int(z["size_left"])
P.num_samples = scenes.CONFIGS[cfg]["samples"]
P.shard_index, P.shard_count = rank, world
ctx = api.Context(local, P)
svm = api.Svm(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/svm_032015_linear_20_20_same"))
ctx.set_svm(svm)
if world > 1:
    shard.setup_peer_gather(ctx, P.num_samples)
host = torch.from_numpy(pts).pin_memory()
dev = host.cuda()
def step():
    g = ctx.localize_device(dev.data_ptr(), pts.strides[0], pts.shape[0], size_left)
    n_per = ctx.gather_wait()[0] if world > 1 else [len(g)]
    return g, n_per
for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(steps):
    g, n_per = step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    tm = ctx.timings()
    total = int(sum(n_per))
    print(json.dumps({"config": cfg, "n_gpus": world, "points": int(pts.shape[0]), "voxels": tm["n_voxels"], "samples": int(P.num_samples),
                      "hypotheses": total, "per_rank": [int(v) for v in n_per], "ms_per_cloud": float(t.item()) * 1e3 / steps,
                      "hyp_per_s": total * steps / float(t.item()), "rank0_stages_ms": {k: round(tm[k], 4) for k in
                      ("preprocess_ms", "quadric_ms", "sweep_ms", "hog_svm_ms", "total_ms")}}), flush=True)
ctx.set_svm(None)
ctx.close()
del svm
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
