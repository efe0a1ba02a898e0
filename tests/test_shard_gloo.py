"""N>1 host logic on CPU: contiguous sample sharding + all-gather of grasp records (gloo, world_size 2)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from agile_grasp_b200.ctypes_defs import GRASP_DTYPE
from agile_grasp_b200.shard import shard_range


def test_shard_ranges_partition():
    for n in (0, 1, 7, 2000, 20001):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from agile_grasp_b200.shard import all_gather_grasps, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a fake global result: hypothesis k belongs to sample k // 2; each rank owns a contiguous sample range
    S = 11
    allg = np.zeros(2 * S - 3, GRASP_DTYPE)
    allg["sample_slot"] = np.arange(len(allg)) // 2
    allg["orientation"] = np.arange(len(allg)) % 2
    allg["width"] = np.arange(len(allg)) * 0.5
    lo, hi = shard_range(S, rank, world)
    local = allg[(allg["sample_slot"] >= lo) & (allg["sample_slot"] < hi)]
    merged, counts = all_gather_grasps(local)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    np.save(os.path.join(out_dir, f"expect_{rank}.npy"), allg)
    assert sum(counts) == len(allg)
    dist.destroy_process_group()


def test_all_gather_grasps_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        m = np.load(tmp_path / f"merged_{r}.npy")
        e = np.load(tmp_path / f"expect_{r}.npy")
        assert len(m) == len(e)
        for nm in ("sample_slot", "orientation", "width"):
            assert np.array_equal(m[nm], e[nm])


def test_shard_range_matches_the_library_partition():
    """shard_range (python helper) and ag_params.shard_index / shard_count (api.cu: k_lo = S * i / n) cut the same
    contiguous shares: they tile [0, S) in order, sizes differ by at most one."""
    from agile_grasp_b200.shard import shard_range
    for S in (0, 1, 7, 2000, 20000):
        for world in (1, 2, 3, 8):
            edges = [shard_range(S, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == S
            for r in range(world):
                lo, hi = edges[r]
                assert lo == (S * r) // world and hi == (S * (r + 1)) // world
                if r:
                    assert lo == edges[r - 1][1]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_shares_merge_back():
    """shard_positions / merge_by_sample (host statement of the library's partition and of the gather's merge):
    contiguous and interleaved shares tile the sample list; lists keyed by the position in the full list merge back
    into sample-major order whatever the partition was."""
    from agile_grasp_b200.shard import merge_by_sample, shard_positions
    rng = np.random.default_rng(3)
    for S in (1, 7, 2000):
        allg = np.zeros(0, GRASP_DTYPE)
        per_sample = rng.integers(0, 4, S)
        recs = []
        for k in range(S):
            for o in sorted(rng.choice(8, per_sample[k], replace=False)):
                g = np.zeros(1, GRASP_DTYPE)
                g["sample_slot"], g["orientation"], g["width"] = k, o, k + 0.1 * o
                recs.append(g)
        allg = np.concatenate(recs) if recs else allg
        for world in (1, 2, 3, 8):
            for il in (False, True):
                pos = [shard_positions(S, r, world, il) for r in range(world)]
                assert sorted(np.concatenate(pos).tolist()) == list(range(S))
                parts = [allg[np.isin(allg["sample_slot"], p)] for p in pos]
                m = merge_by_sample(parts)
                assert m.tobytes() == allg.tobytes()
