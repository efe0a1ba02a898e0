"""Sample sharding across ranks and the all-gather of the grasp list.

Every sample's work is independent (the reference's two hot loops are plain `omp parallel for`,
hand_search.cpp:77-80,135-138, followed by a stable concatenation, :194-200), so samples are split
into contiguous ranges, one per rank; the voxelised cloud and its row index are replicated (each rank
runs the deterministic preprocessing itself).  The only exchange step is one all-gather of the
fixed-stride grasp records, after which every rank holds the reference's sample-major ordering.
Works with any torch.distributed backend: NCCL on device tensors (bench.py), gloo on CPU (tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .ctypes_defs import GRASP_DTYPE


def shard_positions(n_items, rank, world, interleave=False):
    """positions (in the full sample list) of rank's share: a contiguous range (ag_params.shard_interleave = 0) or
    every world-th sample starting at rank (= 1) — the same partition api.cu applies"""
    if interleave:
        return np.arange(rank, n_items, world)
    lo, hi = shard_range(n_items, rank, world)
    return np.arange(lo, hi)


def merge_by_sample(parts):
    """host statement of the gather's merge: lists whose sample_slot is the position in the full sample list ->
    one list ordered by (sample_slot, orientation) = the reference's stable concat (hand_search.cpp:194-200)"""
    parts = [p for p in parts if len(p)]
    if not parts:
        return np.zeros(0, GRASP_DTYPE)
    cat = np.concatenate(parts)
    order = np.lexsort((cat["orientation"], cat["sample_slot"]))
    return cat[order]


def shard_range(n_items, rank, world):
    """contiguous [lo, hi) of rank's share; sizes differ by at most one, order preserved"""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


def all_gather_grasps(local, device=None, group=None):
    """local: numpy structured array (GRASP_DTYPE) of this rank's hypotheses, in sample order.
    Returns the concatenation over ranks in rank order (= global sample order)."""
    world = dist.get_world_size(group)
    dev = torch.device("cpu") if device is None else device
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    item = GRASP_DTYPE.itemsize
    buf = np.zeros(cap * item, np.uint8)
    buf[: len(local) * item] = np.frombuffer(np.ascontiguousarray(local).tobytes(), np.uint8)
    send = torch.from_numpy(buf).to(dev)
    recv = torch.empty(world * cap * item, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    host = recv.cpu().numpy().reshape(world, cap * item)
    parts = [np.frombuffer(host[r, : counts[r] * item].tobytes(), dtype=GRASP_DTYPE) for r in range(world)]
    out = np.concatenate(parts) if parts else np.zeros(0, GRASP_DTYPE)
    out = out.copy()
    out["image_id"] = -1  # images stay on the producing rank
    return out, counts


EXPORT_HEADER_BYTES = 16


def export_buffer_bytes(num_samples):
    """size of the per-rank device export buffer of ag_set_export_buffer for `num_samples` samples"""
    return EXPORT_HEADER_BYTES + 8 * int(num_samples) * GRASP_DTYPE.itemsize


def all_gather_export(send, recv, group=None):
    """ONE fixed-size collective on device memory: every rank contributes its export buffer
    ([n_hyp, n_vox, n_samples, error][records], written by the library inside ag_localize) and receives
    all of them; no host copy, no size exchange.  send: uint8 CUDA tensor (export buffer), recv: uint8 CUDA
    tensor of world * len(send) bytes.  Returns recv viewed as (world, bytes)."""
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.view(dist.get_world_size(group), -1)


def parse_export(buf_u8):
    """host-side decode of one rank's export buffer (numpy uint8 array) -> (header dict, records)"""
    hdr = np.frombuffer(buf_u8[:EXPORT_HEADER_BYTES].tobytes(), dtype=np.int32)
    n = int(hdr[0])
    item = GRASP_DTYPE.itemsize
    recs = np.frombuffer(buf_u8[EXPORT_HEADER_BYTES:EXPORT_HEADER_BYTES + n * item].tobytes(), dtype=GRASP_DTYPE)
    return dict(n_hyp=n, n_vox=int(hdr[1]), n_samples=int(hdr[2]), error=int(hdr[3])), recs


GATHER_SLOT_HEADER = 32


def setup_peer_gather(ctx, num_samples, group=None):
    """Fused export + all-gather over NVLink peer memory (include/ag_b200.h, ag_gather_*): every rank creates
    its gather buffer, the CUDA IPC handles are exchanged once through torch.distributed, every rank maps
    every buffer.  After this, every ctx.localize*() is a collective step: it stores the rank's grasp list into
    all ranks' buffers, waits (on the device) for the other ranks' lists and merges them;
    ctx.gather_result() returns the merged list - no collective call per step."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handle = ctx.gather_create(num_samples, world, rank)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    ctx.gather_connect(handles)
    dist.barrier(group=group)


def read_gathered(n_hyp, slots_ptr, slot_bytes):
    """host copy of the gathered lists (debug / verification): list of GRASP_DTYPE arrays, one per rank"""
    import ctypes as C
    rt = C.CDLL("libcudart.so")
    out = []
    item = GRASP_DTYPE.itemsize
    for r, n in enumerate(n_hyp):
        buf = np.zeros(max(n, 0) * item, np.uint8)
        if n > 0:
            rc = rt.cudaMemcpy(C.c_void_p(buf.ctypes.data), C.c_void_p(slots_ptr + r * slot_bytes + GATHER_SLOT_HEADER),
                               C.c_size_t(n * item), C.c_int(2))
            assert rc == 0, rc
        out.append(np.frombuffer(buf.tobytes(), dtype=GRASP_DTYPE))
    return out
