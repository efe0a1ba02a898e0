#!/usr/bin/env python
"""bench.py — grasp hypotheses/sec on the BASELINE.json configuration (307,200-point organised cloud,
2000 samples, linear SVM), through the C ABI of libag_b200.so.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2] [--normals rand|det]

A step = one pass of the hot path (ag_localize + ag_classify = Localization::localizeHands +
predictAntipodalHands) over one synthetic cloud per rank.  N>1 is launched by torchrun, one rank per
GPU: every rank processes one cloud per step (weak scaling; the same two scenes alternate on every rank and
in the CPU arm, so hyp/s ratios are time ratios) and the grasp lists are exchanged by peer stores fused into
the export kernel, so every rank holds the whole grasp list.

Normal mode: the reference's PRODUCTION default (hand_search.h:84 is_deterministic = false: normals from 50
picks rand() % n, quadric.cpp:177-192) in both arms (--normals rand); the deterministic all-neighbour mode of
the reference's component tests is reported next to it (`other_mode` sub-objects; --normals det makes it
the headline instead).

  value : hyp/s with the cloud already resident in HBM (ag_localize_device), timed with CUDA events on
          the library's own stream around the whole call (ag_timings.total_ms), max over ranks.  The library default:
          no per-stage event records inside the pipeline (ag_set_stage_timing off).
  e2e   : hyp/s through the host-buffer entry points (pinned host cloud in, host grasp list out),
          wall clock around the calls; H2D of the cloud and D2H of the records are inside.
  stage_pass : a second pass of the same K steps with ag_set_stage_timing(1) (a dozen event-record nodes in the CUDA
          graph, ~0.04 ms per step): the source of stages_ms and of the kernel durations behind the two rooflines.
  roofline : the Taubin stage (radius search + moment accumulation): algorithmic bytes (16 B per neighbour +
          292 B out per sample) over the CUDA-event duration of its kernel(s), against MEASURED_PEAKS.json;
          traffic = the kernel's DRAM bytes per launch from the committed ncu capture (profiles/taubin_traffic.json).
  roofline_step : the same for the kernel with the largest share of the step (k_hand_sweep).
  strong_scaling : ONE cloud per step, its samples sharded over the N ranks (ag_params.shard_index / shard_count,
          interleaved shares), every rank ending up with the whole merged grasp list (peer stores + merge inside
          ag_localize): BASELINE configs 5 (2 M points, 20,000 samples) and 4 (VGA clouds, 2000 samples), ms/cloud.
  cpu_baseline : the CPU oracle (a port of the reference path, OpenMP over samples like the reference)
          on this box's host cores, both normal modes.
--impl reference runs only that CPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grasp hypotheses/sec on 307k-pt cloud, 2000 samples; ms/cloud end-to-end"
SVM_PATH = os.path.join(ROOT, "tests", "golden", "svm_032015_linear_20_20_same")
POLY_XZ = os.path.join(ROOT, "tests", "golden", "svm_032015_20_20_same.xz")
SCENES = (0, 1)  # scene offsets of the pool: the same in both arms and on every rank
MODE_NAME = {0: "production (50 picks rand() % n, hand_search.h:84 is_deterministic = false)",
             1: "deterministic (all neighbours, as src/tests/test_taubin.cpp:58)"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML polled from a thread
    every 2 ms (the timed region of a default run is ~40 ms, shorter than nvidia-smi's start-up), with
    `nvidia-smi --query-gpu -lms` as the fallback when the NVML binding is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows = []   # (host time, sm MHz, max MHz, [reason flags])
        self.index = index
        self.t_mark = None
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.source = None

    def mark(self):
        """start of the timed region: only samples from here on are reported"""
        self.t_mark = time.perf_counter()

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            # (CUDA_VISIBLE_DEVICES remaps CUDA ordinals; NVML enumerates physical devices)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            h = N.nvmlDeviceGetHandleByIndex(phys)
            mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = [N.nvmlClocksThrottleReasonHwSlowdown, N.nvmlClocksThrottleReasonHwThermalSlowdown,
                    N.nvmlClocksThrottleReasonSwThermalSlowdown, N.nvmlClocksThrottleReasonSwPowerCap]

            def poll():
                while not self.stop_flag:
                    try:
                        sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                        r = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.perf_counter(), sm, mx, [bool(r & b) for b in bits]))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "NVML, 2 ms period"
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi -lms 20"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            c = [v.strip() for v in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(c[0]), float(c[1]),
                                  [v.lower().startswith("active") for v in c[3:7]]))
            except Exception:
                continue

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread:
            self.thread.join(timeout=1)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        rows = [r for r in self.rows if self.t_mark is None or r[0] >= self.t_mark]
        window = "timed region"
        if not rows:  # the timed region was shorter than one sampling period: report the warm-up samples
            rows, window = self.rows, "warm-up (timed region shorter than the sampling period)"
        reasons = sorted({nm for r in rows for nm, f in zip(self.NAMES, r[3]) if f})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": max(r[2] for r in rows),
                "reasons": reasons, "samples": len(rows), "window": window, "source": self.source}


def bind_to_gpu_cpus(index):
    """N > 1: run this rank on the CPUs next to its GPU (NVML's ideal affinity) before any pinned buffer is allocated, so
    that the pinned host clouds live on the GPU's NUMA node and eight concurrent H2D copies do not share one socket's
    memory controllers.  Best effort."""
    try:
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        N.nvmlDeviceSetCpuAffinity(N.nvmlDeviceGetHandleByIndex(phys))
        return True
    except Exception:
        return False


def make_cloud(config, scene_offset):
    from agile_grasp_b200 import scenes
    pts, size_left, P, S = scenes.config_cloud(config, scene_offset=scene_offset)
    return np.ascontiguousarray(pts), size_left, P, S


def poly_model_path():
    """the launch-file default model (launch/single_camera_grasps.launch:6), committed xz-compressed"""
    import lzma
    import tempfile
    d = tempfile.mkdtemp(prefix="ag_poly_")
    p = os.path.join(d, "svm_032015_20_20_same")
    with open(p, "wb") as f:
        f.write(lzma.open(POLY_XZ).read())
    return p


def cpu_arm(clouds, det, steps, warmup, svm_path=SVM_PATH):
    """The oracle port of the reference path on all host threads over the scene pool; one cloud per step.
    Returns (hyp/s, ms per cloud, mean hypotheses per cloud, cores)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    svm = O.Svm(svm_path)
    times, hyps = [], []
    for it in range(warmup + steps):
        pts, size_left, P = clouds[it % len(clouds)]
        P.num_threads = cores
        P.deterministic_normals = det
        t0 = time.perf_counter()
        H, tm, nv = O.localize(pts, size_left, P, None, 0, svm, True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            hyps.append(len(H))
        del H
    return float(np.sum(hyps) / np.sum(times)), 1e3 * float(np.mean(times)), float(np.mean(hyps)), cores


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on all host threads (same scene pool as the GPU arm)."""
    if rank != 0:
        return
    det = 1 if args.normals == "det" else 0
    clouds = []
    for k in SCENES:
        pts, size_left, P, S = make_cloud(args.config, k)
        clouds.append((pts, size_left, P))
    value, ms, hyp, cores = cpu_arm(clouds, det, args.steps, args.warmup)
    o_value, o_ms, o_hyp, _ = cpu_arm(clouds, 1 - det, max(1, min(args.steps, 2)), 0)
    n_pts = clouds[0][0].shape[0]
    sample = (f"{args.steps} full clouds of the bench workload after {args.warmup} untimed one(s) (scenes {list(SCENES)} "
              "alternating, as the GPU arm), "
              f"OpenMP over samples on all {cores} host threads; std::set voxelisation and kd-tree as the reference; "
              "SVM parsed once")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "hyp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config{args.config}: organised cloud of {n_pts} pts, {S} samples, linear SVM",
                   "normal_mode": MODE_NAME[det], "scenes": list(SCENES), "hypotheses_per_step": hyp},
        "cpu_baseline": {"value": value, "unit": "hyp/s", "cores": cores, "kind": "port", "normal_mode": MODE_NAME[det],
                         "ms_per_cloud": ms, "sample": sample,
                         "other_mode": {"normal_mode": MODE_NAME[1 - det], "value": o_value, "ms_per_cloud": o_ms}},
        "e2e": {"value": value, "unit": "hyp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.poly:
        pv, pms, _, _ = cpu_arm(clouds[:1], det, 1, 0, svm_path=poly_model_path())
        line["cpu_baseline"]["poly_svm"] = {"value": pv, "ms_per_cloud": pms}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--normals", default="rand", choices=["rand", "det"],
                    help="normal mode of the headline numbers: rand = the reference's production default")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--poly", action="store_true", help="reference arm: also time one cloud with the launch-file POLY model")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (profiling runs)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU grasp-list exchange: NVLink peer stores fused into the export kernel, or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from agile_grasp_b200 import api, shard
    from agile_grasp_b200.ctypes_defs import GRASP_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    bound = bind_to_gpu_cpus(local_rank) if world > 1 else False
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    det = 1 if args.normals == "det" else 0

    # the scene pool: successive steps do not see identical inputs; every rank (and the CPU arm) uses the same set
    pool = []
    for k in SCENES:
        pts, size_left, P, S = make_cloud(args.config, scene_offset=k)
        P.deterministic_normals = det
        pinned = torch.from_numpy(pts).pin_memory()
        pool.append(dict(host=pinned, dev=pinned.to(dev), size_left=size_left, P=P, n=pts.shape[0],
                         stride=pts.strides[0]))
    # the library default: no per-stage event records inside the pipeline (they cost ~16 us per call as graph nodes);
    # the stage split and the roofline kernel durations are measured in a second pass of the same steps with them on
    ctx = api.Context(local_rank, pool[0]["P"], stage_timing=False)
    svm = api.Svm(SVM_PATH)
    ctx.set_svm(svm)  # score inside ag_localize; ag_classify then returns the cached decision values
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    item = GRASP_DTYPE.itemsize

    peer_gather = world > 1 and args.gather == "peer"
    if peer_gather:
        shard.setup_peer_gather(ctx, pool[0]["P"].num_samples)
    elif world > 1:
        nbytes = shard.export_buffer_bytes(pool[0]["P"].num_samples)
        send = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        recv = torch.zeros(nbytes * world, dtype=torch.uint8, device=dev)
        ctx.set_export_buffer(send.data_ptr(), nbytes)
    last_gather = [None]

    def gather(local_g):
        if world == 1:
            return len(local_g)
        if peer_gather:  # the export kernel already stored this rank's list into every rank's buffer
            last_gather[0] = ctx.gather_wait()
            return last_gather[0]
        return shard.all_gather_export(send, recv)  # counts are read after the timed region

    def step_device(c):
        g = ctx.localize_device(c["dev"].data_ptr(), c["stride"], c["n"], c["size_left"])
        t1 = ctx.timings()
        g, keep = ctx.classify(svm, g)
        t2 = ctx.timings()
        return g, t1, t2

    def step_host(c):
        g = ctx.localize(c["host"].numpy(), c["size_left"])
        g, keep = ctx.classify(svm, g)
        return g

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def prime():
        # every input buffer goes through the eager call and the CUDA-graph capture once (untimed), so that
        # warm-up and timed steps all run the steady-state (replay) path
        for c in pool:
            for _ in range(2):
                gather(step_device(c)[0])
                gather(step_host(c))

    def timed_device(steps):
        dev_ms, hyps, comm_ms, launches, recs = [], [], [], [], []
        barrier()
        wall0 = time.perf_counter()
        for k in range(steps):
            c = pool[k % len(pool)]
            flush.fill_(k & 0xFF)  # L2 flush between timed iterations (outside the per-step event brackets)
            torch.cuda.synchronize()
            g, t1, t2 = step_device(c)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tg0 = time.perf_counter()
            e0.record()
            gather(g)
            e1.record()
            torch.cuda.synchronize()
            # peer gather: stores, wait for the slowest peer and merge all run inside ag_localize's stream (in total_ms);
            # the NCCL variant is a separate collective, timed here
            cm = 0.0 if (world == 1 or peer_gather) else e0.elapsed_time(e1)
            dev_ms.append(t1["total_ms"] + cm)  # scoring is fused into ag_localize (ag_set_svm): inside total_ms
            comm_ms.append(cm)
            hyps.append(len(g))
            launches.append(t2["kernel_launches"])
            recs.append(t1)
        barrier()
        return dict(dev_ms=dev_ms, hyps=hyps, comm_ms=comm_ms, launches=launches, t=recs,
                    wall=time.perf_counter() - wall0)

    def timed_host(steps):
        e2e_s, e2e_h = [], []
        g = None
        barrier()
        for k in range(steps):
            c = pool[k % len(pool)]
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g = step_host(c)
            gather(g)
            e2e_s.append(time.perf_counter() - t0)
            e2e_h.append(len(g))
        barrier()
        return e2e_s, e2e_h, g

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prime()
    for w in range(args.warmup):
        gather(step_device(pool[w % len(pool)])[0])
        gather(step_host(pool[w % len(pool)]))
    sampler.mark()
    D = timed_device(args.steps)
    e2e_s, e2e_h, g = timed_host(args.steps)
    # second pass, same steps, per-stage events on: stage split + CUDA-event durations of the roofline kernels
    ctx.set_stage_timing(True)
    prime()
    DS = timed_device(args.steps)
    ctx.set_stage_timing(False)
    clocks = sampler.stop() if rank == 0 else None

    gathered = None
    if peer_gather:  # verify the last exchange on the host (outside the timed region): every rank's list is there
        n_per, slots_ptr, slot_bytes = last_gather[0]
        lists = shard.read_gathered(n_per, slots_ptr, slot_bytes)
        assert n_per[rank] == e2e_h[-1], (n_per, e2e_h[-1])
        for nm in ("sample_index", "orientation", "score", "label", "width", "surface", "bottom"):
            assert np.array_equal(lists[rank][nm], g[nm]), "own slot differs from the returned list: " + nm
        for r in range(world):  # every rank's list is well formed (sample-major order)
            assert np.all(np.diff(lists[r]["sample_index"]) >= 0)
        n_per2, merged, _ = ctx.gather_result()
        assert n_per2 == n_per and len(merged) == sum(n_per)
        off = 0  # independent calls (one cloud per rank): the merged list is rank-major, each part sample-major
        for r in range(world):
            part = merged[off:off + n_per[r]]
            assert np.all(np.diff(part["sample_slot"].astype(np.int64) * 8 + part["orientation"]) > 0)
            off += n_per[r]
        gathered = {"per_rank": n_per, "merged": int(len(merged)),
                    "mode": "peer stores over NVLink + on-device merge inside ag_localize (ag_gather_*), no collective per step"}
    elif world > 1:  # decode the last all-gather on the host (outside the timed region)
        host = recv.cpu().numpy().reshape(world, -1)
        parts = [shard.parse_export(host[r]) for r in range(world)]
        gathered = {"per_rank": [p[0]["n_hyp"] for p in parts], "errors": [p[0]["error"] for p in parts],
                    "mode": "NCCL all_gather_into_tensor of the fixed-size export buffer"}
        assert parts[rank][0]["n_hyp"] == e2e_h[-1], (parts[rank][0], e2e_h[-1])

    # ---- reduce over ranks: time = max, hypotheses = sum
    t_dev = torch.tensor([sum(D["dev_ms"]), sum(e2e_s) * 1e3], dtype=torch.float64, device=dev)
    n_h = torch.tensor([sum(D["hyps"]), sum(e2e_h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_h, op=dist.ReduceOp.SUM)
    t_dev, n_h = t_dev.cpu().numpy(), n_h.cpu().numpy()

    line = None
    if rank == 0:
        T = DS["t"]  # (the pass with the stage events)
        mean = lambda nm: float(np.mean([t[nm] for t in T]))  # noqa: E731
        value = n_h[0] / (t_dev[0] * 1e-3)
        e2e = n_h[1] / (t_dev[1] * 1e-3)
        peak, peak_src = measured_peak()
        # graded stage: radius search + moment accumulation of the Taubin fit (one fused kernel when moments_ms == 0)
        tb_bytes = [16 * t["taubin_neighbor_points"] + 292 * t["n_samples"] for t in T]
        tb_ms = [t["search_ms"] + t["moments_ms"] for t in T]
        achieved = float(np.sum(tb_bytes) / (np.sum(tb_ms) * 1e-3) / 1e9) if np.sum(tb_ms) > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "taubin_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # the kernel with the largest share of the step: the hand sweep (16 B per point of the r = 0.08 ball in,
        # 160 B record + 1000 B grasp image out per hypothesis)
        sw_bytes = [16 * t["hand_neighbor_points"] + 1160 * t["n_hyp"] for t in T]
        sw_ms = [t["sweep_ms"] for t in T]
        sw_ach = float(np.sum(sw_bytes) / (np.sum(sw_ms) * 1e-3) / 1e9) if np.sum(sw_ms) > 0 else 0.0
        c0 = pool[0]
        stages = {nm: mean(nm) for nm in ("preprocess_ms", "normals_all_ms", "quadric_ms", "search_ms", "moments_ms",
                                          "axes_ms", "sweep_ms", "hog_svm_ms", "d2h_ms")}
        # (one fused kernel: the "moments" interval is only the gap between two event records)
        fused = float(np.sum([t["moments_ms"] for t in T])) < 0.5 * float(np.sum([t["search_ms"] for t in T]))
        line = {
            "metric": METRIC, "value": float(value), "unit": "hyp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(t_dev[0] / args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config{args.config}: one organised cloud ({c0['n']} pts, {c0['stride']} B/pt; 640x480 "
                                   f"per view) per rank per step, {c0['P'].num_samples} samples, linear SVM "
                                   "svm_032015_linear_20_20_same",
                       "normal_mode": MODE_NAME[det], "scenes": list(SCENES),
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "launch": "CUDA graph replay (captured during untimed priming calls)",
                       "hypotheses_per_step": float(n_h[0] / args.steps / world),
                       "multi_gpu": ("one cloud per rank per step (the same scene pool on every rank); grasp lists "
                                     "exchanged by " + ("NVLink peer stores fused into the export kernel" if peer_gather
                                                        else "an NCCL all-gather")) if world > 1 else "single GPU",
                       "timer": "CUDA events on the library stream around the whole call (ag_timings.total_ms), max over ranks"},
            "e2e": {"value": float(e2e), "unit": "hyp/s", "ms_per_cloud": float(t_dev[1] / args.steps),
                    "h2d_bytes_per_step": int(c0["n"] * c0["stride"]),
                    "d2h_bytes_per_step": int(np.mean(e2e_h) * item + 64),
                    "timer": "wall clock around ag_localize + ag_classify (+ gather wait), pinned host cloud"
                             + (", rank bound to its GPU's CPUs" if bound else "")},
            "gpu_launches": int(np.sum(D["launches"])),
            "roofline": {"bound": "hbm",
                         "kernel": "k_ball_moments (radius search + Taubin moments, one kernel)" if fused
                         else "k_ball_search + k_taubin_moments",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": float(np.mean(tb_bytes)), "launch_ms": float(np.mean(tb_ms)),
                         "note": "16 B per neighbour of the r = 0.03 ball + 292 B of moments per sample; a 2000-sample "
                                 "launch is latency bound, roofline_at_scale is the same stage on a launch that fills "
                                 "the machine"},
            "roofline_step": {"bound": "hbm", "kernel": "k_hand_sweep",
                              "share_of_step": float(np.sum(sw_ms) / np.sum([t["total_ms"] for t in T])),
                              "achieved": sw_ach, "peak": peak, "unit": "GB/s", "frac": sw_ach / peak,
                              "algorithmic_bytes_per_launch": float(np.mean(sw_bytes)), "launch_ms": float(np.mean(sw_ms)),
                              "note": "16 B per point of the r = 0.08 ball + 1160 B per hypothesis out; the ball is "
                                      "L2 resident (1.3 MB cloud), the kernel is issue / latency bound"},
            "stages_ms": stages,
            "stage_pass": {"ms_per_step": float(np.mean(DS["dev_ms"])), "steps": args.steps,
                           "note": "stages_ms, roofline and roofline_step come from a second pass of the same steps with "
                                   "ag_set_stage_timing(1): a dozen event-record nodes inside the CUDA graph; the headline "
                                   "value / e2e run without them (the library default)"},
            "comm_ms": float(np.mean(D["comm_ms"])), "gathered_last_step": gathered,
            "wall_ms_per_step_incl_flush": float(1e3 * D["wall"] / args.steps),
            "clocks": clocks,
        }

    # ---- strong scaling: the samples of ONE cloud sharded over the ranks (north_star: "samples shard naturally
    # across the 8 GPUs"), every rank ending with the whole merged list; runs at every N (N = 1: the unsharded call)
    if not args.no_extras:
        strong = {}
        for cfg, n_clouds, k_steps in ((5, 1, 5), (4, 8, 2)):
            try:
                clouds = []
                for k in range(n_clouds if cfg == 4 else 1):
                    pts, size_left, Ps, S = make_cloud(5 if cfg == 5 else 2, scene_offset=k % 4)
                    Ps.deterministic_normals = det
                    Ps.shard_index, Ps.shard_count, Ps.shard_interleave = rank, world, 1
                    pin = torch.from_numpy(pts).pin_memory()
                    clouds.append(dict(host=pin, dev=pin.to(dev), size_left=size_left, P=Ps, n=pts.shape[0],
                                       stride=pts.strides[0]))
                c2 = api.Context(local_rank, clouds[0]["P"], stage_timing=False)
                c2.set_svm(svm)
                if world > 1:
                    shard.setup_peer_gather(c2, clouds[0]["P"].num_samples)
                for _ in range(3):
                    for c in clouds:
                        c2.localize_device(c["dev"].data_ptr(), c["stride"], c["n"], c["size_left"])
                        c2.localize(c["host"].numpy(), c["size_left"])
                dms, ems, hyp = [], [], []
                barrier()
                for k in range(k_steps):
                    for c in clouds:
                        flush.fill_(k & 0xFF)
                        barrier()
                        g = c2.localize_device(c["dev"].data_ptr(), c["stride"], c["n"], c["size_left"])
                        dms.append(c2.timings()["total_ms"])
                        hyp.append(sum(c2.gather_result(copy=False)[0]) if world > 1 else len(g))
                for k in range(k_steps):
                    for c in clouds:
                        flush.fill_(k & 0xFF)
                        barrier()
                        t0 = time.perf_counter()
                        g = c2.localize(c["host"].numpy(), c["size_left"])
                        ems.append((time.perf_counter() - t0) * 1e3)
                t2 = torch.tensor([float(np.mean(dms)), float(np.mean(ems))], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
                t2 = t2.cpu().numpy()
                c2.set_stage_timing(True)  # (stage split of this rank: extra calls after the timed ones)
                for _ in range(3):
                    barrier()
                    c2.localize_device(clouds[0]["dev"].data_ptr(), clouds[0]["stride"], clouds[0]["n"], clouds[0]["size_left"])
                tm = c2.timings()
                strong[f"config{cfg}"] = {
                    "workload": ("one fused 7-view cloud (%d pts), 20000 samples" % clouds[0]["n"]) if cfg == 5 else
                                ("%d VGA clouds, 2000 samples each, every cloud sharded over the ranks" % n_clouds),
                    "ms_per_cloud": float(t2[0]), "e2e_ms_per_cloud": float(t2[1]),
                    "hypotheses_per_cloud": float(np.mean(hyp)), "hyp_per_s": float(np.mean(hyp) / (t2[0] * 1e-3)),
                    "rank0_stages_ms": {nm: round(tm[nm], 4) for nm in ("preprocess_ms", "quadric_ms", "sweep_ms",
                                                                         "hog_svm_ms", "total_ms")},
                    "timer": "ms_per_cloud: CUDA events around the whole call incl. the exchange + merge (device-resident "
                             "cloud), e2e: wall clock from the pinned host cloud; max over ranks"}
                c2.set_svm(None)
                c2.close()
                del c2, clouds
            except Exception as e:
                strong[f"config{cfg}"] = {"error": str(e)}
        if rank == 0:
            strong["sharding"] = ("samples interleaved over %d ranks (ag_params.shard_interleave), cloud voxelised on every "
                                  "rank, lists exchanged by peer stores and merged on the device" % world) if world > 1 \
                else "single GPU (the unsharded call)"
            line["strong_scaling"] = strong

    extras = world == 1 and not args.no_extras
    if extras:
        # ---- the other normal mode, same steps (outside the headline's timed region)
        try:
            for c in pool:
                c["P"].deterministic_normals = 1 - det
            ctx.set_params(pool[0]["P"])
            prime()
            k2 = max(4, min(args.steps, 10))
            D2 = timed_device(k2)
            s2, h2, _ = timed_host(k2)
            ctx.set_stage_timing(True)
            prime()
            D2s = timed_device(k2)
            ctx.set_stage_timing(False)
            line["other_mode"] = {"normal_mode": MODE_NAME[1 - det],
                                  "value": float(np.sum(D2["hyps"]) / (np.sum(D2["dev_ms"]) * 1e-3)),
                                  "ms_per_step": float(np.mean(D2["dev_ms"])),
                                  "e2e": {"value": float(np.sum(h2) / np.sum(s2)), "ms_per_cloud": float(1e3 * np.mean(s2))},
                                  "stages_ms": {nm: float(np.mean([t[nm] for t in D2s["t"]])) for nm in
                                                ("preprocess_ms", "quadric_ms", "search_ms", "moments_ms", "axes_ms",
                                                 "sweep_ms", "hog_svm_ms")}, "steps": k2}
        except Exception as e:
            line["other_mode"] = {"error": str(e)}
        finally:
            for c in pool:
                c["P"].deterministic_normals = det
            ctx.set_params(pool[0]["P"])
        # ---- the Taubin stage on a launch that fills the machine: every voxel of the bench cloud as a sample,
        # r = 0.03 — the size of the reference's all-points pass (hand_search.cpp:17-26)
        try:
            c0 = pool[0]
            ctx.set_stage_timing(True)
            step_device(c0)
            n_vox = ctx.timings()["n_voxels"]
            all_idx = np.arange(n_vox, dtype=np.int32)
            c0["P"].deterministic_normals = 1  # (no pick ranking on the side stream while the stage is timed)
            ctx.set_params(c0["P"])
            best = None
            for _ in range(4):
                flush.fill_(1)
                torch.cuda.synchronize()
                ctx.fit_quadrics(all_idx, c0["P"].nn_radius_taubin)
                t = ctx.timings()
                if best is None or t["moments_ms"] + t["search_ms"] < best["moments_ms"] + best["search_ms"]:
                    best = t
            b_alg = 16 * best["taubin_neighbor_points"] + 292 * n_vox
            ms = best["moments_ms"] + best["search_ms"]
            a_sc = b_alg / (ms * 1e-3) / 1e9
            peak, _ = measured_peak()
            line["roofline_at_scale"] = {
                "kernel": "k_ball_search + k_taubin_moments (launches > 16k samples keep the two-kernel variant)",
                "samples": int(n_vox), "algorithmic_bytes_per_launch": float(b_alg),
                "launch_ms": float(ms), "search_ms": float(best["search_ms"]), "moments_ms": float(best["moments_ms"]),
                "achieved": float(a_sc), "peak": peak, "unit": "GB/s", "frac": float(a_sc / peak),
                "moments_only_frac": float(b_alg / (best["moments_ms"] * 1e-3) / 1e9 / peak),
                "axes_ms": float(best["axes_ms"]),
                "note": "all voxels of the bench cloud as samples (L2 flushed before each of 4 launches, best taken)"}
        except Exception as e:  # never lose the bench line over the extra measurement
            line["roofline_at_scale"] = {"error": str(e)}
        finally:
            ctx.set_stage_timing(False)
            c0["P"].deterministic_normals = det
            ctx.set_params(c0["P"])
        # ---- throughput mode (BASELINE config 4 on one GPU): 16 clouds through ag_localize_batch, pinned host
        # buffers in, host grasp lists out (wall clock, best of 3); the headline above stays one cloud per call
        try:
            bc = [pool[i % len(pool)]["host"].numpy() for i in range(16)]
            bs = [pool[i % len(pool)]["size_left"] for i in range(16)]
            ctx.localize_batch(bc[:8], bs[:8])
            ctx.localize_batch(bc[:8], bs[:8])
            best_dt, nh = None, 0
            for _ in range(3):
                t0 = time.perf_counter()
                outs = ctx.localize_batch(bc, bs)
                dt = time.perf_counter() - t0
                if best_dt is None or dt < best_dt:
                    best_dt, nh = dt, sum(len(o) for o in outs)
            line["batched_e2e"] = {"clouds": 16, "lanes": 4, "value": float(nh / best_dt), "unit": "hyp/s",
                                   "ms_per_cloud": float(best_dt * 1e3 / 16),
                                   "note": "ag_localize_batch: 4 clouds in flight on separate streams, each lane "
                                           "replaying its CUDA graph; results identical to sequential calls"}
        except Exception as e:
            line["batched_e2e"] = {"error": str(e)}
        # ---- the launch-file default model (POLY, 588 support vectors): same step with that model attached
        try:
            poly = api.Svm(poly_model_path())
            ctx.set_svm(poly)
            for c in pool:
                for _ in range(3):
                    ctx.localize(c["host"].numpy(), c["size_left"])
            ts, hs, hg = [], [], []
            for k in range(max(4, min(args.steps, 10))):
                c = pool[k % len(pool)]
                flush.fill_(k & 0xFF)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gp = ctx.localize(c["host"].numpy(), c["size_left"])
                gp, keep = ctx.classify(poly, gp)
                ts.append(time.perf_counter() - t0)
                hs.append(len(gp))
            ctx.set_stage_timing(True)
            for k in range(4):
                c = pool[k % len(pool)]
                ctx.localize(c["host"].numpy(), c["size_left"])
                if k >= 2:
                    hg.append(ctx.timings()["hog_svm_ms"])
            ctx.set_stage_timing(False)
            line["poly_svm"] = {"model": "svm_032015_20_20_same (POLY degree 2, 588 support vectors; "
                                         "launch/single_camera_grasps.launch:6)",
                                "e2e": {"value": float(np.sum(hs) / np.sum(ts)), "unit": "hyp/s",
                                        "ms_per_cloud": float(1e3 * np.mean(ts))},
                                "hog_svm_ms": float(np.mean(hg)), "positives_last_step": int(keep.sum())}
        except Exception as e:
            line["poly_svm"] = {"error": str(e)}
        finally:
            ctx.set_svm(svm)
        # ---- CPU baseline: the oracle port on this box's host cores, both normal modes, same scene pool — in a fresh
        # process (= the --impl reference arm): inside this one torch's and numpy's idle worker threads compete with
        # the oracle's OpenMP team and triple its time
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps),
                   "--warmup", "1", "--normals", args.normals, "--config", str(args.config)]
            if "poly_svm" in line and "e2e" in line["poly_svm"]:
                cmd.append("--poly")
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                                 env={k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"})
            ref = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            cb = ref["cpu_baseline"]
            if "poly_svm" in cb:
                line["poly_svm"]["cpu_baseline"] = dict(cb.pop("poly_svm"), cores=cb["cores"])
            line["cpu_baseline"] = cb
        except Exception as e:  # the oracle is test infrastructure; never let it break the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "hyp/s", "cores": 0, "kind": "port",
                                    "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
