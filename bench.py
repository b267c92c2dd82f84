#!/usr/bin/env python
"""Benchmark of the scoring-and-selection hot path (BASELINE.json metric: pool frames scored+selected / sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

Workload (BASELINE.json configs[1], "C2"): Panoptic 19 joints, 8 views, 100k-frame pool per GPU: heat-map decode
(arg-max) + multi-view RANSAC/DLT triangulation + reprojection-uncertainty score for every frame, then the top-k
ranking (strategy.py:945-949).  One step = one pass over the whole pool.  A 100k x 8 x 19 float32 heat-map pool is
249 GB, more than one GPU's HBM, so the pool is streamed as chunk passes over a resident buffer of
`resident_frames` distinct frames (far larger than the 126 MB L2, so no pass is served from cache).
For N > 1 every rank owns its own 100k-frame shard (weak scaling, contiguous frame sharding, no data-path
collective; one small all_gather for the ranking merge).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pool frames scored+selected/sec"
UNIT = "frames/s"
V, J, H, W, STRIDE = 8, 19, 64, 64, 4
POOL_FRAMES_PER_GPU = 100_000
TOPK = 1000
FRAME_HEATMAP_BYTES = V * J * H * W * 4
FRAME_ALGO_BYTES = FRAME_HEATMAP_BYTES + V * 96 + J + J * 24 + 16


def workload_config(args, n_gpus):
    return {
        "workload": ("C2: Panoptic 19-joint, 8 views, 100k-frame pool per GPU: heatmap decode + RANSAC/DLT "
                     "triangulation + reprojection uncertainty + top-k ranking") if (V, J) == (8, 19) else
                    ("%d-joint, %d views, %d-frame pool per GPU: heatmap decode + RANSAC/DLT triangulation + reprojection "
                     "uncertainty + top-k ranking" % (J, V, args.pool_frames)),
        "views": V, "joints": J, "heatmap": [H, W], "pool_frames_per_gpu": args.pool_frames,
        "pool_frames_total": args.pool_frames * n_gpus, "resident_frames": args.resident_frames, "topk": TOPK,
        "n_iters": 64, "epsilon_px": 5.0,
        "cache": "inputs (resident buffer %.1f GB) far larger than the 126 MB L2; no flush needed"
                 % (args.resident_frames * FRAME_HEATMAP_BYTES / 1e9),
        "sharding": "contiguous frames per rank, no data-path collective; ranking = 1 all_gather of k (idx, score) + device merge",
        "launches": "one persistent fused launch per step over the whole shard (chunk passes over the resident buffer as segments)",
    }


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region.  In-process NVML polling from a thread (a
    spawned `nvidia-smi -lms` takes the driver lock on every query and was measured to stall kernel launches of
    a 36 ms step by tens of ms); falls back to one nvidia-smi query if NVML is unavailable."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=0.05):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _run(self):
        nv, h = self._nvml, self._handle
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            self._handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM))
            self._nvml = nv
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml"}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.max_mhz, samples=len(self.sm),
                       reasons=sorted(self.reasons))
            return out
        try:  # fallback: a single nvidia-smi query right after the timed region
            q = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(x) for x in q.strip().split(",")[:2]]
            out.update(sm_mhz=a, sm_max_mhz=b, samples=1, source="nvidia-smi (after the timed region)")
        except Exception:
            pass
        return out


# ----------------------------------------------------------------------------------------------------------------
# the reference arm: the reference's CPU algorithm (oracle port; the reference itself is Python and cannot travel
# to the GPU box) on all host cores, on a bounded sample of the same workload
# ----------------------------------------------------------------------------------------------------------------
def _ref_worker_init(seed, frames):
    global _W
    from multi_view_active_learning_b200 import synthetic as S

    pool = S.make_pool(frames, V, J, seed=seed, p_outlier=0.1)
    _W = (S.render_heatmaps(pool["centres"], noise=0.05, seed=seed + 1), pool["P"], pool["valid"])


def _ref_worker_step(_):
    from oracle import triangulation_oracle as O

    hm, P, valid = _W
    out = O.triangulate_pool(hm, P, STRIDE, valid)
    return out["metric"]


def cpu_port_single_core(hm, P, valid, topk):
    """The oracle port on ONE core over the given host arrays; returns (frames/s, seconds)."""
    from oracle import scores_oracle as SO
    from oracle import triangulation_oracle as O

    t0 = time.perf_counter()
    out = O.triangulate_pool(hm, P, STRIDE, valid)
    SO.rank_nlargest({i: float(m) for i, m in enumerate(out["metric"])}, topk)
    dt = time.perf_counter() - t0
    return hm.shape[0] / dt, dt, out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    from oracle import scores_oracle as SO

    cores = os.cpu_count() or 1
    per_core = args.ref_frames_per_core
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, initializer=_ref_worker_init, initargs=(1000 + c, per_core)) for c in range(cores)]
    frames = per_core * cores

    def step():
        res = [p.map_async(_ref_worker_step, [0]) for p in pools]
        metrics = np.concatenate([r.get()[0] for r in res])
        SO.rank_nlargest({i: float(m) for i, m in enumerate(metrics)}, TOPK)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for p in pools:
        p.terminate()
    value = frames * args.steps / dt
    sample = "%d frames per step (%d per core x %d processes), oracle port of utils/triangulation.py + nlargest" % (
        frames, per_core, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (not present on the GPU box); this arm is the vectorised numpy oracle "
                "port, which is ~60x faster per core than the reference's own per-frame loop (SURVEY.md probe: "
                "239 ms/frame at 8 views)",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# our arm
def gpu_numa_info(index):
    """NUMA placement of GPU `index` and of this process (context for the host-buffer e2e number)."""
    info = {}
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:  # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bdf = bdf[4:]
        base = "/sys/bus/pci/devices/" + bdf
        info["gpu_numa_node"] = int(open(base + "/numa_node").read())
        info["gpu_local_cpulist"] = open(base + "/local_cpulist").read().strip()
    except Exception as e:  # informational only
        info["error"] = repr(e)[:120]
    try:
        info["process_cpu_count"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return info


def parse_cpulist(text):
    cpus = set()
    for part in (text or "").split(","):
        if part.strip():
            a, _, b = part.strip().partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from multi_view_active_learning_b200 import _lib, ops, pool as poolmod
    from multi_view_active_learning_b200 import synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- mval_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    R = args.resident_frames
    pool_frames = args.pool_frames
    shard_start = rank * pool_frames  # weak scaling: every rank owns pool_frames frames of the global pool

    # ---- synthetic resident buffer (generated on the device; seeds recorded in the JSON line)
    seed = 1234 + rank
    host_pool = S.make_pool(R, V, J, seed=seed, p_outlier=0.1)
    centres = torch.from_numpy(host_pool["centres"]).to(dev)
    P = torch.from_numpy(host_pool["P"]).to(dev)
    hm = torch.empty((R, V, J, H, W), dtype=torch.float32, device=dev)
    ops.synth_heatmaps(centres, H, W, 1.0, 0.05, seed, out=hm)
    torch.cuda.synchronize()

    # the shard as chunk passes over the resident buffer, scored by ONE persistent launch (mval_score_pool_segments): frame f
    # of the shard reads heat maps of resident frame f % R, with its own projection matrices / outputs
    seg_list, done = [], 0
    while done < pool_frames:
        n = min(R, pool_frames - done)
        seg_list.append(hm[:n])
        done += n
    assert len(seg_list) <= _lib.MAX_SEGMENTS, "raise --resident-frames: more than %d chunk passes per shard" % _lib.MAX_SEGMENTS
    P_shard = torch.cat([P[: t.shape[0]] for t in seg_list])
    exchange = poolmod.RankingExchange(TOPK, dev)
    sel_host = torch.empty(TOPK, dtype=torch.int64).pin_memory()
    work = {"out": None}
    k_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def step(ev=None):
        if ev is not None:
            ev[0].record()
        work["out"] = ops.score_pool_segments(seg_list, P_shard, STRIDE, None, pair_seed=0, frame_offset=shard_start,
                                              out=work["out"])
        if ev is not None:
            ev[1].record()
        idx, val, cnt = exchange(work["out"]["metric"], shard_start)  # local top-k, one all_gather, device-side merge
        sel_host.copy_(idx, non_blocking=True)  # the selection lands in pinned host memory, no synchronisation here
        return idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ops.check_async()
    # the clock sampler (NVML initialisation, tens of ms) starts BEFORE the barrier: started after it, on rank 0 only, it
    # delayed rank 0's first step and every other rank waited for it in the ranking exchange inside its own timed region
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0.record()
    for k in range(args.steps):
        step(k_events[k])
        marks[k].record()
    ev1.record()
    barrier()
    ops.check_async()  # a tripped mbarrier watchdog would have left incomplete outputs: fail loudly instead
    launches = _lib.launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    step_ms = [round(a.elapsed_time(b), 3) for a, b in zip([ev0] + marks[:-1], marks)]
    fused_ms = float(np.mean([a.elapsed_time(b) for a, b in k_events]))  # the dominant kernel, live, inside the timed region
    sel = [sel_host.tolist()]
    clocks = sampler.stop() if rank == 0 else None
    per_rank_fused_ms = [fused_ms]
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        g = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g, torch.tensor([fused_ms], dtype=torch.float64, device=dev))
        per_rank_fused_ms = g.cpu().tolist()
    ms_per_step = elapsed_ms / args.steps
    value = pool_frames * n_gpus / (ms_per_step * 1e-3)

    # ---- roofline of the individual kernels on the resident buffer (CUDA events on the launching stream)
    def time_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    xy = ops.decode_argmax(hm, STRIDE)
    dec_ms = time_ms(lambda: ops.decode_argmax(hm, STRIDE), 5)
    tri_ms = time_ms(lambda: ops.triangulate_ransac(xy, P), 5)
    del xy
    dec_gbs = R * FRAME_HEATMAP_BYTES / (dec_ms * 1e-3) / 1e9
    pool_bytes = pool_frames * FRAME_ALGO_BYTES
    pool_gbs = pool_bytes / (fused_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the tracked ncu capture (tools/ncu_traffic.py writes profiles/traffic.json from
    # the `ncu --set full` report: dram__bytes_read.sum + dram__bytes_write.sum per launch and the frames of that launch)
    traffic, traffic_src = None, "no ncu capture for this shape in profiles/traffic.json"
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("score_pool_fused_kernel_v%d_j%d" % (V, J))
        if rec:
            traffic = float(rec["dram_bytes"]) / float(rec["frames"]) * pool_frames
            traffic_src = "%s: %.0f B over %d frames, scaled per frame to this launch" % (rec["source"], rec["dram_bytes"], rec["frames"])
    except Exception:
        pass
    # dominant kernel: ONE persistent launch of the fused kernel scores the whole shard; its duration is measured live inside
    # the timed region with CUDA events on the launching stream
    roofline = {
        "bound": "hbm", "kernel": "score_pool_fused_kernel (mval_score_pool_segments: one launch per step over the shard)",
        "achieved": pool_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": pool_gbs / hbm_peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": pool_bytes,
        "algorithmic_bytes_per_frame": FRAME_ALGO_BYTES, "frames_per_launch": pool_frames, "avg_launch_ms": fused_ms,
        "launch_timing": "CUDA events around the launch in every timed step (mean of %d)" % args.steps,
        "share_of_step": fused_ms / ms_per_step,
        # every rank scores the same number of frames; under the power cap the GPUs of a box settle at different clocks and
        # the step (max over ranks, coupled by the ranking exchange) follows the slowest one
        "per_rank_avg_launch_ms": [round(x, 3) for x in per_rank_fused_ms],
        "share_of_step_slowest_rank": max(per_rank_fused_ms) / ms_per_step,
        "other_kernels": {
            "decode_argmax_kernel": {"achieved": dec_gbs, "frac": dec_gbs / hbm_peak, "avg_launch_ms": dec_ms,
                                     "algorithmic_bytes_per_launch": R * FRAME_HEATMAP_BYTES},
            "ransac_vote+final+frame_reduce (stand-alone, from key-points)": {"avg_ms_per_resident_chunk": tri_ms,
                                                                             "bound": "fp64 pipe"},
        },
    }

    # ---- end-to-end through the host-buffer entry (every rank, concurrently): the H2D copy of the step's inputs
    # from pinned host memory, the kernels, the D2H copy of the results and the ranking exchange are all inside the
    # timed region; max over ranks per step, median over steps.  Same sample size at every N.
    E = min(args.e2e_frames, R)
    pin = lambda t: t.cpu().pin_memory()
    # pinned pages are placed by first touch: allocate them from the CPUs next to this rank's GPU when the process is
    # allowed to run there, so that the H2D stream does not cross the socket interconnect
    numa = gpu_numa_info(local_rank)
    allowed = os.sched_getaffinity(0)
    near = allowed & parse_cpulist(numa.get("gpu_local_cpulist"))
    if near:
        os.sched_setaffinity(0, near)
    h_hm, h_P = pin(hm[:E]), pin(P[:E])
    if near:
        os.sched_setaffinity(0, allowed)
    numa["pinned_from_gpu_local_cpus"] = bool(near)
    # the raw link: one cudaMemcpyAsync of the same pinned heat maps, CUDA events (what e2e can reach at most), first on
    # this rank ALONE (ranks take turns), then on all ranks at once -- the second is what a shared PCIe uplink leaves
    probe_dst = torch.empty_like(hm[: min(E, 1024)])
    probe_src = h_hm[: probe_dst.shape[0]]
    probe_dst.copy_(probe_src, non_blocking=True)
    link_alone = None
    for r in range(world):
        barrier()
        if r == rank:
            link_alone = probe_src.numel() * 4 / (time_ms(lambda: probe_dst.copy_(probe_src, non_blocking=True), 3) * 1e-3) / 1e9
    barrier()
    link_gbs = probe_src.numel() * 4 / (time_ms(lambda: probe_dst.copy_(probe_src, non_blocking=True), 3) * 1e-3) / 1e9
    del probe_dst
    pipe = ops.HostPipeline(V, J, H, W, chunk_frames=args.e2e_chunk, n_slots=3, device=dev)
    e2e_exchange = poolmod.RankingExchange(TOPK, dev)
    outs = None
    e2e_times, own_times = [], []
    for it in range(2 + args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        outs = pipe.score_pool(h_hm, h_P, STRIDE, None, frame_offset=shard_start, out=outs)
        idx, _, _ = e2e_exchange(outs["metric"].to(dev, non_blocking=True), shard_start)
        sel_e2e = idx.cpu()  # the step ends with the selected indices on the host
        dt = time.perf_counter() - t0
        own = dt
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if it > 1:
            e2e_times.append(dt)
            own_times.append(own)
    pipe.close()
    e2e_val = E * n_gpus / float(np.median(e2e_times))
    h2d = E * (FRAME_HEATMAP_BYTES + V * 96)
    d2h = sum(t.numel() * t.element_size() for t in outs.values()) + sel_e2e.numel() * 8
    per_rank = [h2d / float(np.median(own_times)) / 1e9, link_alone, link_gbs]
    if world > 1:
        g = torch.empty(world * 3, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g, torch.tensor(per_rank, dtype=torch.float64, device=dev))
        per_rank = g.view(world, 3).cpu().tolist()
    else:
        per_rank = [per_rank]
    topo = None
    if rank == 0 and world > 1:
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:6000]
        except Exception:
            pass
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d * n_gpus, "d2h_bytes_per_step": d2h * n_gpus,
           "frames_per_step": E * n_gpus, "steps": args.e2e_steps, "aggregate": "max over ranks per step, median over steps",
           "step_ms": [round(1e3 * t, 2) for t in e2e_times],
           "h2d_gbs": h2d / float(np.median(e2e_times)) / 1e9, "link_h2d_gbs_measured": link_gbs,
           "per_rank": {"columns": ["e2e_h2d_gbs (own step time)", "link_h2d_gbs, this rank copying alone",
                                    "link_h2d_gbs, all ranks copying at once"], "rows": per_rank},
           "link_note": "link_* = one cudaMemcpyAsync of the same pinned heat maps (CUDA events): the bound of any end-to-end "
                        "number whose inputs start in host memory; when the all-ranks figure is below the alone figure the "
                        "ranks share a PCIe uplink / host memory controller and e2e cannot scale past it", "numa": numa,
           "nvidia_smi_topo": topo,
           "call": "ops.HostPipeline.score_pool = mval_pipeline_score_pool (pinned host heat maps -> chunked H2D over 3 "
                   "persistent staging slots / streams -> fused kernel -> D2H; nothing allocated per call) + mval_topk_desc + "
                   "all_gather + mval_topk_merge, on every rank concurrently"}
    del h_hm
    # Variant at the reference's own boundary: triangulation() receives the heat maps as CUDA tensors straight from
    # the backbone (strategy.py:1027-1045) and only the projection matrices / validity masks live on the host.  Per step:
    # H2D of P (+ valid) from pinned memory, fused kernel over the resident heat maps, ranking, D2H of every result.
    h_P_all = P.cpu().pin_memory()
    dv_times = []
    res_host = None
    for it in range(1 + args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        dP = h_P_all.to(dev, non_blocking=True)
        out_dv = ops.score_pool(hm, dP, STRIDE, None, frame_offset=shard_start, return_keypoints_2d=True)
        idx_dv, _, _ = e2e_exchange(out_dv["metric"], shard_start)
        res_host = {k: v.to("cpu", non_blocking=True) for k, v in out_dv.items()}
        sel_dv = (idx_dv.cpu(),)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if it > 0:
            dv_times.append(dt)
    e2e["device_heatmaps_variant"] = {
        "value": R * n_gpus / float(np.median(dv_times)), "unit": UNIT, "frames_per_step": R * n_gpus,
        "h2d_bytes_per_step": R * V * 96 * n_gpus,
        "d2h_bytes_per_step": (sum(t.numel() * t.element_size() for t in res_host.values()) + len(sel_dv[0]) * 16) * n_gpus,
        "step_ms": [round(1e3 * t, 2) for t in dv_times],
        "call": "ops.score_pool on CUDA heat maps (the reference's triangulation() boundary: heat maps come from the "
                "backbone on the device, strategy.py:1027-1045) with P copied from pinned host memory and all results "
                "copied back, + mval_topk_desc + ranking merge"}

    # ---- the other BASELINE.json configurations, measured in the same run so that they reach the driver's records
    extra = None
    if not args.no_extra:
        extra = extra_records({"world": world, "rank": rank, "dev": dev}, args, hm, P, hbm_peak, peak_src, time_ms)

    line = None
    if rank == 0:
        # ---- CPU baseline (oracle port, one core) on a bounded sample of the same frames, N = 1 only
        cpu = None
        if n_gpus == 1 and args.cpu_frames > 0:
            S_ = min(args.cpu_frames, R)
            c_hm = hm[:S_].cpu().numpy()
            fps, dt, ref = cpu_port_single_core(c_hm, host_pool["P"][:S_], host_pool["valid"][:S_], TOPK)
            got = ops.score_pool(hm[:S_], P[:S_], STRIDE, None, frame_offset=0)
            parity = bool(np.array_equal(got["inlier_count"].cpu().numpy(), ref["inlier_count"])
                          and np.allclose(got["metric"].cpu().numpy(), ref["metric"], rtol=0, atol=1e-4)
                          and np.array_equal(got["keypoints_2d"].cpu().numpy(), ref["keypoints_2d"]))
            cpu = {"value": fps, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "%d frames of the same pool, vectorised numpy oracle of utils/triangulation.py + "
                             "nlargest, %.1f s; host has %d cores" % (S_, dt, os.cpu_count() or 0),
                   "gpu_matches_oracle_on_sample": parity}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, n_gpus), "seed": 1234,
            "step_ms": step_ms, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "selected_head": [int(i) for i in sel[0][:5]], "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def rig_record(dev, views, joints, frames, hbm_peak, peak_src, time_ms, label):
    """The fused scoring kernel on another rig (BASELINE.json configs[2] / [4]): one launch over `frames` resident frames."""
    import torch

    from multi_view_active_learning_b200 import ops
    from multi_view_active_learning_b200 import synthetic as S

    pool = S.make_pool(min(frames, 256), views, joints, seed=4321, p_outlier=0.1, box=100.0 if joints == 42 else 400.0,
                       radius=500.0 if joints == 42 else 3000.0)
    reps = -(-frames // pool["centres"].shape[0])
    centres = torch.from_numpy(np.tile(pool["centres"], (reps, 1, 1, 1))[:frames]).to(dev)
    Pr = torch.from_numpy(np.tile(pool["P"], (reps, 1, 1, 1))[:frames]).to(dev)
    hm_r = ops.synth_heatmaps(centres, H, W, 1.0, 0.05, 99)  # distinct noise per frame: nothing repeats in the buffer
    out = {"out": None}

    def run():
        out["out"] = ops.score_pool_segments([hm_r], Pr, STRIDE, None, pair_seed=1234, out=out["out"])

    run()
    ms = time_ms(run, 3)
    ops.check_async()
    frame_bytes = views * joints * H * W * 4 + views * 96 + joints + joints * 24 + 16
    gbs = frames * frame_bytes / (ms * 1e-3) / 1e9
    rec = {"workload": label, "views": views, "joints": joints, "frames_per_launch": frames, "avg_launch_ms": ms,
           "value": frames / (ms * 1e-3), "unit": UNIT,
           "roofline": {"bound": "hbm", "kernel": "score_pool_fused_kernel", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_frame": frame_bytes, "algorithmic_bytes_per_launch": frames * frame_bytes},
           "view_pairs": "64 of %d per joint (counter-based subsets, utils/triangulation.py:279-282)" % (views * (views - 1) // 2),
           "metric_mean": float(out["out"]["metric"].mean().item())}
    del hm_r, out
    torch.cuda.empty_cache()
    return rec


def extra_records(ctx, args, hm, P_res, hbm_peak, peak_src, time_ms):
    """Sub-records of the default line for the configurations BASELINE.json names besides C2: the north-star target through
    the API (T), the coreset (C4, strong scaling over the ranks of this run), and the fused kernel on the C3 / C5 rigs."""
    import torch

    world, rank = ctx["world"], ctx["rank"]
    extra = {}
    api = api_selection(ctx, hm, P_res, args.extra_api_frames, 8192, 10_000, 1000, 1000, 5)
    if rank == 0:
        extra["T_api_sample_next_batch"] = {
            "workload": "north-star target through the API: %d frames per GPU x %d GPU(s) = %d frames, 8 views, 19 joints; "
                        "ActiveLearningStrategy.sample_next_batch with device heat maps" % (args.extra_api_frames, world,
                                                                                              args.extra_api_frames * world),
            "variants": api}
    c4 = coreset_record(ctx, args.coreset_rows, args.coreset_dim, 1000, 10_000, with_cpu=False)
    if rank == 0:
        extra["C4_coreset"] = c4
    if rank == 0:  # kernel-level records of the other rigs: one GPU's worth, measured on rank 0
        # whole frames per CTA, one persistent CTA per SM: a launch of k x (SM count) frames has no ragged last wave (a real
        # shard -- 62 500 frames per GPU for C3 -- has hundreds of frames per SM and does not care; 1 536 frames were 10.4 per SM,
        # i.e. 6 % of the launch spent in an eleventh wave of 56 CTAs)
        sms = torch.cuda.get_device_properties(ctx["dev"]).multi_processor_count
        extra["C3_interhand_20v_42j"] = rig_record(ctx["dev"], 20, 42, 11 * sms, hbm_peak, peak_src, time_ms,
                                                   "C3 rig: InterHand 42 joints, 20 views (fused decode + RANSAC + uncertainty)")
        extra["C5_panoptic_31v_19j"] = rig_record(ctx["dev"], 31, 19, 14 * sms, hbm_peak, peak_src, time_ms,
                                                  "C5 rig: Panoptic 19 joints, 31 views (fused decode + RANSAC + uncertainty)")
    if world > 1:
        torch.distributed.barrier()
    return extra if rank == 0 else None


# ----------------------------------------------------------------------------------------------------------------
# secondary workload: coreset k-center greedy (BASELINE.json configs[3], "C4"): not the driver's default line
# ----------------------------------------------------------------------------------------------------------------
def _timed(fn, iters=1):
    import torch

    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


def coreset_record(ctx, n_total, d, L, budget, path="auto", data="gaussian", kslots=0, cpu_rows=50000, with_cpu=True, pad=0):
    """C4-style workload: k-center greedy over n_total x d float32 features sharded by rows over the ranks (STRONG scaling:
    the total is fixed).  Returns (on rank 0) the record of the whole selection (norms + labeled fold-in + budget picks in
    exact rounds, device-timed, max over ranks) with two rooflines: the tensor-pipe fraction of the tcgen05 screening GEMM
    that dominates the batched update, and the HBM fraction of the single-centre step the reference's loop is made of."""
    import torch
    import torch.distributed as dist

    from multi_view_active_learning_b200 import _lib, ops, pool as poolmod

    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    flags = {"auto": 0, "ffma": 1, "tc": 2}[path]
    lo, hi = poolmod.shard_range(n_total, world, rank)
    n = hi - lo
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    feat = torch.randn((n, d), generator=g, device=dev, dtype=torch.float32)
    gl = torch.Generator(device=dev).manual_seed(7)
    labeled = torch.randn((L, d), generator=gl, device=dev, dtype=torch.float32)
    if pad:  # zero columns up to a multiple of `pad`: bit-identical distances (pool.pad_features), tensor-core screen applicable
        feat, labeled = poolmod.pad_features(feat, pad), poolmod.pad_features(labeled, pad)
        d = feat.shape[1]
    if data == "clustered":  # 64 tight clusters: a pick collapses the minima of its whole cluster
        gc = torch.Generator(device=dev).manual_seed(5)
        cent = torch.randn((64, d), generator=gc, device=dev, dtype=torch.float32) * 4.0
        feat = feat * 0.25 + cent[torch.randint(0, 64, (n,), generator=g, device=dev)]
        labeled = labeled * 0.25 + cent[torch.randint(0, 64, (L,), generator=gl, device=dev)]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # tcgen05 kind::tf32 runs at half the dense bf16 rate; the measured bf16 figure (a cuBLAS GEMM) halved is the peak
    tf32_peak = float(peaks.get("bf16_tflops", 2250.0)) / 2.0
    tf32_sustained = float(peaks["bf16_tflops_sustained"]) / 2.0 if "bf16_tflops_sustained" in peaks else None
    tf32_src = ("MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 issues at half the bf16 rate)" if "bf16_tflops" in peaks
                else "fallback: nominal 2250 dense bf16 TFLOP/s / 2")

    # ---- kernels timed alone on this rank's shard (CUDA events on the launching stream, after a warm-up call)
    norms = ops.kcenter_norms(feat)
    lab_norms = ops.kcenter_norms(labeled)
    norms_ms, _ = _timed(lambda: ops.kcenter_norms(feat), 3)
    min_d = torch.full((n,), float("inf"), dtype=torch.float32, device=dev)
    ops.kcenter_update(feat, norms, labeled[0], min_d)
    single_ms, _ = _timed(lambda: ops.kcenter_update_batch(feat, norms, labeled[:1], lab_norms[:1], min_d, 1), 5)
    T = min(256, n)
    cidx = torch.randint(0, n, (T,), generator=g, device=dev)
    cent_rows, cent_norms = feat[cidx].contiguous(), norms[cidx].contiguous()
    poolmod.kcenter_fold_centres([{"feat": feat, "norms": norms, "min": min_d, "off": lo}], labeled, lab_norms, flags=1)
    scratch = min_d.clone()

    def update(fl):
        scratch.copy_(min_d)
        ops.kcenter_update_batch(feat, norms, cent_rows, cent_norms, scratch, fl)

    update(1)
    ffma_ms, _ = _timed(lambda: update(1), 2)
    auto_ms, tc_survivors, tc_capacity = None, None, None
    if flags != 1:
        update(flags)
        auto_ms, _ = _timed(lambda: update(flags), 3)
        tc_survivors, tc_capacity = ops.kcenter_tc_stats()
    clone_ms, _ = _timed(lambda: scratch.copy_(min_d), 3)
    rec = ops.kcenter_select(feat, norms, min_d, lo, 256)
    select_ms, _ = _timed(lambda: ops.kcenter_select(feat, norms, min_d, lo, 256, out=rec), 5)
    del scratch

    # ---- the whole selection, device-timed, max over ranks (second of two runs: the first warms the allocator / NCCL)
    total_ms, stats, sel, launches = None, [], None, 0
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        stats = []
        total_ms, (sel, fin) = _timed(lambda: poolmod.kcenter_greedy_sharded([(feat, lo)], labeled, budget, flags=flags, stats=stats,
                                                                                 k_slots=kslots or None))
        launches = _lib.launch_count() - l0
    # the same 256-centre update against the SETTLED minima the rounds work on (after the labeled fold alone the running
    # minima are still high and the exact recheck of the screen's survivors is several times longer than in any round)
    steady_ms, steady_survivors = None, None
    if flags != 1 and auto_ms is not None:
        min_fin = fin[0]
        scratch = min_fin.clone()
        own = sel[(sel >= lo) & (sel < hi)] - lo  # centres = the last picks that live in this shard: far points, as in a round
        if own.numel() >= T:
            cent_rows, cent_norms = feat[own[-T:]].contiguous(), norms[own[-T:]].contiguous()

        def update_steady():
            scratch.copy_(min_fin)
            ops.kcenter_update_batch(feat, norms, cent_rows, cent_norms, scratch, flags)

        for _ in range(8):  # ~20 ms of the same load first: three launches after the host-side pause above caught the clocks low
            update_steady()
        steady_ms, _ = _timed(update_steady, 8)
        steady_survivors, _ = ops.kcenter_tc_stats()
        del scratch, min_fin
    del fin
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- CPU arm on a bounded sample (rank 0): the C oracle (same arithmetic) on all host cores
    cpu = None
    if rank == 0 and with_cpu:
        from oracle import coreset_oracle as CO

        ns = min(n, cpu_rows)
        Fs = np.concatenate([feat[:ns].cpu().numpy(), labeled[: min(L, 8)].cpu().numpy()])
        steps = 8
        t0 = time.perf_counter()
        cpu_sel, _ = CO.kcenter_greedy_f32(Fs, ns, steps)
        cpu_s = time.perf_counter() - t0
        passes = min(L, 8) + steps + 1  # centre passes + the norms pass
        cpu_rows_per_s = ns * passes / cpu_s  # row-passes per second
        full_passes = L + budget + 1
        cpu = {"value": n_total / (full_passes * n_total / cpu_rows_per_s), "unit": "rows/s", "cores": os.cpu_count(),
               "kind": "port", "sample": "C oracle (OpenMP, fmaf) on %d x %d rows: %d centre passes in %.2f s; extrapolated "
               "linearly to %d passes over %d rows (the sequential algorithm has no batching)" % (ns, d, passes, cpu_s,
                                                                                                   full_passes, n_total)}
        # parity of the sample against the GPU on the same rows
        gsel, _ = ops.kcenter_greedy(torch.from_numpy(Fs).to(dev), ns, steps)
        cpu["gpu_matches_oracle_on_sample"] = bool(gsel.cpu().tolist() == cpu_sel)
    out = None
    if rank == 0:
        fold_ms = (auto_ms if auto_ms is not None else ffma_ms) - clone_ms
        upd_ms = (steady_ms - clone_ms) if steady_ms is not None else fold_ms
        if steady_survivors is not None:
            tc_survivors = steady_survivors
        ffma_only = ffma_ms - clone_ms
        flops = 2.0 * n * T * d
        step_bytes = n * (d * 4 + 12)
        tc_ran = auto_ms is not None and tc_survivors is not None and tc_capacity
        out = {
            "metric": "coreset k-center greedy: pool rows selected-from / sec", "value": n_total / (total_ms * 1e-3),
            "unit": "rows/s", "n_gpus": world, "ms_total": total_ms, "higher_is_better": True, "dtype": "f32",
            "scaling": "strong", "data": "synthetic (%s)" % data,
            "config": {"workload": "C4-style coreset: %d x %d float32 features, %d labeled centres, budget %d, rows sharded "
                                   "over %d GPU(s); exact greedy in rounds" % (n_total, d, L, budget, world),
                       "update_path": path},
            "rounds": {"count": len(stats), "picks_per_round_mean": float(np.mean(stats)), "picks_per_round_min": int(min(stats)),
                       "picks_per_round_max": int(max(stats))},
            "ms_per_pick": total_ms / budget,
            "equivalent_sequential_ms": (L + budget) * single_ms,
            "kernels_ms": {"norms": norms_ms, "single_centre_update": single_ms, "update_256_centres_ffma": ffma_only,
                           "update_256_centres_selected_path": upd_ms, "update_256_centres_selected_path_minima_after_labeled_fold_only": fold_ms,
                           "select_256": select_ms},
            "roofline": {"bound": "tensor", "kernel": "kc_screen_tc2_kernel (+ kc_recheck_kernel + the no-op FFMA guard): one "
                         "batched update of %d centres (the selection's last picks in this shard) over this rank's %d rows, running "
                         "minima as the greedy rounds see them (the selection's final state)" % (T, n) if tc_ran else
                         "kc_batch_kernel (FFMA; the tensor-core screen does not apply to this shape)",
                         "achieved": flops / (upd_ms * 1e-3) / 1e12, "peak": tf32_peak if tc_ran else 72.0, "unit": "TFLOP/s",
                         "frac": flops / (upd_ms * 1e-3) / 1e12 / (tf32_peak if tc_ran else 72.0), "traffic": None,
                         "peak_source": tf32_src if tc_ran else "fp32 FFMA pipe: 148 SMs x 128 lanes x 2 x 1.9 GHz",
                         "algorithmic_flops_per_launch": flops, "avg_launch_ms": upd_ms,
                         "frac_of_sustained_peak": (flops / (upd_ms * 1e-3) / 1e12 / tf32_sustained) if tc_ran and tf32_sustained else None,
                         "sustained_peak": tf32_sustained if tc_ran else None,
                         "sustained_peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (under the power cap the SM clock "
                                                  "settles near 1.1-1.3 GHz while the tensor pipe is busy; profiles/r2r_ncu_screen_c4_"
                                                  "key_metrics.json: pipe active 91 % of the cycles at 1.13 GHz)" if tc_ran else None,
                         "tc_survivor_pairs": tc_survivors, "tc_pair_capacity": tc_capacity,
                         "ffma_tflops_same_update": flops / (ffma_only * 1e-3) / 1e12},
            "roofline_single_step": {"bound": "hbm", "kernel": "kc_rowdot_kernel<1> (single-centre update, one greedy step of the reference)",
                                     "achieved": step_bytes / (single_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": step_bytes / (single_ms * 1e-3) / 1e9 / hbm_peak,
                                     "algorithmic_bytes_per_launch": step_bytes, "avg_launch_ms": single_ms},
            "gpu_launches": int(launches), "selected_head": sel[:5].cpu().tolist()}
        if cpu is not None:
            out["cpu_baseline"] = cpu
    del feat, norms, min_d
    torch.cuda.empty_cache()
    return out


def run_coreset(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = coreset_record({"world": world, "rank": rank, "dev": dev}, args.coreset_rows, args.coreset_dim, args.coreset_labeled,
                         args.coreset_budget, args.coreset_path, args.coreset_data, args.coreset_kslots, args.coreset_cpu_rows,
                         with_cpu=args.cpu_frames > 0, pad=args.coreset_pad)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------------------------
# secondary workload: hybrid selection (BASELINE.json configs[4] / north-star target): uncertainty scoring of the whole
# pool + top-k ranking + coreset k-center over the predicted poses, end to end
# ----------------------------------------------------------------------------------------------------------------
def run_hybrid(args):
    """Every rank owns `pool_frames` frames (north-star target: 8 x 125k = 1M frames, 8 views, 19 joints).  One step =
    fused decode + RANSAC triangulation + uncertainty over the whole shard (resident chunk passes, as in the scoring
    workload) -> ranking (top-k by uncertainty, one all_gather) -> root-relative pose features (coreset.py:35-47,
    d = 3J = 57) -> k-center greedy (budget picks, `coreset_labeled` labeled poses) in exact rounds (one all_gather per
    round).  Distinct poses per pool frame: the resident heat maps repeat, the projection matrices do not (each pool
    frame's cameras are composed with its own similarity transform of the world, which leaves every reprojection
    error unchanged and rotates / scales the triangulated pose)."""
    import torch
    import torch.distributed as dist

    from multi_view_active_learning_b200 import _lib, ops, pool as poolmod
    from multi_view_active_learning_b200 import synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R, n, L, budget = args.resident_frames, args.pool_frames, args.coreset_labeled, args.coreset_budget
    shard_start = rank * n
    seed = 1234 + rank
    host_pool = S.make_pool(R, V, J, seed=seed, p_outlier=0.1)
    hm = ops.synth_heatmaps(torch.from_numpy(host_pool["centres"]).to(dev), H, W, 1.0, 0.05, seed)
    P_res = torch.from_numpy(host_pool["P"]).to(dev)
    # per-pool-frame similarity transform M_f = [s R | t]: P_f = P_res[f % R] @ M_f
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    q = torch.randn((n, 4), generator=g, device=dev, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w_, x_, y_, z_ = q.unbind(1)
    Rm = torch.stack([1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_),
                      2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_),
                      2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)], dim=1).reshape(n, 3, 3)
    sc = 0.5 + torch.rand((n, 1, 1), generator=g, device=dev, dtype=torch.float64)
    M = torch.zeros((n, 4, 4), dtype=torch.float64, device=dev)
    M[:, :3, :3] = Rm * sc
    M[:, :3, 3] = torch.randn((n, 3), generator=g, device=dev, dtype=torch.float64) * 100
    M[:, 3, 3] = 1.0
    idx = torch.arange(n, device=dev) % R
    P_pool = torch.matmul(P_res[idx], M[:, None])  # [n, V, 3, 4]
    gl = torch.Generator(device=dev).manual_seed(7)
    labeled = (torch.randn((L, J, 3), generator=gl, device=dev) * 300.0)
    root = 2
    labeled = (labeled - labeled[:, root:root + 1]).permute(0, 2, 1).reshape(L, 3 * J).contiguous()
    chunks = [(o, min(R, n - o)) for o in range(0, n, R)]

    def step(stats=None):
        metrics, feats = [], []
        for off, m in chunks:
            out = ops.score_pool(hm[:m], P_pool[off:off + m], STRIDE, None, pair_seed=0, frame_offset=shard_start + off,
                                 return_keypoints_2d=False)
            metrics.append(out["metric"])
            feats.append(ops.pose_features(out["keypoints_3d"], root))  # utils/coreset.py:35-47 (mval_pose_features)
        metric = torch.cat(metrics)
        ranked = poolmod.distributed_topk(ops.topk_desc(metric, TOPK, index_offset=shard_start), TOPK)
        feat = torch.cat(feats)
        sel, _ = poolmod.kcenter_greedy_sharded([(feat, shard_start)], labeled, budget, stats=stats,
                                                pad_to=args.coreset_pad or None, k_slots=args.coreset_kslots or None)
        return ranked, sel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = _lib.launch_count()
    times, stats = [], []
    for _ in range(args.steps):
        barrier()
        stats = []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ranked, sel = step(stats)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    launches = (_lib.launch_count() - l0) // max(args.steps, 1)
    verified = None
    if args.verify:
        # the sharded selection (one all_gather per round) against the single-device loop over the gathered features
        feats = []
        for off, m in chunks:
            o = ops.score_pool(hm[:m], P_pool[off:off + m], STRIDE, None, pair_seed=0, frame_offset=shard_start + off,
                               return_keypoints_2d=False)
            feats.append(ops.pose_features(o["keypoints_3d"], root))
        feat = torch.cat(feats)
        if world > 1:
            parts = [torch.empty_like(feat) for _ in range(world)]
            dist.all_gather(parts, feat)
            feat = torch.cat(parts)
        if rank == 0:
            ref_sel, _ = ops.kcenter_greedy(torch.cat([feat, labeled]), feat.shape[0], budget)
            verified = bool(torch.equal(ref_sel.cpu(), sel.cpu()))
        del feat
    score_ms, _ = _timed(lambda: [ops.score_pool(hm[:m], P_pool[o:o + m], STRIDE, None, frame_offset=shard_start + o,
                                                 return_keypoints_2d=False) for o, m in chunks])
    if rank == 0:
        ms = float(np.median(times))
        print(json.dumps({
            "metric": "hybrid selection: pool frames scored + ranked + coreset-selected / sec", "value": n * world / (ms * 1e-3),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "step_ms": times,
            "higher_is_better": True, "scaling": "weak", "dtype": "f64 (triangulation) / f32 (coreset)", "data": "synthetic",
            "config": {"workload": "north-star target / C5-style hybrid: %d frames per GPU x %d GPU(s), %d views, %d joints: "
                                   "fused decode + RANSAC triangulation + uncertainty, top-%d ranking, coreset k-center over "
                                   "root-relative predicted poses (d = %d), %d labeled, budget %d"
                                   % (n, world, V, J, TOPK, 3 * J, L, budget), "resident_frames": R},
            "breakdown_ms": {"scoring_kernels": score_ms, "ranking_features_coreset": ms - score_ms},
            "coreset_rounds": {"count": len(stats), "picks_per_round_mean": float(np.mean(stats)) if stats else None},
            "selection_matches_single_device": verified,
            "gpu_launches": int(launches), "selected_head": sel[:5].cpu().tolist(), "ranked_head": [int(i) for i in ranked[0][:5]]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------------------------
# the north-star target THROUGH THE REPOSITORY'S OWN API: ActiveLearningStrategy.sample_next_batch (reference
# strategy.py:54-135 -> _sal_pseudo_labeling :915-1002 -> _compute_sal_dict :1004-1147 -> nlargest / CoreSet
# utils/coreset.py:13-95) over a pool whose heat maps are on the device (what the pose estimator's forward leaves there)
# ----------------------------------------------------------------------------------------------------------------
def similarity_proj(P_res, n, R, seed, dev):
    """Distinct cameras for every pool frame although the resident heat maps repeat every R frames: frame f uses
    P_res[f % R] @ M_f with M_f a similarity transform of the world, which leaves every reprojection error unchanged
    and rotates / scales / shifts the triangulated pose (so the coreset sees n different poses)."""
    import torch

    g = torch.Generator(device=dev).manual_seed(seed)
    q = torch.randn((n, 4), generator=g, device=dev, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w_, x_, y_, z_ = q.unbind(1)
    Rm = torch.stack([1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_),
                      2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_),
                      2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)], dim=1).reshape(n, 3, 3)
    sc = 0.5 + torch.rand((n, 1, 1), generator=g, device=dev, dtype=torch.float64)
    M = torch.zeros((n, 4, 4), dtype=torch.float64, device=dev)
    M[:, :3, :3] = Rm * sc
    M[:, :3, 3] = torch.randn((n, 3), generator=g, device=dev, dtype=torch.float64) * 100
    M[:, 3, 3] = 1.0
    idx = torch.arange(n, device=dev) % R
    return torch.matmul(P_res[idx], M[:, None])  # [n, V, 3, 4]


class DevicePoolDataset:
    """The surface of dataset/dataset.py's ActiveLearningDataset that the selection path touches (:47-74, 98-110), over a
    synthetic pool whose heat maps are resident on the device.  Rank r owns the frames r, r + G, r + 2G, ... (the
    DistributedSampler's deal without its shuffle), so the gathered table is in global frame order."""

    POSE_ID = 160422

    def __init__(self, hm, P_pool, n_local, rank, world, batch, n_labeled, joints, dev):
        import torch

        self.hm, self.P, self.n_local, self.rank, self.world, self.batch, self.dev = hm, P_pool, n_local, rank, world, batch, dev
        self.R = hm.shape[0]
        gl = torch.Generator().manual_seed(7)
        lab = torch.randn((n_labeled, 4, joints), generator=gl, dtype=torch.float64) * 300.0
        self.labeled_data = [{"3d_keypoints": lab[i].numpy()} for i in range(n_labeled)]
        self.pseudo_label_guids, self.pseudo_labeled_data, self.labeled_guids = [], [], []
        self.valid = torch.ones((batch, joints), dtype=torch.float32, device=dev)
        self.gt = torch.zeros((batch, 4, joints), dtype=torch.float32, device=dev)
        self.frame_ids = torch.arange(n_local, device=dev, dtype=torch.int64) * world + rank
        self.pose_ids = torch.full((n_local,), self.POSE_ID, device=dev, dtype=torch.int64)

    # dataset/dataset.py:98-102, 47-51, 61-74
    def resample_unlabeled_data(self):
        pass

    def get_al_dict_for_coreset(self):
        return {i: np.array(self.labeled_data[i]["3d_keypoints"]).transpose([1, 0]) for i in range(len(self.labeled_data))}

    def label_by_frame_guids(self, guids):
        self.labeled_guids = list(guids)

    def pseudo_label_by_frame_guids(self, guids, pseudo_labels):
        self.pseudo_label_guids = guids
        self.pseudo_labeled_data = [{"pseudo_3d_keypoints": np.array(pseudo_labels[g]).transpose([1, 0])} for g in guids]

    def loader(self):
        """What DataLoader(dataset, batch_size, sampler=DistributedSampler(dataset)) yields (dataset/dataset.py:142-155)
        when the backbone's output is already on the device: "images" carries the heat maps and the pose estimator is the
        identity (the forward is timed separately, north star)."""
        for o in range(0, self.n_local, self.batch):
            m = min(self.batch, self.n_local - o)
            r0 = o % self.R
            if r0 + m > self.R:
                r0 = 0
            yield {"images": self.hm[r0:r0 + m], "proj_matrices": self.P[o:o + m], "joint_valid": self.valid[:m],
                   "3d_keypoints": self.gt[:m], "pose": self.pose_ids[o:o + m], "frame_id": self.frame_ids[o:o + m]}


def api_cfg(strategy, expr, batch, joints):
    from types import SimpleNamespace as NS

    return NS(EXPR_TYPE=expr, RANDOM_SEED=1307, DATA=NS(NUM_JOINTS=joints, TYPE="panoptic"), POSE_ESTIMATOR=NS(STRIDE=STRIDE),
              SAL=NS(INLIER_THRESHOLD=4, CLUSTER_FILE_PATH="", NUM_CLUSTERS=10),
              AL=NS(STRATEGY=strategy, USE_SOFTARGMAX=False, USE_REPROJECTION_XE=False, REPROJECTION_SIGMA=1.0, HP_CONFIG="AVG",
                    MPE_CONFIG="AVG", BSB_CONFIG="AVG", INFERENCE=NS(BATCH_SIZE=batch, NUM_WORKERS=0)))


def api_selection(ctx, hm, P_res, n_local, batch, budget, n_labeled, pseudo, reps, variants=None):
    """Times ActiveLearningStrategy.sample_next_batch(iteration >= 1) end to end (wall clock between a barrier +
    synchronize on both sides, max over ranks, median over `reps` calls after one warm-up call) for the given
    (strategy, EXPR_TYPE) variants.  Returns {variant: record}."""
    import gc
    import random

    import torch
    import torch.distributed as dist

    from multi_view_active_learning_b200 import _lib
    from multi_view_active_learning_b200.strategy import ActiveLearningStrategy

    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    P_pool = similarity_proj(P_res, n_local, hm.shape[0], 77 + rank, dev)
    out = {}
    for strategy, expr in (variants or (("TRIANGULATION", "AL"), ("CORESET", "AL"), ("TRIANGULATION", "SAL"))):
        cfg = api_cfg(strategy, expr, batch, J)
        st = ActiveLearningStrategy(cfg)
        times, sel, launches = [], None, 0
        for it in range(1 + reps):
            ds = DevicePoolDataset(hm, P_pool, n_local, rank, world, batch, n_labeled, J, dev)
            st._get_dataloader = lambda d, bs, nw, ds=ds: ds.loader()
            random.seed(5)
            gc.collect()  # the previous call's dataset / table: a generation-2 pass inside a 60 ms call was a 30 ms outlier
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            t0 = time.perf_counter()
            st.sample_next_batch(ds, budget, pseudo, torch.nn.Identity(), iteration=1, rank=rank)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            launches = _lib.launch_count() - l0
            if world > 1:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            if it > 0:
                times.append(dt)
            sel = (st.last_al_guids, st.last_sal_guids)
        same = None
        if world > 1:  # every rank must end with the same selection (strategy.py:945-950)
            import hashlib

            h = int(hashlib.sha1(json.dumps(sel).encode()).hexdigest()[:15], 16)
            hs = torch.tensor([h], dtype=torch.int64, device=dev)
            all_h = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_h, hs)
            same = bool((all_h == all_h[0]).all().item())
        ms = 1e3 * float(np.median(times))
        out["%s/%s" % (strategy, expr)] = {
            "ms_per_call": ms, "call_ms": [round(1e3 * t, 2) for t in times], "value": n_local * world / (ms * 1e-3), "unit": UNIT,
            "frames_total": n_local * world, "al_num_frames": budget, "sal_num_frames": pseudo if expr == "SAL" else 0,
            "labeled": n_labeled, "al_guids_head": sel[0][:3], "n_al_guids": len(sel[0]), "n_sal_guids": len(sel[1]),
            "same_selection_on_every_rank": same, "gpu_launches": int(launches),
            "call": "ActiveLearningStrategy.sample_next_batch(train_dataset, al_num_frames, sal_num_frames, pose_estimator, "
                    "iteration=1, rank) -- strategy.py:54; heat maps on the device (identity pose estimator, batches of %d "
                    "frames), device-resident sal_dict (table.py)" % batch}
    return out


def run_api(args):
    """--workload api: the north-star target through sample_next_batch: `pool_frames` frames per GPU (default for this
    workload: 125 000 = 1M over 8 GPUs), 8 views, 19 joints."""
    import torch
    import torch.distributed as dist

    from multi_view_active_learning_b200 import ops
    from multi_view_active_learning_b200 import synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R = args.resident_frames
    host_pool = S.make_pool(R, V, J, seed=1234 + rank, p_outlier=0.1)
    hm = ops.synth_heatmaps(torch.from_numpy(host_pool["centres"]).to(dev), H, W, 1.0, 0.05, 1234 + rank)
    P_res = torch.from_numpy(host_pool["P"]).to(dev)
    ctx = {"world": world, "rank": rank, "dev": dev}
    n_local = args.pool_frames
    variants = [tuple(v.split("/")) for v in args.api_variants.split(",")] if args.api_variants else None
    rec = api_selection(ctx, hm, P_res, n_local, args.api_batch, args.coreset_budget, args.coreset_labeled, args.api_pseudo,
                        args.steps, variants)
    if rank == 0:
        head = rec.get("CORESET/AL") or next(iter(rec.values()))
        print(json.dumps({
            "metric": "selection through sample_next_batch: pool frames scored + selected / sec", "value": head["value"],
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "ms_per_step": head["ms_per_call"], "higher_is_better": True,
            "scaling": "weak", "dtype": "f64 (triangulation) / f32 (coreset)", "data": "synthetic",
            "config": {"workload": "north-star target through the API: %d frames per GPU x %d GPU(s), %d views, %d joints, "
                                   "sample_next_batch (strategy.py:54) with device heat maps" % (n_local, world, V, J),
                       "resident_frames": R, "batch": args.api_batch}, "variants": rec}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------------------------
# secondary workload: the per-map heat-map kernels (soft-arg-max, HP, MPE, BSB, XE) on a resident chunk
# ----------------------------------------------------------------------------------------------------------------
def run_scores(args):
    """Roofline of every per-map kernel of SURVEY.md rows a1/a3/a8/a9: one launch over `resident_frames` frames of
    V x J heat maps (16 KiB each, read once), CUDA events, algorithmic bytes = maps x H*W*4."""
    import torch

    from multi_view_active_learning_b200 import ops
    from multi_view_active_learning_b200 import synthetic as S

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    n = args.resident_frames
    pool = S.make_pool(min(n, 2048), V, J, seed=1234)
    reps = (n + pool["centres"].shape[0] - 1) // pool["centres"].shape[0]
    centres = torch.from_numpy(np.tile(pool["centres"], (reps, 1, 1, 1))[:n]).to(dev)
    P = torch.from_numpy(np.tile(pool["P"], (reps, 1, 1, 1))[:n]).to(dev)
    hm = ops.synth_heatmaps(centres, noise=0.05, seed=1)
    xyz = torch.from_numpy(np.tile(pool["X"], (reps, 1, 1))[:n]).to(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    algo = n * FRAME_HEATMAP_BYTES
    kernels = {
        "decode_argmax_kernel (a1)": lambda: ops.decode_argmax(hm, STRIDE),
        "map_stream_kernel<SoftArgmaxOp> (a3)": lambda: ops.decode_softargmax(hm, STRIDE),
        "map_stream_kernel<HpOp> (a8 HP)": lambda: ops.score_hp(hm),
        "map_stream_kernel<PeaksOp<0>> MPE (a8)": lambda: ops.score_peaks(hm, "MPE"),
        "map_stream_kernel<PeaksOp<1>> BSB (a8)": lambda: ops.score_peaks(hm, "BSB"),
        "map_stream_kernel<XeOp> + xe_frame_reduce (a9)": lambda: ops.score_xe(hm, P, xyz, 2.0),
        "score_pool_fused_kernel (a1+a4..a7)": lambda: ops.score_pool(hm, P, STRIDE, return_keypoints_2d=False),
    }
    # the fused pass with the AL score of the same maps evaluated by its decode warps (one read of the pool instead of
    # two: compare with the sum of the fused line and the map_stream line of that score)
    for kind in ("HP", "MPE", "BSB"):
        kernels["mval_score_pool_scored<%s> (a1+a4..a8; default path)" % kind] = (
            lambda kind=kind: ops.score_pool(hm, P, STRIDE, return_keypoints_2d=False, map_score=kind))
    # arg-max flavour A/B (csrc/fused.cu: launch_score_pool_fused): the non-default flavour of each variant
    def flavoured(value, kind):
        def run():
            os.environ["MVAL_ROW_ARGMAX"] = value
            try:
                return ops.score_pool(hm, P, STRIDE, return_keypoints_2d=False, map_score=kind)
            finally:
                del os.environ["MVAL_ROW_ARGMAX"]
        return run

    def unsplit(kind, value):
        def run():
            os.environ["MVAL_SCORED_SPLIT"] = value
            try:
                return ops.score_pool(hm, P, STRIDE, return_keypoints_2d=False, map_score=kind)
            finally:
                del os.environ["MVAL_SCORED_SPLIT"]
        return run

    # the default is the faster of the two per score (MPE, BSB: stream kernel + RANSAC launches); both forced:
    kernels["score_pool_fused_kernel<MPE> in ONE launch (MVAL_SCORED_SPLIT=0)"] = unsplit("MPE", "0")
    kernels["score_pool_fused_kernel<BSB> in ONE launch (MVAL_SCORED_SPLIT=0)"] = unsplit("BSB", "0")
    kernels["map_stream<argmax+MPE> + RANSAC launches (MVAL_SCORED_SPLIT=1)"] = unsplit("MPE", "1")
    kernels["map_stream<argmax+BSB> + RANSAC launches (MVAL_SCORED_SPLIT=1)"] = unsplit("BSB", "1")
    kernels["score_pool_fused_kernel, lane=row arg-max (MVAL_ROW_ARGMAX=1)"] = flavoured("1", None)
    kernels["score_pool_fused_kernel<MPE>, generic arg-max scan (MVAL_ROW_ARGMAX=0)"] = flavoured("0", "MPE")
    kernels["score_pool_fused_kernel<BSB>, generic arg-max scan (MVAL_ROW_ARGMAX=0)"] = flavoured("0", "BSB")
    def with_env(kind, env):
        def run():
            os.environ.update(env)
            try:
                return ops.score_pool(hm, P, STRIDE, return_keypoints_2d=False, map_score=kind)
            finally:
                for k in env:
                    del os.environ[k]
        return run

    # warp split of the scored fused launch: (12 - s) decode + (3 + s) RANSAC warps
    for kind in ("HP", "MPE", "BSB"):
        for shape in ("0", "1"):
            kernels["score_pool_fused_kernel<%s> in ONE launch, %d decode + %d RANSAC warps (MVAL_FUSED_SHAPE=%s)"
                    % (kind, 12 - int(shape), 3 + int(shape), shape)] = with_env(kind, {"MVAL_SCORED_SPLIT": "0", "MVAL_FUSED_SHAPE": shape})
    if args.scores_only:
        kernels = {k: f for k, f in kernels.items() if args.scores_only in k}
    out = {}
    for _ in range(40):  # ~0.25 s of load before the first timed kernel: the first entry of a process otherwise ran 2x slow
        ops.decode_argmax(hm, STRIDE)
    for name, fn in kernels.items():
        for _ in range(3):
            fn()
        ms, _ = _timed(fn, 5)
        out[name] = {"avg_launch_ms": ms, "achieved": algo / (ms * 1e-3) / 1e9, "frac": algo / (ms * 1e-3) / 1e9 / hbm_peak}
    print(json.dumps({"metric": "per-map heat-map kernels: algorithmic GB/s", "unit": "GB/s", "n_gpus": 1,
                      "config": {"workload": "%d frames x %d views x %d joints x %dx%d float32 heat maps resident (%.1f GB), "
                                 "one launch each" % (n, V, J, H, W, algo / 1e9)},
                      "roofline": {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "algorithmic_bytes_per_launch": algo,
                                   "peak_source": peak_src, "kernels": out}}))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# context only: the pose-estimator forward that produces the heat maps (stays in PyTorch / cuDNN, north star: "timed
# separately").  Random-init PoseResNet-50 ("Simple Baselines": ResNet-50 trunk, three 256-channel stride-2 deconvolutions,
# 1x1 head -> J x 64 x 64 for a 256 x 256 crop; the architecture pose_estimators/pose_resnet.py builds).
# ----------------------------------------------------------------------------------------------------------------
def run_backbone(args):
    import torch
    import torch.nn as nn

    def bottleneck(cin, planes, stride):
        class B(nn.Module):
            def __init__(self):
                super().__init__()
                self.body = nn.Sequential(
                    nn.Conv2d(cin, planes, 1, bias=False), nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                    nn.Conv2d(planes, planes, 3, stride, 1, bias=False), nn.BatchNorm2d(planes), nn.ReLU(inplace=True),
                    nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4))
                self.down = None
                if stride != 1 or cin != planes * 4:
                    self.down = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))

            def forward(self, x):
                return torch.relu(self.body(x) + (x if self.down is None else self.down(x)))
        return B()

    layers = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1)]
    cin = 64
    for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            layers.append(bottleneck(cin, planes, stride if b == 0 else 1))
            cin = planes * 4
    for _ in range(3):
        layers += [nn.ConvTranspose2d(cin, 256, 4, 2, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True)]
        cin = 256
    layers.append(nn.Conv2d(256, J, 1))
    torch.cuda.set_device(0)
    model = nn.Sequential(*layers).cuda().eval().to(memory_format=torch.channels_last)
    out = {}
    frames = args.backbone_frames
    for name, dtype in (("fp32 (TF32 convolutions off)", torch.float32), ("bf16 autocast", torch.bfloat16)):
        torch.backends.cudnn.allow_tf32 = False
        x = torch.randn((frames * V, 3, 256, 256), device="cuda").to(memory_format=torch.channels_last)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
            for _ in range(3):
                y = model(x)
            ms, y = _timed(lambda: model(x), 5)
        assert tuple(y.shape) == (frames * V, J, H, W)
        out[name] = {"ms_per_batch": ms, "frames_per_s": frames / (ms * 1e-3), "images_per_s": frames * V / (ms * 1e-3)}
    print(json.dumps({"metric": "pose-estimator forward (context only, not part of the scoring metric)", "unit": "frames/s",
                      "n_gpus": 1, "config": {"workload": "PoseResNet-50, random init, %d frames x %d views of 3 x 256 x 256 -> %d x "
                                              "64 x 64 heat maps, torch/cuDNN, channels_last" % (frames, V, J)}, "results": out}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool-frames", type=int, default=None, help="frames per GPU (default 100000; api / hybrid: 125000)")
    ap.add_argument("--resident-frames", type=int, default=16384)
    ap.add_argument("--e2e-frames", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=9)
    ap.add_argument("--e2e-chunk", type=int, default=0, help="frames per staging chunk of the host pipeline (0 = ~256 MiB)")
    ap.add_argument("--cpu-frames", type=int, default=4096)
    ap.add_argument("--ref-frames-per-core", type=int, default=128)
    ap.add_argument("--workload", default="scoring", choices=["scoring", "coreset", "scores", "hybrid", "backbone", "api"])
    ap.add_argument("--api-batch", type=int, default=8192, help="api workload: frames per loader batch")
    ap.add_argument("--api-pseudo", type=int, default=1000, help="api workload: sal_num_frames of the SAL variant")
    ap.add_argument("--api-variants", default="", help="api workload: comma list of STRATEGY/EXPR_TYPE (default: TRIANGULATION/AL,"
                    "CORESET/AL,TRIANGULATION/SAL), e.g. HP/AL,MPE/AL,BSB/AL")
    ap.add_argument("--coreset-rows", type=int, default=1_000_000)
    ap.add_argument("--coreset-dim", type=int, default=2048)
    ap.add_argument("--coreset-labeled", type=int, default=None, help="labeled centres (coreset: 64; api / hybrid: 1000)")
    ap.add_argument("--coreset-budget", type=int, default=None, help="picks (coreset: 2048; api / hybrid: 10000)")
    ap.add_argument("--coreset-path", default="auto", choices=["auto", "ffma", "tc"])
    ap.add_argument("--coreset-data", default="gaussian", choices=["gaussian", "clustered"])
    ap.add_argument("--coreset-cpu-rows", type=int, default=50000)
    ap.add_argument("--scores-only", default="", help="scores workload: only the kernels whose name contains this string")
    ap.add_argument("--backbone-frames", type=int, default=16, help="backbone workload: frames (x views images) per batch")
    ap.add_argument("--verify", action="store_true", help="hybrid: gather the pose features and check the sharded coreset "
                    "selection against the single-device loop on rank 0")
    ap.add_argument("--coreset-kslots", type=int, default=0, help="candidate slots per shard and round (0 = pool.py default)")
    ap.add_argument("--coreset-pad", type=int, default=0, help="coreset / hybrid: zero-pad the features to a multiple of this")
    ap.add_argument("--no-extra", action="store_true", help="scoring workload: skip the extra sub-records (T / C4 / C3 / C5)")
    ap.add_argument("--extra-api-frames", type=int, default=125_000, help="frames per GPU of the T record in `extra`")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample clocks during the timed region")
    ap.add_argument("--views", type=int, default=V, help="camera views per frame (C2: 8; C3: 20; C5: 31)")
    ap.add_argument("--joints", type=int, default=J, help="joints per frame (Panoptic 19; InterHand 42)")
    args = ap.parse_args()
    big = args.workload in ("api", "hybrid")
    if args.pool_frames is None:
        args.pool_frames = 125_000 if big else POOL_FRAMES_PER_GPU
    if args.coreset_labeled is None:
        args.coreset_labeled = 1000 if big else 64
    if args.coreset_budget is None:
        args.coreset_budget = 10_000 if big else 2048
    if (args.views, args.joints) != (V, J):  # other BASELINE.json shapes (C3 / C5): same code path, different rig
        g = globals()
        g["V"], g["J"] = args.views, args.joints
        g["FRAME_HEATMAP_BYTES"] = args.views * args.joints * H * W * 4
        g["FRAME_ALGO_BYTES"] = g["FRAME_HEATMAP_BYTES"] + args.views * 96 + args.joints + args.joints * 24 + 16
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "coreset":
        return run_coreset(args)
    if args.workload == "scores":
        return run_scores(args)
    if args.workload == "hybrid":
        return run_hybrid(args)
    if args.workload == "backbone":
        return run_backbone(args)
    if args.workload == "api":
        return run_api(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
