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
        "workload": "C2: Panoptic 19-joint, 8 views, 100k-frame pool per GPU: heatmap decode + RANSAC/DLT "
                    "triangulation + reprojection uncertainty + top-k ranking",
        "views": V, "joints": J, "heatmap": [H, W], "pool_frames_per_gpu": args.pool_frames,
        "pool_frames_total": args.pool_frames * n_gpus, "resident_frames": args.resident_frames, "topk": TOPK,
        "n_iters": 64, "epsilon_px": 5.0,
        "cache": "inputs (resident buffer %.1f GB) far larger than the 126 MB L2; no flush needed"
                 % (args.resident_frames * FRAME_HEATMAP_BYTES / 1e9),
        "sharding": "contiguous frames per rank, no data-path collective; ranking = 1 all_gather of k (idx, score)",
    }


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = tempfile.mktemp(prefix="mval_clocks_", suffix=".csv")
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0]))
                    mx.append(float(p[1]))
                except ValueError:
                    continue
                for name, flag in zip(names, p[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
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

    chunks = []
    done = 0
    while done < pool_frames:
        n = min(R, pool_frames - done)
        chunks.append((done, n))
        done += n

    def step():
        metrics = []
        for off, n in chunks:
            out = ops.score_pool(hm[:n], P[:n], STRIDE, None, pair_seed=0, frame_offset=shard_start + off,
                                 return_keypoints_2d=False)
            metrics.append(out["metric"])
        metric = torch.cat(metrics)
        local = ops.topk_desc(metric, TOPK, index_offset=shard_start)
        return poolmod.distributed_topk(local, TOPK)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sel = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        sel = step()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
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
    pool_ms = time_ms(lambda: ops.score_pool(hm, P, STRIDE, return_keypoints_2d=False), 5)
    dec_bytes = R * FRAME_HEATMAP_BYTES
    dec_gbs = dec_bytes / (dec_ms * 1e-3) / 1e9
    pool_gbs = R * FRAME_ALGO_BYTES / (pool_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": "decode_argmax_kernel", "achieved": dec_gbs, "peak": hbm_peak, "unit": "GB/s",
        "frac": dec_gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dec_bytes, "avg_launch_ms": dec_ms,
        "share_of_step": dec_ms / pool_ms,
        "other_kernels": {"ransac_vote+final+frame_reduce_ms_per_resident_chunk": tri_ms,
                          "triangulation_share_of_step": tri_ms / pool_ms},
        "whole_scoring_step": {"achieved": pool_gbs, "frac": pool_gbs / hbm_peak, "ms_per_resident_chunk": pool_ms,
                               "algorithmic_bytes_per_frame": FRAME_ALGO_BYTES},
    }

    line = None
    if rank == 0:
        # ---- end-to-end through the host-buffer entry: H2D of the step's inputs and D2H of its results inside
        E = min(args.e2e_frames, R)
        pin = lambda t: t.cpu().pin_memory()
        h_hm, h_P = pin(hm[:E]), pin(P[:E])
        outs = None
        e2e_times = []
        for it in range(1 + args.e2e_steps):
            t0 = time.perf_counter()
            outs = ops.score_pool_host(h_hm, h_P, STRIDE, None, frame_offset=shard_start, out=outs)
            idx, val = ops.topk_desc(outs["metric"].to(dev, non_blocking=True), TOPK, index_offset=shard_start)
            idx_h = idx.cpu()
            torch.cuda.synchronize()
            if it > 0:
                e2e_times.append(time.perf_counter() - t0)
        e2e_val = E / (sum(e2e_times) / len(e2e_times))
        h2d = E * (FRAME_HEATMAP_BYTES + V * 96)
        d2h = sum(t.numel() * t.element_size() for t in outs.values()) + idx_h.numel() * 8
        e2e = {"value": e2e_val * n_gpus if n_gpus > 1 else e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "frames_per_step": E, "steps": args.e2e_steps,
               "call": "mval_score_pool_host (pinned host heat maps -> chunked H2D on 2 streams -> kernels -> D2H) + "
                       "mval_topk_desc; measured on rank 0" + (", scaled by n_gpus" if n_gpus > 1 else "")}
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
            "dtype": "f64", "data": "synthetic", "config": dict(workload_config(args, n_gpus), seed=1234),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "selected_head": [int(i) for i in sel[0][:5]],
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool-frames", type=int, default=POOL_FRAMES_PER_GPU)
    ap.add_argument("--resident-frames", type=int, default=16384)
    ap.add_argument("--e2e-frames", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=4096)
    ap.add_argument("--ref-frames-per-core", type=int, default=128)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
