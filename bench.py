#!/usr/bin/env python
"""bench.py - NMF multiplicative-update iterations/s on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one NMF iteration (W update, H update, Frobenius error; compute_w = compute_h =
compute_err = True) over the whole synthetic matrix.  Workload (default): BASELINE.json
configs[1], 4096 x 262144 fp32, k = 32 PER GPU; with N GPUs the matrix is N x 262144 columns
(column shards, weak scaling) and `value` counts shard-iterations/s summed over ranks, so at
N = 1 it is plain iterations/s.

Prints ONE JSON line (rank 0).  `value`: X resident in HBM, device-timed with CUDA events, max
over ranks.  `e2e`: the same metric through pymf_b200.NMF with HOST buffers (upload of X/W/H,
K iterations, download of W/H/ferr all inside the timed region).  `roofline`: the streaming
kernels' algorithmic bytes / their CUDA-event time vs MEASURED_PEAKS.json.  `cpu_baseline`: the
numpy oracle port of the reference path timed on this box's host cores on a column prefix.
--impl reference: only the CPU arm (the reference is pure Python/numpy; the oracle port is it).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (d, n per GPU, k, description)
    "cfg1": (1000, 500, 10, "cfg1: 1000x500, k=10"),
    "cfg2": (4096, 262144, 32, "cfg2: 4096x262144 fp32, k=32"),
    "cfg3": (16384, 1048576, 128, "cfg3: 16384x1M fp32, k=128 (STRONG: columns split over GPUs)"),
    "cfg4k64": (8192, 524288, 64, "cfg4: 8192x524288, k=64"),
    "cfg4k16": (8192, 524288, 16, "cfg4: 8192x524288, k=16"),
    "cfg4k32": (8192, 524288, 32, "cfg4: 8192x524288, k=32"),
    "cfg4k128": (8192, 524288, 128, "cfg4: 8192x524288, k=128"),
    "cfg4k256": (8192, 524288, 256, "cfg4: 8192x524288, k=256"),
    "cfg4k512": (8192, 524288, 512, "cfg4: 8192x524288, k=512"),
    "cfg5": (32768, 524288, 64, "cfg5: 32768x524288 per GPU (64 GiB), k=64, weak scaling"),
}
# --mode: which of factorize()'s flag combinations a step is (tests/test_pymf.py:92-95 of the reference)
MODES = {
    "full": dict(compute_w=True, compute_h=True, compute_err=True),        # the BASELINE metric
    "h_only": dict(compute_w=False, compute_h=True, compute_err=False),    # projection on a fixed basis: one pass over X
    "w_only": dict(compute_w=True, compute_h=False, compute_err=True),     # H fixed: A, B are loop invariant, no pass over X
}
METRIC = "NMF MU iterations/sec"
UNIT = "iterations/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured"
    return 6650.0, 1590.0, "fallback"


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        threading.Thread.__init__(self)
        self.daemon = True
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.002)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(2.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU arm
def cpu_port_rate(d, n_full, k, budget_s, steps, warmup):
    """Time the numpy oracle port (same arithmetic as pymf/nmf.py) on a column prefix of the
    workload and extrapolate to n_full (time is linear in n).  Returns (it/s at n_full,
    sample description, cores, measured seconds/iteration on the prefix, n_cpu)."""
    from oracle import nmf_oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    rng = np.random.RandomState(0)

    def one(n_cpu, iters):
        X = rng.random_sample((d, n_cpu))
        W = rng.random_sample((d, k))
        H = rng.random_sample((k, n_cpu))
        O.factorize(X, W, H, niter=1, early_stop=False)          # warm BLAS threads
        t0 = time.perf_counter()
        O.factorize(X, W, H, niter=iters, early_stop=False)
        return (time.perf_counter() - t0) / iters

    n_probe = min(n_full, 2048)
    t_probe = one(n_probe, 2)
    total_iters = max(1, steps + warmup)
    # largest power-of-two prefix such that all iterations fit the budget (and RAM: 5 d*n f64 temps)
    n_cpu = n_probe
    while (n_cpu * 2 <= n_full and t_probe * (n_cpu * 2 / n_probe) * total_iters <= budget_s
           and 5 * 8 * d * n_cpu * 2 <= 8e9):
        n_cpu *= 2
    X = rng.random_sample((d, n_cpu))
    W = rng.random_sample((d, k))
    H = rng.random_sample((k, n_cpu))
    O.factorize(X, W, H, niter=max(1, warmup), early_stop=False)
    t0 = time.perf_counter()
    O.factorize(X, W, H, niter=steps, early_stop=False)
    sec = (time.perf_counter() - t0) / steps
    rate_full = 1.0 / (sec * (float(n_full) / n_cpu))
    sample = ("numpy float64 port of pymf/nmf.py update_w/update_h/frobenius_norm, %d iterations on a "
              "%dx%d column prefix (k=%d), %.3f s/iter measured, extrapolated linearly in n to %d columns"
              % (steps, d, n_cpu, k, sec, n_full))
    return rate_full, sample, cores, sec, n_cpu


def run_reference(args, d, n_per_gpu, k, wl_desc):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    rate, sample, cores, sec, n_cpu = cpu_port_rate(d, n_per_gpu, k, budget_s=150.0, steps=steps,
                                                    warmup=max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1000.0 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_desc, "d": d, "n_per_gpu": n_per_gpu, "k": k,
                   "note": "CPU arm: one host does every shard, so shard-iterations/s does not grow with N"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def run_ours(args, d, n_per_gpu, k, wl_name, wl_desc):
    import torch
    import torch.distributed as dist
    import pymf_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("--gpus %d needs torchrun (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    strong = wl_name == "cfg3"
    if strong:
        n_global = n_per_gpu
        bounds = [(n_global * r) // world for r in range(world + 1)]
        n_loc, col0 = bounds[rank + 1] - bounds[rank], bounds[rank]
    else:
        n_global, n_loc, col0 = n_per_gpu * world, n_per_gpu, n_per_gpu * rank

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = pymf_b200.Engine(d, n_loc, k, device=local_rank, n_global=n_global, col0=col0,
                           path=args.path)
    if world > 1:
        box = [pymf_b200.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], world, rank)
    eng.gen_x(1234)
    eng.gen_w(1235)
    eng.gen_h(1236)
    eng.sync()

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    mode_kw = MODES[args.mode]
    eng.enqueue(warmup, **mode_kw)
    eng.sync()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # per-kernel CUDA events only on streaming-sized problems: small ones are launch-bound and replay the
    # iteration body as a CUDA graph, which per-kernel events would break up
    small = float(d) * n_loc <= 2 ** 24
    eng.kernel_timing(not small)
    l0, g0 = eng.launch_count, eng.graph_replays
    e0, e1 = eng.event(), eng.event()
    barrier()
    eng.record(e0)
    eng.enqueue(steps, **mode_kw)
    eng.record(e1)
    eng.sync()
    barrier()
    ms = max_over_ranks(eng.elapsed_ms(e0, e1))
    launches = (eng.launch_count - l0) * world
    graph_replays = eng.graph_replays - g0
    clocks = sampler.finish() if sampler else None
    t_h, n_h = eng.kernel_timing_read(0)
    t_x, n_x = eng.kernel_timing_read(1)
    eng.kernel_timing(False)
    ferr_check = eng.frobenius()
    path = eng.active_path

    ms_per_step = ms / steps
    units_per_step = 1.0 if strong else float(world)          # shard-iterations per step
    value = units_per_step * 1000.0 / ms_per_step

    # ---- roofline of the streaming pass: bytes_alg = 4 d n_loc + 8 k n_loc per iteration
    hbm_peak, bf16_peak, peak_kind = measured_peaks()
    bytes_alg = 4.0 * d * n_loc + 8.0 * k * n_loc
    if args.mode == "w_only":
        bytes_alg = 0.0                            # no streaming pass at all
    flops_alg = 4.0 * d * n_loc * k + 4.0 * n_loc * k * k + 4.0 * d * k * k
    if args.mode == "h_only":
        flops_alg = 2.0 * d * n_loc * k + 2.0 * n_loc * k * k      # W^T X and G H only
    t_stream_ms = t_h + t_x
    if small:                                     # no per-kernel events: the whole iteration is the unit
        t_stream_ms = ms_per_step
    achieved = bytes_alg / (t_stream_ms * 1e-3) / 1e9 if t_stream_ms > 0 else 0.0
    traffic = None
    try:                                          # DRAM bytes per iteration of the two streaming kernels (ncu capture)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl_name)
        if tj and world == 1 and path == "tc":
            traffic = tj["h_update_bytes"] + tj["xht_bytes"]
    except Exception:
        traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
        "kernel": "streaming pass = H-update kernel + X.H^T kernel (each reads X once per iteration)",
        "alg_bytes_per_iteration": bytes_alg, "h_update_ms": t_h, "xht_ms": t_x,
        "kernel_launches_timed": int(n_h + n_x),
        # each streaming kernel on its own: its bytes (X once + H read/write resp. X once + [H_hi;H_lo]) / its time
        "h_update_frac_of_peak": (bytes_alg / (t_h * 1e-3) / 1e9 / hbm_peak) if t_h > 0 else None,
        "xht_frac_of_peak": (bytes_alg / (t_x * 1e-3) / 1e9 / hbm_peak) if t_x > 0 else None,
        "fp32_equiv_tflops": flops_alg / (t_stream_ms * 1e-3) / 1e12 if t_stream_ms > 0 else 0.0,
        "tensor_frac_3xtf32_of_half_bf16_peak": (3.0 * flops_alg / (t_stream_ms * 1e-3) / 1e12) / (0.5 * bf16_peak)
        if t_stream_ms > 0 else 0.0,
    }
    eng.close()

    # ---- e2e: pymf_b200.NMF with host buffers (upload + K iterations + download in the timed region)
    e2e = None
    if not args.no_e2e and args.mode == "full":
        rng = np.random.default_rng(1234 + rank)
        if args.e2e_source == "pinned":          # page-locked host buffers (the contract's e2e source)
            Xh = pymf_b200.pinned_empty((d, n_loc), np.float32)
            W0 = pymf_b200.pinned_empty((d, k), np.float64)
            H0 = pymf_b200.pinned_empty((k, n_loc), np.float64)
            rng.random(out=Xh, dtype=np.float32)
            rng.random(out=W0)
            rng.random(out=H0)
        else:                                    # ordinary pageable numpy arrays
            Xh = rng.random((d, n_loc), dtype=np.float32)
            W0 = rng.random((d, k))
            H0 = rng.random((k, n_loc))
        # Two passes of the identical user-level sequence; the second is the one reported.  The first pays this
        # process's one-off costs for a matrix of this size (first 4 GiB device allocation and DMA mapping: measured
        # 0.18-0.29 s vs 0.116 s, tests/_e2e_probe2.py) and is reported beside it as `first_call_seconds`.
        rec = []
        for _pass in range(2):
            m = pymf_b200.NMF(Xh, num_bases=k, device=local_rank, path=args.path,
                              process_group=(True if world > 1 else None))
            barrier()
            t0 = time.perf_counter()
            m.W, m.H = W0, H0
            m._sync_to_device()        # X / W / H host -> device (part of factorize; split out for the breakdown)
            t1 = time.perf_counter()
            m.factorize(niter=steps)
            t2 = time.perf_counter()
            _ = (m.W, m.H, m.ferr)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            rec.append((max_over_ranks(t3 - t0), t1 - t0, t2 - t1, t3 - t2, bool(m._engine.last_upload_pinned)))
            del m
            if _pass == 0:             # fresh starting factors for the reported pass (W0 / H0 were updated in place)
                rng = np.random.default_rng(4321 + rank)
                W0[...] = rng.random(W0.shape)
                H0[...] = rng.random(H0.shape)
        t_e2e, t_up, t_it, t_dn, direct = rec[1]
        h2d = (Xh.nbytes + W0.nbytes + H0.nbytes) / float(steps)
        d2h = (W0.nbytes + H0.nbytes + 8 * steps) / float(steps)
        e2e = {"value": units_per_step * steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "seconds_total": t_e2e,
               "seconds_upload": t_up, "seconds_iterations": t_it, "seconds_download": t_dn,
               "first_call_seconds": rec[0][0],
               "host_source": args.e2e_source, "x_upload_direct_dma": direct,
               "numa": dict(zip(("gpu_node", "x_host_node"), pymf_b200.numa_info(Xh, local_rank))),
               "what": "NMF(X_host).factorize(niter=%d) incl. X/W/H upload and W/H/ferr download; second of two "
                       "identical calls in this process" % steps}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, sample, cores, sec, n_cpu = cpu_port_rate(d, n_per_gpu, k, budget_s=20.0, steps=3, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl_desc + (" per GPU, n_global = %d" % n_global if not strong else ""),
                       "mode": args.mode, "d": d, "n_local": n_loc, "n_global": n_global, "k": k, "kernel_path": path,
                       "arithmetic": "fp32 storage; 3xTF32 tcgen05 products (tc path) or fp32 FMA (simt path); fp64 error combine",
                       "l2": ("inputs larger than L2 (X shard = %.2f GiB), no flush needed" % (4.0 * d * n_loc / 2 ** 30))
                       if 4.0 * d * n_loc > 2 * 126e6 else
                       ("X shard = %.1f MB fits the 126 MB L2 and is NOT flushed between iterations: the loop re-reads "
                        "the same matrix every step by construction (latency-bound workload, no HBM roofline claim)"
                        % (4.0 * d * n_loc / 1e6)),
                       "value_units": "shard-iterations/s summed over ranks" if not strong else "iterations/s of the global problem",
                       "ferr_after": ferr_check},
            "e2e": e2e, "gpu_launches": int(launches), "graph_replays": int(graph_replays), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default=None, choices=[None, "auto", "simt", "tc"])
    ap.add_argument("--mode", default="full", choices=sorted(MODES),
                    help="full = W, H and error every step (the BASELINE metric); h_only / w_only = serving modes")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-source", default="pinned", choices=["pinned", "pageable"],
                    help="host memory the e2e leg reads X/W/H from (default: page-locked)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    d, n, k, desc = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, d, n, k, desc)
    else:
        run_ours(args, d, n, k, args.workload, desc)


if __name__ == "__main__":
    main()
