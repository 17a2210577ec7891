#!/usr/bin/env python
"""Headline benchmark: observations/s for fused posterior + logpdf (fp64) of the FiniteBLR path.

    python bench.py --gpus N --steps K --warmup W            # our arm (libblr_cuda on B200)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU op sequence on host cores

Workload (BASELINE.json configs[2], the one the metric is quoted on): diagonal-noise BLR, N = 2^24 observations,
D = 1024 features, ColVecs layout, synthetic data generated in place on the device (Philox), prior mw = 0, Λw = I.
A "step" is one full posterior+logpdf inference over all N observations.  N is fixed as GPUs are added
(observations are sharded; one NCCL sum-allreduce of the packed statistics per step) => strong scaling.
The 128 GiB design matrix is far larger than the 126 MB L2, so no explicit L2 flush is needed between steps.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "obs/s for posterior+logpdf (fp64, N=16M, D=1024)"
UNIT = "obs/s"
N_FULL, D_FULL = 1 << 24, 1024
NOMINAL_FP64_TFLOPS = 37.2  # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz


def gram_flops(N, D):  # algorithmic flops of the Gram kernel: SYRK lower + r
    return float(N) * D * (D + 1) + 2.0 * N * D


def path_flops(N, D):  # whole posterior+logpdf path
    return float(N) * D * (D + 1) + 4.0 * N * D + D**3 / 3.0 + 4.0 * D * D


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_pass(ref, f, X, y, σ2):
    """What a user of the reference incurs for posterior + logpdf: two independent calls, each running
    __compute_inference_quantities (src/bayesian_linear_regression.jl:56,:61)."""
    fx = f(ref.ColVecs(X), σ2)
    post = ref.posterior(fx, y)
    lp = ref.logpdf(fx, y)
    return post, lp


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (BLAS thread count is stated)."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:  # pragma: no cover
        pass
    return cores


def cpu_sample(D, n_sample, seed=0):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((D, n_sample)))
    σ2 = np.exp(rng.standard_normal(n_sample))
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(n_sample)
    return X, y, σ2


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its literal op sequence on
    scipy/OpenBLAS -- Julia itself is not installable here) with all host threads, on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import blr_oracle as ref

    cores = use_all_host_threads()
    D, n_sample = args.dim, args.cpu_sample
    X, y, σ2 = cpu_sample(D, n_sample)
    f = ref.BayesianLinearRegressor(np.zeros(D), ref.Diagonal(np.ones(D)))
    for _ in range(args.warmup):
        cpu_reference_pass(ref, f, X, y, σ2)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(ref, f, X, y, σ2)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_sample / dt
    sample = f"literal reference op sequence (posterior + logpdf as two calls), N={n_sample} of the D={D} workload per step, scipy/OpenBLAS fp64"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"diagonal-noise BLR posterior+logpdf, D={D}, ColVecs, CPU sample N={n_sample} (cost is linear in N)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch

    import blr_b200 as blr
    from blr_b200 import _lib as L
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    ctx = blr.Context(local)
    blr.set_default_context(ctx)
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner on stdout at communicator creation: keep stdout to the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            ctx.init_comm_from_torch()
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    N, D = args.n_obs, args.dim
    lo, hi = blr.ShardPlan(N, world).bounds(rank)
    n_loc = hi - lo

    # ---- synthetic shard, generated in place on the device (identical data however N is partitioned)
    X = blr.DeviceMatrix.alloc(ctx, D, n_loc).synth_(args.seed, lo)
    σ2 = blr.DeviceVector.alloc(ctx, n_loc)
    y = blr.DeviceVector.alloc(ctx, n_loc)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, σ2.handle, args.seed, lo))
    ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, X.handle, σ2.handle, args.seed, lo, y.handle))
    ctx.sync()

    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    fx = f(blr.ColVecs(X), 0.37 if args.scalar_noise else σ2)
    fx.ctx = ctx

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    gram_ms, solve_ms = [], []

    def step():
        post, lp = blr.posterior_and_logpdf(fx, y)  # public API; returns m', Λ' (host) and logpdf
        t = ctx.last_timings()
        gram_ms.append(t["gram_ms"]); solve_ms.append(t["solve_ms"])
        return post, lp

    for _ in range(args.warmup):
        post, lp = step()
    gram_ms.clear(); solve_ms.clear()

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        post, lp = step()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = N / (ms_per_step * 1e-3)

    out = None
    if rank == 0:
        g_ms = statistics.mean(gram_ms)
        achieved = gram_flops(n_loc, D) / (g_ms * 1e-3) / 1e12
        cal = ctx.calibrate() if not args.no_calibrate else {}
        peak = cal.get("dmma_tflops") or NOMINAL_FP64_TFLOPS
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gram_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"D{D}_N{n_loc}")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"diagonal-noise BLR posterior+logpdf, N={N}, D={D}, fp64, ColVecs, N-sharded over {world} GPU(s)",
                       "n_obs": N, "dim": D, "prior": "mw=0, Λw=I",
                       "noise": "homoscedastic 0.37 I (secondary measurement)" if args.scalar_noise else "heteroscedastic diagonal exp(N(0,1))",
                       "l2": "inputs (%.1f GiB per GPU) far larger than L2; no flush needed" % (n_loc * D * 8 / 2**30),
                       "parallelism": f"obs-sharded x{world}, one NCCL allreduce of D^2+D+3 doubles"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gram_tma_kernel (fp64 DMMA.8x8x4 + TMA bulk copies)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "on-box pure-DMMA issue loop (blr_calibrate_dmma); MEASURED_PEAKS.json has no fp64 entry"
                         if cal else "nominal 148 SM x 64 DFMA/clk x 1.965 GHz",
                         "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                         "calibration": cal, "kernel_ms": g_ms, "solve_ms": statistics.mean(solve_ms),
                         "kernel_share_of_step": g_ms / ms_per_step,
                         "algorithmic_flops_per_launch": gram_flops(n_loc, D), "traffic": traffic},
            "path_tflops": path_flops(N, D) / (ms_per_step * 1e-3) / 1e12 / world,
            "logpdf": lp,
        }

    # ---- e2e: same metric through the public API with HOST buffers (H2D of every chunk + D2H of results timed)
    if (rank == 0 or world > 1) and not args.no_e2e:
        e2e = run_e2e(args, ctx, blr, torch, rank, world)
        if out is not None:
            out["e2e"] = e2e
    # ---- CPU baseline (rank 0, N = 1 only): the reference's literal op sequence on this box's host cores
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import blr_oracle as ref

        use_all_host_threads()
        Xs, ys, ss = cpu_sample(D, args.cpu_sample)
        fo = ref.BayesianLinearRegressor(np.zeros(D), ref.Diagonal(np.ones(D)))
        t0 = time.perf_counter()
        cpu_reference_pass(ref, fo, Xs, ys, ss)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        ref.infer_streaming(np.zeros(D), ref.Diagonal(np.ones(D)), Xs, ys, ss)
        dts = time.perf_counter() - t1
        out["cpu_baseline"] = {
            "value": args.cpu_sample / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"literal reference op sequence (posterior + logpdf as two calls) on N={args.cpu_sample} of the D={D} workload, scipy/OpenBLAS fp64, one pass",
            "streaming_gram_variant_obs_per_s": args.cpu_sample / dts,
        }
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, ctx, blr, torch, rank, world):
    """posterior+logpdf through the public API from pinned HOST memory: every step uploads its observations
    chunk by chunk (H2D inside the timed region), accumulates statistics, all-reduces, solves, and reads back
    the posterior mean / precision / logpdf (D2H)."""
    import ctypes as C
    from blr_b200.runtime import make_noise

    D = args.dim
    n_host = args.e2e_obs  # observations resident in EACH rank's pinned host buffer (every GPU has its own PCIe link)
    chunk = min(args.e2e_chunk, n_host)
    Xh = torch.empty((n_host, D), dtype=torch.float64, pin_memory=True)  # = column-major D x n_host
    yh = torch.empty(n_host, dtype=torch.float64, pin_memory=True)
    sh = torch.empty(n_host, dtype=torch.float64, pin_memory=True)
    g = torch.Generator().manual_seed(1234 + rank)
    for a in range(0, n_host, 1 << 16):  # fill in slices to bound temporary memory
        b = min(n_host, a + (1 << 16))
        Xh[a:b].normal_(generator=g)
    sh.normal_(generator=g).exp_()
    yh.normal_(generator=g)
    Xn, yn, sn = Xh.numpy(), yh.numpy(), sh.numpy()
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))

    def step():
        return blr.posterior_and_logpdf_streamed(f, Xn.T, yn, sn, chunk=chunk, ctx=ctx)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    torch.cuda.synchronize(); ctx.sync()
    if world > 1:
        torch.distributed.barrier()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        post, lp = step()
    ctx.sync()
    dt = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt = float(t.item())
    return {"value": n_host * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(n_host * (D + 2) * 8),
            "d2h_bytes_per_step": int((D * D + D + 1) * 8), "obs_per_step": n_host * world, "chunk_obs": chunk,
            "ms_per_step": dt * 1e3,
            "note": "host-resident sample of the workload (the full 128 GiB matrix does not fit host RAM); pinned memory; wall clock around the public API call"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-obs", type=int, default=N_FULL)
    ap.add_argument("--dim", type=int, default=D_FULL)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=1 << 15, help="observations in the CPU-baseline sample")
    ap.add_argument("--e2e-obs", type=int, default=1 << 19, help="host-resident observations for the e2e leg")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-calibrate", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scalar-noise", action="store_true", help="secondary measurement: Σy = σ² I instead of the headline's heteroscedastic noise")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
