#!/usr/bin/env python
"""BASELINE.json config 5: BasisFunctionRegressor with random Fourier features (D=4096, d_in=32), N=4M observations sharded
over the ranks of one box, posterior+logpdf (fp64).  Launch with torchrun (one rank per GPU) or plainly for one GPU.
The feature map is re-evaluated on the device at every call (as the reference re-evaluates ϕ per call,
src/basis_function_regression.jl:41); ϕ(x) never leaves the GPU."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-obs", type=int, default=1 << 22)
    ap.add_argument("--dim", type=int, default=4096)
    ap.add_argument("--din", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    ctx = blr.Context(local)
    blr.set_default_context(ctx)
    if world > 1:
        import torch.distributed as dist

        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            ctx.init_comm_from_torch()
            dist.barrier()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    N, D, din = args.n_obs, args.dim, args.din
    lo, hi = blr.ShardPlan(N, world).bounds(rank)
    n = hi - lo
    rng = np.random.default_rng(0)  # same feature map on every rank
    rff = blr.RandomFourierFeatures(rng.standard_normal((D, din)), rng.uniform(0, 2 * np.pi, D), ctx)
    xin = blr.DeviceMatrix.alloc(ctx, din, n).synth_(3, lo)
    y, s2 = blr.DeviceVector.alloc(ctx, n), blr.DeviceVector.alloc(ctx, n)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, 5, lo))
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, y.handle, 6, lo))
    bfr = blr.BasisFunctionRegressor(blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D))), rff)
    fx = bfr(blr.ColVecs(xin), s2)
    fx.ctx = ctx

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(args.warmup):
        post, lp = blr.posterior_and_logpdf(fx, y)
    barrier()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        post, lp = blr.posterior_and_logpdf(fx, y)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        fl = N * D * (D + 1) + 4 * N * D + D**3 / 3 + 4 * D * D + 2 * N * D * din
        tm = ctx.last_timings()
        print(json.dumps({"config": f"cfg5 BasisFunctionRegressor RFF D={D} d_in={din} N={N} over {world} GPU(s), posterior+logpdf fp64, phi on device per call",
                          "obs_per_s": N / ms * 1e3, "ms_per_step": ms, "n_gpus": world, "tflops_per_gpu": fl / ms / 1e9 / world,
                          "gram_ms": tm["gram_ms"], "solve_ms": tm["solve_ms"], "logpdf": lp}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
