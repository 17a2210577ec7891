#!/usr/bin/env python
"""BASELINE.json config 4: posterior marginals (mean_and_var) on 64M test points at D = 512, plus rand with S = 64 function
samples, test points sharded over the ranks of one box (no collective on the prediction side: only the D-sized
posterior is replicated).  Launch with torchrun (one rank per GPU) or plainly for one GPU; with fewer than 8 GPUs
the per-GPU share is capped by --max-log2-per-gpu (the full 256 GiB test matrix needs 8 x 32 GiB).
Inputs and outputs stay on the device (blr_mean_var_dev / blr_rand_finite_dev); draws are device Philox."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402
from blr_b200.runtime import make_noise  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-test", type=int, default=1 << 26)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--n-fit", type=int, default=1 << 20)
    ap.add_argument("--max-log2-per-gpu", type=int, default=24)
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
    D, S = args.dim, args.samples
    Nt = min(args.n_test, world << args.max_log2_per_gpu)
    lo, hi = blr.ShardPlan(Nt, world).bounds(rank)
    n = hi - lo

    # ---- fit: observations sharded, one allreduce, posterior replicated on every rank
    flo, fhi = blr.ShardPlan(args.n_fit, world).bounds(rank)
    Xf = blr.DeviceMatrix.alloc(ctx, D, fhi - flo).synth_(0, flo)
    s2f, yf = blr.DeviceVector.alloc(ctx, fhi - flo), blr.DeviceVector.alloc(ctx, fhi - flo)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2f.handle, 0, flo))
    ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, Xf.handle, s2f.handle, 0, flo, yf.handle))
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    fx = f(blr.ColVecs(Xf), s2f)
    fx.ctx = ctx
    post, lp = blr.posterior_and_logpdf(fx, yf)
    del Xf, s2f, yf, fx
    dpost = post._device(ctx)

    # ---- this rank's test points
    Xt = blr.DeviceMatrix.alloc(ctx, D, n).synth_(1, lo)
    mv = torch.empty(2 * n, dtype=torch.float64, device="cuda")
    Y = torch.empty(n * S, dtype=torch.float64, device="cuda")
    noise, keep = make_noise(ctx, 0.1, n)

    def mean_var():
        ctx.check(ctx.lib.blr_mean_var_dev(ctx.handle, dpost.handle, Xt.handle, C.byref(noise), C.c_void_p(mv.data_ptr()),
                                           C.c_void_p(mv.data_ptr() + 8 * n)))

    def rand():
        ctx.check(ctx.lib.blr_rand_finite_dev(ctx.handle, dpost.handle, Xt.handle, C.byref(noise), S, None, None, 7 + rank,
                                              C.c_void_p(Y.data_ptr())))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ms_mv = timed(mean_var)
    ms_r = timed(rand)
    ok = bool(torch.isfinite(mv).all()) and bool(torch.isfinite(Y[: 1 << 20]).all())
    if rank == 0:
        fl_mv = Nt * D * (D + 1) + 2 * Nt * D
        fl_r = 2 * Nt * D * S + D * D * S
        print(json.dumps({
            "config": f"cfg4 marginals + rand: D={D}, N*={Nt} test points over {world} GPU(s) ({n} per GPU), S={S}, fp64, device-resident in/out",
            "n_gpus": world,
            "mean_and_var": {"points_per_s": Nt / ms_mv * 1e3, "ms": ms_mv, "tflops_per_gpu": fl_mv / ms_mv / 1e9 / world,
                             "hbm_gbs_algorithmic_per_gpu": 8 * Nt * (D + 3) / ms_mv / 1e6 / world},
            "rand": {"points_per_s": Nt / ms_r * 1e3, "ms": ms_r, "tflops_per_gpu": fl_r / ms_r / 1e9 / world,
                     "hbm_gbs_algorithmic_per_gpu": 8 * Nt * (D + S) / ms_r / 1e6 / world},
            "finite": ok, "fit_logpdf": lp}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
