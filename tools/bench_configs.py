#!/usr/bin/env python
"""Secondary BASELINE.json configs on ONE GPU (reported beside the headline; results -> profiles/):
  cfg2  posterior+logpdf, N=2^20, D=256
  cfg4  mean_and_var on test points at D=512 (N* = 2^24 per GPU: the 64M-point config is 256 GiB = 8 GPUs x 2^23 ... here
        one GPU's share at 2x), plus rand with 64 function samples on 2^22 points
  cfg5  BasisFunctionRegressor with device-resident RFF (D=4096, d_in=32), posterior+logpdf, N=2^19 per GPU
        (the N=4M config sharded over 8 GPUs)
Each line: config, units/s, ms, algorithmic TFLOP/s and fraction of the on-box DMMA peak.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402
from blr_b200.runtime import make_noise  # noqa: E402


def timed(ctx, fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    stream = torch.cuda.ExternalStream(ctx.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ctx = blr.Context(0)
    blr.set_default_context(ctx)
    peak = ctx.calibrate()["dmma_tflops"]
    out = []

    def synth(D, N, seed=0):
        X = blr.DeviceMatrix.alloc(ctx, D, N).synth_(seed)
        s2 = blr.DeviceVector.alloc(ctx, N)
        y = blr.DeviceVector.alloc(ctx, N)
        ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, seed, 0))
        ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, X.handle, s2.handle, seed, 0, y.handle))
        return X, y, s2

    # ---- cfg2
    D, N = 256, 1 << 20
    X, y, s2 = synth(D, N)
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    ms = timed(ctx, lambda: blr.posterior_and_logpdf(f(blr.ColVecs(X), s2), y), 3, 10)
    fl = N * D * (D + 1) + 4 * N * D + D**3 / 3 + 4 * D * D
    out.append({"config": "cfg2 posterior+logpdf N=2^20 D=256", "obs_per_s": N / ms * 1e3, "ms": ms, "tflops": fl / ms / 1e9,
                "frac_of_dmma_peak": fl / ms / 1e9 / peak, "timings": ctx.last_timings()})
    del X, y, s2

    # ---- cfg4: fit at D=512, then marginals on N* points and rand S=64
    D, N = 512, 1 << 20
    X, y, s2 = synth(D, N)
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    post, lp = blr.posterior_and_logpdf(f(blr.ColVecs(X), s2), y)
    del X, y, s2
    Nt = 1 << 24
    Xt = blr.DeviceMatrix.alloc(ctx, D, Nt).synth_(1)
    dpost = post._device(ctx)
    mv = torch.empty(2 * Nt, dtype=torch.float64, device="cuda")
    noise, keep = make_noise(ctx, 0.1, Nt)

    def mean_var():
        ctx.check(ctx.lib.blr_mean_var_dev(ctx.handle, dpost.handle, Xt.handle, C.byref(noise), C.c_void_p(mv.data_ptr()),
                                           C.c_void_p(mv.data_ptr() + 8 * Nt)))

    ms = timed(ctx, mean_var, 2, 3)
    fl = Nt * D * (D + 1) + 2 * Nt * D
    out.append({"config": "cfg4 mean_and_var D=512 N*=2^24 (device-resident in/out)", "points_per_s": Nt / ms * 1e3, "ms": ms,
                "tflops": fl / ms / 1e9, "frac_of_dmma_peak": fl / ms / 1e9 / peak,
                "hbm_gbs_algorithmic": 8 * Nt * (D + 3) / ms / 1e6})
    del mv
    Nr, S = 1 << 22, 64
    Y = torch.empty(Nr * S, dtype=torch.float64, device="cuda")
    Xr = blr.DeviceMatrix.alloc(ctx, D, Nr).synth_(2)
    noise_r, keep_r = make_noise(ctx, 0.1, Nr)

    def rand():
        ctx.check(ctx.lib.blr_rand_finite_dev(ctx.handle, dpost.handle, Xr.handle, C.byref(noise_r), S, None, None, 7,
                                              C.c_void_p(Y.data_ptr())))

    ms = timed(ctx, rand, 2, 3)
    fl = 2 * Nr * D * S + D * D * S
    out.append({"config": "cfg4 rand S=64 D=512 N*=2^22 (device Philox draws, device-resident out)", "points_per_s": Nr / ms * 1e3,
                "ms": ms, "tflops": fl / ms / 1e9, "frac_of_dmma_peak": fl / ms / 1e9 / peak,
                "hbm_gbs_algorithmic": 8 * Nr * (D + S) / ms / 1e6})
    del Y, Xr, Xt

    # ---- cfg5: RFF features resident on device, D=4096
    din, D, N = 32, 4096, 1 << 19
    rng = np.random.default_rng(0)
    xin = blr.DeviceMatrix.alloc(ctx, din, N).synth_(3)
    rff = blr.RandomFourierFeatures(rng.standard_normal((D, din)), rng.uniform(0, 2 * np.pi, D), ctx)
    y = blr.DeviceVector.alloc(ctx, N)
    s2 = blr.DeviceVector.alloc(ctx, N)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, 5, 0))
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, y.handle, 6, 0))
    bfr = blr.BasisFunctionRegressor(blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D))), rff)
    t0 = time.perf_counter()
    ms = timed(ctx, lambda: blr.posterior_and_logpdf(bfr(blr.ColVecs(xin), s2), y), 1, 2)
    fl = N * D * (D + 1) + 4 * N * D + D**3 / 3 + 4 * D * D + 2 * N * D * din
    out.append({"config": "cfg5 BasisFunctionRegressor RFF D=4096 d_in=32 N=2^19 (one GPU's share of N=4M over 8), phi recomputed per call",
                "obs_per_s": N / ms * 1e3, "ms": ms, "tflops": fl / ms / 1e9, "frac_of_dmma_peak": fl / ms / 1e9 / peak,
                "timings": ctx.last_timings()})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
