#!/usr/bin/env python
"""Small-D regime (few features, many observations): posterior+logpdf through the HBM-bound streaming Gram kernel.
Reports obs/s and the algorithmic HBM rate 8*N*(D+2) bytes / step time against the measured copy bandwidth."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402

ctx = blr.Context(0)
blr.set_default_context(ctx)
hbm = ctx.calibrate()["hbm_gbs"]
SHAPES = [(2, 1 << 26), (8, 1 << 26), (16, 1 << 25), (24, 1 << 25), (32, 1 << 25), (40, 1 << 24), (48, 1 << 24), (50, 1 << 24), (64, 1 << 24)]
if os.environ.get("BLR_BENCH_DS"):  # e.g. BLR_BENCH_DS=72,96,100,127,128 -- the trough between the small-D kernels and the tiled ones
    SHAPES = [(int(d), (1 << 30) // (int(d) * 8) // 1024 * 1024) for d in os.environ["BLR_BENCH_DS"].split(",")]
for D, N in SHAPES:
    X = blr.DeviceMatrix.alloc(ctx, D, N).synth_(0)
    s2, y = blr.DeviceVector.alloc(ctx, N), blr.DeviceVector.alloc(ctx, N)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, 0, 0))
    ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, X.handle, s2.handle, 0, 0, y.handle))
    mw0 = 0.1 * np.random.default_rng(3).standard_normal(D) if os.environ.get("BLR_BENCH_MW") == "1" else np.zeros(D)
    f = blr.BayesianLinearRegressor(mw0, blr.Diagonal(np.ones(D)))
    fx = f(blr.ColVecs(X), s2)
    for _ in range(3):
        blr.posterior_and_logpdf(fx, y)
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        post, lp = blr.posterior_and_logpdf(fx, y)
    e1.record(st)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tm = ctx.last_timings()
    gbs = 8.0 * N * (D + 2) / ms / 1e6
    import ctypes as C
    from blr_b200.runtime import make_noise
    dpost = post._device(ctx)
    mv = torch.empty(2 * N, dtype=torch.float64, device="cuda")
    noise, keep = make_noise(ctx, 0.1, N)
    def mean_var():
        ctx.check(ctx.lib.blr_mean_var_dev(ctx.handle, dpost.handle, X.handle, C.byref(noise), C.c_void_p(mv.data_ptr()),
                                           C.c_void_p(mv.data_ptr() + 8 * N)))
    mean_var(); mean_var(); ctx.sync()
    e0.record(st)
    for _ in range(5):
        mean_var()
    e1.record(st)
    ctx.sync(); torch.cuda.synchronize()
    ms_mv = e0.elapsed_time(e1) / 5
    print(json.dumps({"config": f"small-D mean_and_var D={D} N*={N}", "points_per_s": N / ms_mv * 1e3, "ms": ms_mv,
                      "hbm_gbs_algorithmic": 8.0 * N * (D + 2) / ms_mv / 1e6, "frac_of_measured_hbm": 8.0 * N * (D + 2) / ms_mv / 1e6 / hbm,
                      "tflops_triangular": N * D * (D + 1.0) / ms_mv / 1e9}))
    del mv
    print(json.dumps({"config": f"small-D posterior+logpdf D={D} N={N}", "obs_per_s": N / ms * 1e3, "ms": ms, "gram_ms": tm["gram_ms"],
                      "prep_ms": tm["prep_ms"], "hbm_gbs_algorithmic": gbs, "frac_of_measured_hbm": gbs / hbm,
                      "gram_tflops": N * D * (D + 3.0) / tm["gram_ms"] / 1e9}))
    del X, s2, y, fx
