#!/usr/bin/env python
"""Small driver for ncu captures of the prediction-side kernels: mean_and_var + rand at D, N* given on the command line."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402
from blr_b200.runtime import make_noise  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ctx = blr.Context(0)
blr.set_default_context(ctx)
rng = np.random.default_rng(0)
B = rng.standard_normal((D, D))
f = blr.BayesianLinearRegressor(rng.standard_normal(D), B @ B.T / D + np.eye(D))
dpost = f._device(ctx)
Xt = blr.DeviceMatrix.alloc(ctx, D, Nt).synth_(1)
mv = torch.empty(2 * Nt, dtype=torch.float64, device="cuda")
Y = torch.empty(Nt * S, dtype=torch.float64, device="cuda")
noise, keep = make_noise(ctx, 0.1, Nt)
for _ in range(2):
    ctx.check(ctx.lib.blr_mean_var_dev(ctx.handle, dpost.handle, Xt.handle, C.byref(noise), C.c_void_p(mv.data_ptr()),
                                       C.c_void_p(mv.data_ptr() + 8 * Nt)))
    ctx.check(ctx.lib.blr_rand_finite_dev(ctx.handle, dpost.handle, Xt.handle, C.byref(noise), S, None, None, 7,
                                          C.c_void_p(Y.data_ptr())))
ctx.sync()
print("ok", float(mv[:4].sum()), float(Y[:4].sum()))
