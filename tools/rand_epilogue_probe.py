#!/usr/bin/env python
"""What bounds the rand kernel?  Times blr_rand_finite_dev at a cfg4-sized shape with (a) device Philox + Box-Muller draws
(the epilogue costs ~68 scalar fp64 instructions per pair of draws, on the pipe DMMA runs on) and (b) supplied draws Zy (the
epilogue is two loads, two FMAs and two stores per pair), for the kernel selected by BLR_RAND_PP (0 single-group, 2 two-group, default 1 = by mode).

    python tools/rand_epilogue_probe.py [D] [N*] [S]
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402
from blr_b200.runtime import make_noise  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 512
Nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ctx = blr.Context(0)
rng = np.random.default_rng(0)
B = rng.standard_normal((D, D))
f = blr.BayesianLinearRegressor(rng.standard_normal(D), B @ B.T / D + np.eye(D))
dpost = f._device(ctx)
Xt = blr.DeviceMatrix.alloc(ctx, D, Nt).synth_(1)
Y = torch.empty(Nt * S, dtype=torch.float64, device="cuda")
Zy = torch.randn(Nt * S, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
noise, keep = make_noise(ctx, 0.1, Nt)
st = torch.cuda.ExternalStream(ctx.stream())
out = {"D": D, "N": Nt, "S": S, "BLR_RAND_PP": os.environ.get("BLR_RAND_PP", "default (1: two-group for supplied draws, single-group for device draws)"),
       "BLR_RAND_UNFUSED": os.environ.get("BLR_RAND_UNFUSED", "default (0)")}
for label, zy in (("device_philox", None), ("supplied_draws", C.c_void_p(Zy.data_ptr()))):
    def run():
        ctx.check(ctx.lib.blr_rand_finite_dev(ctx.handle, dpost.handle, Xt.handle, C.byref(noise), S, None, zy, 7, C.c_void_p(Y.data_ptr())))
    run(); run(); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        run()
    e1.record(st)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[label] = {"ms": ms, "tflops": 2.0 * Nt * D * S / ms / 1e9, "finite": bool(torch.isfinite(Y[: 1 << 16]).all())}
print(json.dumps(out))
