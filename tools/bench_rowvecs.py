#!/usr/bin/env python
"""ColVecs vs RowVecs (feature-major, native TMA ring) Gram-kernel time on device-resident data."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402

ctx = blr.Context(0)
blr.set_default_context(ctx)
D, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
for layout in (0, 1):
    buf = torch.randn(D * N, dtype=torch.float64, device="cuda")
    X = blr.DeviceMatrix.wrap_torch(ctx, buf.view(N, D) if layout == 0 else buf.view(D, N), layout)
    s2 = blr.DeviceVector.wrap_torch(ctx, torch.rand(N, dtype=torch.float64, device="cuda") + 0.5)
    y = blr.DeviceVector.wrap_torch(ctx, torch.randn(N, dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    fx = f((blr.ColVecs if layout == 0 else blr.RowVecs)(X), s2)
    for _ in range(3):
        blr.posterior_and_logpdf(fx, y)
    tm = ctx.last_timings()
    fl = N * D * (D + 1) + 2.0 * N * D
    print(json.dumps({"layout": "ColVecs" if layout == 0 else "RowVecs", "D": D, "N": N, **tm, "gram_tflops": fl / tm["gram_ms"] / 1e9}))
