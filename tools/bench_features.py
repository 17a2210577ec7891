#!/usr/bin/env python
"""Device feature map alone: ϕ(x) = sqrt(2/D) cos(Wx + b), d_in = 32, D = 4096 (cfg5's map) on N device-resident inputs.
Reports ms per call (CUDA events on the library stream; includes the 1 MiB upload of W and the output allocation), the fp64
instruction-bound estimate and the output write rate; checks sampled columns against numpy."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import blr_b200 as blr  # noqa: E402

ctx = blr.Context(0)
blr.set_default_context(ctx)
D, din = int(os.environ.get("D", 4096)), int(os.environ.get("DIN", 32))
N = int(os.environ.get("N", 1 << 19))
rng = np.random.default_rng(0)
W, b = rng.standard_normal((D, din)), rng.uniform(0, 2 * np.pi, D)
for act in ("cos", "tanh"):
    fm = blr.RandomFourierFeatures(W, b, ctx) if act == "cos" else blr.AffineFeatures(W, b, "tanh", 1.0, ctx)
    X = blr.DeviceMatrix.alloc(ctx, din, N).synth_(1)
    x = blr.ColVecs(X)
    out = fm(x)
    ctx.sync()
    st = torch.cuda.ExternalStream(ctx.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    del out
    e0.record(st)
    for _ in range(reps):
        out = fm(x)
        del out
    e1.record(st)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = fm(x)
    ctx.sync()
    idx = rng.choice(N, 64, replace=False)
    ti = torch.from_numpy(idx).cuda()
    Xh = X.as_torch()[ti].cpu().numpy().T          # d_in x 64
    got = out.X.as_torch()[ti].cpu().numpy().T     # D x 64
    z = W @ Xh + b[:, None]
    want = np.sqrt(2.0 / D) * np.cos(z) if act == "cos" else np.tanh(z)
    err = float(np.abs(got - want).max()) if got is not None else None
    print(json.dumps({"act": act, "D": D, "din": din, "N": N, "ms": ms, "elements_per_s": N * D / ms * 1e3,
                      "write_gbs": 8.0 * N * D / ms / 1e6, "max_abs_err_vs_numpy": err,
                      "lib": os.environ.get("LIBBLR_CUDA", "default")}))
    del out, X, x
