#!/usr/bin/env python
"""Where does a posterior+logpdf step spend its time outside the Gram kernel?  (cfg2's whole step is
~0.45 ms longer than its Gram kernel.)  Prints, for one shape, the device-event phases of the library (prep / Gram / reduce /
D x D), the wall clock of the bare C call with and without result downloads, and the wall clock of the public Python API.

    python tools/step_breakdown.py 256 1048576 [reps]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blr_b200 as blr  # noqa: E402


def main():
    D, N = int(sys.argv[1]), int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    ctx = blr.default_context()
    X = blr.DeviceMatrix.alloc(ctx, D, N).synth_(0, 0)
    σ2, y = blr.DeviceVector.alloc(ctx, N), blr.DeviceVector.alloc(ctx, N)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, σ2.handle, 0, 0))
    ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, X.handle, σ2.handle, 0, 0, y.handle))
    ctx.sync()
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    fx = f(blr.ColVecs(X), σ2)
    fx.ctx = ctx

    def wall(fn):
        for _ in range(5):
            fn()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.sync()
        return (time.perf_counter() - t0) / reps * 1e3

    out = {"D": D, "N": N, "reps": reps}
    out["api_posterior_and_logpdf_ms"] = wall(lambda: blr.posterior_and_logpdf(fx, y))
    out["api_logpdf_only_ms"] = wall(lambda: blr.logpdf(fx, y))
    t = ctx.last_timings()
    out["device_phases_ms"] = t

    # the bare C call, outputs in page-locked / pageable host memory, with and without the D x D downloads
    from blr_b200.model import make_noise

    noise, keep_noise = make_noise(ctx, σ2, N)
    prior, keep_prior = f._prior_struct()
    lp = C.c_double()
    m = np.empty(D)
    for label, Λ in (("pageable", np.empty((D, D), order="F")), ("pinned", ctx.empty_pinned((D, D), min_bytes=0))):
        def call(Λ=Λ):
            h = C.c_void_p()
            ctx.check(ctx.lib.blr_infer(ctx.handle, C.byref(prior), X.handle, y.handle, C.byref(noise), C.byref(lp),
                                        m.ctypes.data_as(C.c_void_p), None, Λ.ctypes.data_as(C.c_void_p), C.byref(h)))
            ctx.lib.blr_post_free(ctx.handle, h)
        out[f"c_call_full_outputs_{label}_ms"] = wall(call)

    def call_lp():
        ctx.check(ctx.lib.blr_infer(ctx.handle, C.byref(prior), X.handle, y.handle, C.byref(noise), C.byref(lp), None, None, None,
                                    None))
    out["c_call_logpdf_only_ms"] = wall(call_lp)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
