#!/usr/bin/env python
"""Run the Gram kernel once per shape so that `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (a single-pass
collection: no replay, so it is affordable at the full cfg3 shape) records its DRAM traffic at EXACTLY the per-GPU shapes
bench.py runs: N = 2^24 / {1, 2, 4, 8}, D = 1024 (and cfg2 / cfg5 shares).  tools/parse_traffic.py turns the CSV into
profiles/gram_traffic.json, which bench.py reports as roofline.traffic for the shape that ran.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gram_tma --csv \
        --log-file gpurun_out/<tag>/traffic.csv python tools/gram_traffic.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blr_b200 as blr  # noqa: E402
from blr_b200 import _lib as L  # noqa: E402

SHAPES = [(1024, 1 << 24), (1024, 1 << 23), (1024, 1 << 22), (1024, 1 << 21), (256, 1 << 20), (4096, 1 << 19)]


def main():
    ctx = blr.Context(0)
    shapes = SHAPES
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for D, N in shapes:
        X = blr.DeviceMatrix.alloc(ctx, D, N).synth_(0, 0)
        s2, y = blr.DeviceVector.alloc(ctx, N), blr.DeviceVector.alloc(ctx, N)
        ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, 0, 0))
        ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, X.handle, s2.handle, 0, 0, y.handle))
        st = blr.Stats(ctx, D)
        noise = L.Noise(L.NOISE_VECTOR, 0.0, s2.handle, None, 0)
        mw = np.zeros(D)
        ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, mw.ctypes.data_as(C.c_void_p), X.handle, y.handle, C.byref(noise)))
        ctx.sync()
        print(f"shape D={D} N={N} done", flush=True)
        del X, s2, y, st
        ctx.sync()


if __name__ == "__main__":
    main()
