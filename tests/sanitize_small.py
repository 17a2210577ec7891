#!/usr/bin/env python
"""TEST INFRASTRUCTURE (lives under tests/ because it checks against the oracle).  One small-shape pass over every kernel family of libblr_cuda, to be run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck): SURVEY.md section 5 asks for race / failure evidence for kernels that use
mbarrier rings, setmaxnreg, cross-CTA spin flags and soft grid barriers.  No torch import (keeps the tool's overhead down).

    compute-sanitizer --tool racecheck python tests/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blr_b200 as blr  # noqa: E402
from oracle import blr_oracle as ref  # noqa: E402  (checker only)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def one(ctx, D, N, layout, dense, scalar_noise, Nt, S, tag):
    rng = np.random.default_rng(D + N)
    X = rng.standard_normal((D, N))
    σ2 = 0.37 if scalar_noise else np.exp(rng.standard_normal(N))
    mw = rng.standard_normal(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    if dense:
        B = rng.standard_normal((D, D)) / np.sqrt(D)
        Λ, Λo = B @ B.T + np.eye(D), None
        Λo = Λ
    else:
        lam = np.linspace(0.5, 2.0, D)
        Λ, Λo = blr.Diagonal(lam), ref.Diagonal(lam)
    f = blr.BayesianLinearRegressor(mw, Λ)
    Xd = blr.DeviceMatrix.upload(ctx, X if layout == "col" else np.ascontiguousarray(X.T), 0 if layout == "col" else 1)
    x = blr.ColVecs(Xd) if layout == "col" else blr.RowVecs(Xd)
    s2d = σ2 if scalar_noise else blr.DeviceVector.upload(ctx, σ2)
    fx = f(x, s2d)
    fx.ctx = ctx
    post, lp = blr.posterior_and_logpdf(fx, blr.DeviceVector.upload(ctx, y))
    lp_o, m_o, T_o = ref.infer_streaming(mw, Λo, X, y, σ2)
    e = [abs(lp - lp_o) / abs(lp_o), rel(post.mw, m_o)]
    Xt = rng.standard_normal((D, Nt))
    fp = post(blr.ColVecs(Xt), 0.2)
    fp.ctx = ctx
    m, v = blr.mean_and_var(fp)
    Zw, Zy = rng.standard_normal((D, S)), rng.standard_normal((Nt, S))
    Y = blr.rand_with_draws(fp, Zw, Zy)
    po = ref.BayesianLinearRegressor(m_o, T_o.T @ T_o)
    mo, vo = ref.mean_and_var(po(ref.ColVecs(Xt), 0.2))
    e += [rel(m, mo), rel(v, vo), rel(Y, ref.rand(po(ref.ColVecs(Xt), 0.2), Zw, Zy))]
    Yr = blr.rand(blr.DeviceRNG(3), fp, S)  # device Philox epilogue
    assert np.isfinite(Yr).all()
    # one-pass multi-column logpdf (rhs_multi_kernel, both layouts; k = 3 is a partial column block)
    Ym = np.stack([y, 2.0 * y + 1.0, -y], axis=1)
    lps = blr.logpdf(fx, Ym)
    fo = ref.BayesianLinearRegressor(mw, Λo)(ref.ColVecs(X), σ2)
    e += [abs(lps[j] - ref.logpdf(fo, Ym[:, j])) / abs(lps[j]) for j in range(3)]
    print(f"{tag}: D={D} N={N} {layout} dense={dense} scalar={scalar_noise}  max rel err {max(e):.1e}", flush=True)
    assert max(e) < 1e-9, e


def main():
    which = os.environ.get("BLR_SANITIZE_SET", "all")
    ctx = blr.Context(0)
    cases = [
        # D, N, layout, dense prior, scalar noise, test points, samples, tag
        (2, 300, "col", False, False, 200, 3, "tiny Gram / var"),
        (12, 300, "row", True, False, 200, 3, "small fused Gram"),
        (40, 400, "col", True, True, 200, 3, "small Gram + prep"),
        (130, 700, "col", True, False, 300, 5, "TMA Gram hybrid, narrow tile, var/rand TMA"),
        (256, 1500, "col", False, True, 300, 70, "TMA Gram unit-noise, rand 2 sample blocks"),
        (192, 1024, "row", True, False, 128, 4, "feature-major ring"),
        (320, 900, "col", True, False, 200, 4, "3 tile rows, multi-segment CTAs, D x D fused nb=5"),
        # every tile and every pipeline stage FULL (D % 128 == 0, N % 32 == 0, test points % 128 == 0, S == 64): no slot is ever
        # read beyond what the current phase wrote -- the control for racecheck's reports on the ragged cases above
        (256, 1536, "col", True, False, 384, 64, "aligned control: full tiles and stages only"),
        # the cases above give each of the 148 Gram CTAs at most ONE pipeline stage: this one makes every CTA run ~10 stages
        # through its 3-slot ring (slot reuse = the empty-barrier hand-over), still with full tiles and stages only
        (256, 16384, "col", True, False, 384, 64, "aligned, Gram ring slots reused"),
    ]
    late = [  # kernels added late in round 2 (BLR_SANITIZE_SET=late runs only these)
        (129, 700, "col", True, False, 300, 5, "odd D: repack_colvecs + Gram on D + 1 features, staged var / rand"),
        (129, 700, "row", True, False, 300, 5, "odd D feature-major"),
        (96, 500, "col", True, False, 300, 5, "marginals 128-row pass on 128-point tiles"),
        (24, 400, "col", True, False, 300, 3, "var_small<3>"),
        (80, 700, "col", True, False, 300, 5, "K1m team ring, one stage per team; streaming marginals with W in dynamic shared memory"),
        (90, 40000, "col", True, False, 100, 3, "K1m team ring, slots reused (empty-barrier hand-over), split leftover row"),
    ]
    if which == "late":
        for c in late:
            one(ctx, *c)
        ctxw = blr.Context(0)
        ctxw.set_form("whitened")  # trsm_lower_kernel<false / true>, dxd_whitened, literal var / rand
        one(ctxw, 100, 300, "col", True, False, 100, 5, "whitened form, dense prior")
        one(ctxw, 70, 300, "row", False, True, 40, 3, "whitened form, diagonal prior")
        print("sanitize_small: late set done", flush=True)
        return
    cases += late
    if which == "quick":
        cases = cases[3:5]
    if which.startswith("case:"):  # one case alone (to attribute a sanitizer report to a kernel configuration)
        one(ctx, *cases[int(which[5:])])
        print("sanitize_small: single case done", flush=True)
        return
    for c in cases:
        one(ctx, *c)
    # periodic schedule (soft grid barrier) and the legacy D x D path (wavefront flags)
    os.environ["BLR_GRAM_PERIOD_OBS"] = "256"
    os.environ["BLR_DXD"] = "legacy"
    ctx2 = blr.Context(0)
    del os.environ["BLR_GRAM_PERIOD_OBS"], os.environ["BLR_DXD"]
    one(ctx2, 256, 2000, "col", True, False, 100, 2, "periodic schedule + legacy D x D")
    print("sanitize_small: all passes done", flush=True)


if __name__ == "__main__":
    main()
