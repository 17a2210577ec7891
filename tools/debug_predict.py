import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blr_b200 as blr
from oracle import blr_oracle as ref

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for D, N in [(128, 64), (256, 96), (512, 64), (1024, 64)]:
    rng = np.random.default_rng(D)
    X = rng.standard_normal((D, N)); mw = rng.standard_normal(D)
    B = rng.standard_normal((D, D)); Lam = B @ B.T + np.eye(D)
    f = blr.BayesianLinearRegressor(mw, Lam); fo = ref.BayesianLinearRegressor(mw, Lam)
    m, v = blr.mean_and_var(f(blr.ColVecs(X), 0.1))
    mo, vo = ref.mean_and_var(fo(ref.ColVecs(X), 0.1))
    print(D, N, "mean", rel(m, mo), "var", rel(v, vo))
    bad = np.argsort(-np.abs(m - mo))[:6]
    print("  worst mean idx", bad, (m - mo)[bad])
    # which features would explain the mean error? project: m - mo = X' (mw_eff - mw)
    d, *_ = np.linalg.lstsq(X.T, m - mo, rcond=None) if N >= D else (None,)
    if d is not None:
        nz = np.where(np.abs(d) > 1e-8)[0]
        print("  mw error support:", nz[:40], d[nz][:8], "true mw there:", mw[nz][:8])
    bv = np.argsort(-np.abs(v - vo))[:6]
    print("  worst var idx", bv, ((v - vo) / vo)[bv])
