"""GPU: device-resident feature maps for BasisFunctionRegressor beyond random Fourier features (SURVEY.md 8f.4,
src/basis_function_regression.jl:7-9,34-37: ϕ is any callable).

  * AffineFeatures (blr_x_features): ϕ(x) = scale * act(Wx + b) evaluated by the library, output resident on the device;
  * TorchFeatureMap: ϕ is arbitrary user code in torch; inputs reach it as a zero-copy view of the library's device matrix and
    its output is borrowed back (blr_x_wrap_device + stream ordering) -- ϕ(x) never visits the host;
  * the reference suite's own nonlinear map (test/test_utils.jl:27-30: hcat(ones, prod.(x))) written that way.
Each against the oracle evaluating the same ϕ in numpy (BFR ≡ BLR∘ϕ, test/basis_function_regression.jl:13-28)."""
import math

import numpy as np
import pytest

import blr_b200 as blr
from oracle import blr_oracle as ref

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_bfr(ϕ_dev, ϕ_np, x, D, seed=0):
    rng = np.random.default_rng(seed)
    N = x.shape[1]
    σ2 = np.exp(rng.standard_normal(N))
    Φ = ϕ_np(x)
    assert Φ.shape == (D, N)
    y = Φ.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    mw = rng.standard_normal(D)
    B = rng.standard_normal((D, D)) / math.sqrt(D)
    Λ = B @ B.T + np.eye(D)
    bfr = blr.BasisFunctionRegressor(blr.BayesianLinearRegressor(mw, Λ), ϕ_dev)
    post, lp = blr.posterior_and_logpdf(bfr(blr.ColVecs(x), σ2), y)
    fo = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(Φ), σ2)
    po = ref.posterior(fo, y)
    assert abs(lp - ref.logpdf(fo, y)) <= RTOL * abs(lp)
    assert relerr(post.blr.mw, po.mw) < RTOL
    xt = x[:, : min(N, 200)]
    m, v = blr.mean_and_var(post(blr.ColVecs(xt), 0.3))
    mo, vo = ref.mean_and_var(po(ref.ColVecs(ϕ_np(xt)), 0.3))
    assert relerr(m, mo) < RTOL and relerr(v, vo) < RTOL
    # function samples evaluate ϕ on the device too (src/sampling_functions.jl:17-19)
    g = blr.rand(np.random.default_rng(5), post)
    assert relerr(g(blr.ColVecs(xt)), ϕ_np(xt).T @ g.w) < RTOL


@pytest.mark.parametrize("act,scale", [("tanh", 1.0), ("relu", 0.5), ("identity", 2.0), ("sin", 1.0), ("cos", math.sqrt(2.0 / 160))])
def test_affine_features(act, scale):
    rng = np.random.default_rng(1)
    din, D, N = 6, 160, 2500
    x = rng.standard_normal((din, N))
    W, b = rng.standard_normal((D, din)) / math.sqrt(din), rng.standard_normal(D)
    fn = {"tanh": np.tanh, "relu": lambda z: np.maximum(z, 0.0), "identity": lambda z: z, "sin": np.sin, "cos": np.cos}[act]
    check_bfr(blr.AffineFeatures(W, b, act, scale), lambda z: scale * fn(W @ z + b[:, None]), x, D)


def test_torch_feature_map_two_layer_network():
    """A user-defined two-layer feature extractor in torch (the shape of examples/nn-blr.jl): hidden features stay on the GPU."""
    import torch

    rng = np.random.default_rng(2)
    din, H, D, N = 5, 40, 192, 3000
    x = rng.standard_normal((din, N))
    W1, b1 = rng.standard_normal((H, din)) / math.sqrt(din), rng.standard_normal(H)
    W2, b2 = rng.standard_normal((D, H)) / math.sqrt(H), rng.standard_normal(D)
    W1t, b1t, W2t, b2t = (torch.from_numpy(a).cuda() for a in (W1, b1, W2, b2))

    def net(xin):  # (N, d_in) CUDA tensor -> (N, D) CUDA tensor
        return torch.tanh(torch.tanh(xin @ W1t.T + b1t) @ W2t.T + b2t)

    def net_np(z):
        return np.tanh(W2 @ np.tanh(W1 @ z + b1[:, None]) + b2[:, None])

    check_bfr(blr.TorchFeatureMap(net), net_np, x, D)
    # inputs already resident on the device (a DeviceMatrix): zero-copy into torch and back
    ctx = blr.default_context()
    xd = blr.DeviceMatrix.upload(ctx, x, 0)
    Φ = blr.TorchFeatureMap(net)(blr.ColVecs(xd)).X
    assert Φ.is_cuda and relerr(Φ.cpu().numpy().T, net_np(x)) < 1e-12


def test_reference_suite_feature_map_on_device():
    """test/test_utils.jl:27-30: ϕ(x) = [1; prod(x)] per input, as a TorchFeatureMap (D = 2: the tiny-D kernels)."""
    import torch

    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 400))
    ϕ = blr.TorchFeatureMap(lambda xin: torch.stack([torch.ones(xin.shape[0], dtype=xin.dtype, device=xin.device), xin.prod(dim=1)], dim=1))
    check_bfr(ϕ, lambda z: np.vstack([np.ones(z.shape[1]), np.prod(z, axis=0)]), x, 2)
