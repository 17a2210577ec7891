# Replacement bodies for BayesianLinearRegressors.jl/src/sampling_functions.jl:27-49: the weight draw
# `blr.mw .+ _cholesky(blr.Λw).U \ randn(rng, D, S)` runs on the device (blr_rand_weights); the callable
# evaluates ϕ(X)'w through blr_apply_weights.  UNEXECUTED -- see INTEGRATION.md.
function _device_weights(rng::AbstractRNG, blr::BayesianLinearRegressor, S::Int)
    ctx = LibBLR.default_context()
    Z = randn(rng, length(blr.mw), S)
    return LibBLR.rand_weights(ctx, _device_post(ctx, blr), Z)   # cached device factor (bayesian_linear_regression.jl)
end

function Random.rand(rng::AbstractRNG, b::BLRorBasisFunction)                                      # :27-31
    blr, ϕ = _blr_and_mapping(b)
    return BLRFunctionSample(vec(_device_weights(rng, blr, 1)), ϕ)
end

function Random.rand(rng::AbstractRNG, b::BLRorBasisFunction, dims::Dims)                          # :33-38
    blr, ϕ = _blr_and_mapping(b)
    ws = _device_weights(rng, blr, prod(dims))
    return reshape([BLRFunctionSample(collect(w), ϕ) for w in eachcol(ws)], dims)
end

function Random.rand!(rng::AbstractRNG, A::AbstractArray{<:BLRFunctionSample}, b::BLRorBasisFunction)   # :40-49
    blr, ϕ = _blr_and_mapping(b)
    ws = _device_weights(rng, blr, prod(size(A)))
    for i in LinearIndices(A)
        @inbounds A[i] = BLRFunctionSample(ws[:, i], ϕ)
    end
    return A
end

# (s::BLRFunctionSample)(X) = ϕ(X)'w  (:17-19) on the device
(s::BLRFunctionSample)(X::ColVecs) = LibBLR.apply_weights(LibBLR.default_context(), _device_x(LibBLR.default_context(), s.ϕ(X)), collect(Float64, s.w))
(s::BLRFunctionSample)(X::RowVecs) = LibBLR.apply_weights(LibBLR.default_context(), _device_x(LibBLR.default_context(), s.ϕ(X)), collect(Float64, s.w))
(s::BLRFunctionSample)(X::AbstractMatrix{<:Real}) = s(ColVecs(X))
