# src/basis_function_regression.jl needs NO change: every FiniteBFR method forwards through
# `_to_finite_blr(fx) = fx.f.blr(fx.f.ϕ(fx.x), fx.Σy)` (:41) to the FiniteBLR methods replaced in
# bayesian_linear_regression.jl.  The only addition is a device-native feature map whose output stays resident
# on the GPU (BASELINE config 5): a ϕ that returns a ColVecs wrapping a device handle instead of a host matrix.
# UNEXECUTED -- see INTEGRATION.md.
struct RandomFourierFeatures
    W::Matrix{Float64}   # D x d_in
    b::Vector{Float64}   # D
end

struct DeviceColVecs <: AbstractVector{Vector{Float64}}
    x::LibBLR.DeviceX
end
Base.size(x::DeviceColVecs) = (x.x.N,)
_device_x(ctx, x::DeviceColVecs) = x.x

function (ϕ::RandomFourierFeatures)(x::ColVecs)
    ctx = LibBLR.default_context()
    xin = LibBLR.upload_x(ctx, x.X, LibBLR.COLVECS)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    LibBLR.check(ctx, ccall((:blr_x_rff, LibBLR.libblr), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
        ctx.ptr, xin.ptr, ϕ.W, ϕ.b, size(ϕ.W, 1), r))
    h = LibBLR.DeviceX(r[], ctx, size(ϕ.W, 1), xin.N)
    finalizer(h -> ccall((:blr_x_free, LibBLR.libblr), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ctx.ptr, h.ptr), h)
    return DeviceColVecs(h)
end

# Any one-layer map  ϕ(x) = scale * act(W x + b)  evaluated by the library (blr_x_features); RandomFourierFeatures is
# act = :cos, scale = sqrt(2 / D).  `act` ∈ (:cos, :tanh, :relu, :identity, :sin).
struct AffineFeatures
    W::Matrix{Float64}   # D x d_in
    b::Vector{Float64}   # D
    act::Symbol
    scale::Float64
end
const _ACT = Dict(:cos => 0, :tanh => 1, :relu => 2, :identity => 3, :sin => 4)

function (ϕ::AffineFeatures)(x::ColVecs)
    ctx = LibBLR.default_context()
    size(ϕ.W, 2) == size(x.X, 1) || throw(DimensionMismatch("size(W, 2) != size(x, 1)"))
    xin = LibBLR.upload_x(ctx, x.X, LibBLR.COLVECS)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    LibBLR.check(ctx, ccall((:blr_x_features, LibBLR.libblr), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Float64, Ref{Ptr{Cvoid}}),
        ctx.ptr, xin.ptr, ϕ.W, ϕ.b, size(ϕ.W, 1), _ACT[ϕ.act], ϕ.scale, r))
    h = LibBLR.DeviceX(r[], ctx, size(ϕ.W, 1), xin.N)
    finalizer(h -> ccall((:blr_x_free, LibBLR.libblr), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ctx.ptr, h.ptr), h)
    return DeviceColVecs(h)
end

# An arbitrary ϕ written on CUDA.jl arrays: borrow its output (D x N CuMatrix{Float64}, column-major = ColVecs) without a copy.
# `stream` is the CUDA.jl stream ϕ ran on (CUDA.stream().handle): the library's stream is ordered after it before reading.
function device_colvecs(Φ_ptr::Ptr{Float64}, D::Integer, N::Integer, ld::Integer, stream::Ptr{Cvoid}; keep=nothing)
    ctx = LibBLR.default_context()
    r = Ref{Ptr{Cvoid}}(C_NULL)
    LibBLR.check(ctx, ccall((:blr_x_wrap_device, LibBLR.libblr), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Cint, Ref{Ptr{Cvoid}}), ctx.ptr, Φ_ptr, D, N, ld, LibBLR.COLVECS, r))
    LibBLR.wait_stream(ctx, stream)
    h = LibBLR.DeviceX(r[], ctx, D, N)
    finalizer(h -> (keep; ccall((:blr_x_free, LibBLR.libblr), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ctx.ptr, h.ptr)), h)  # `keep` pins the CuArray
    return DeviceColVecs(h)
end
