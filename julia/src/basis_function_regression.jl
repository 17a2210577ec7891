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
