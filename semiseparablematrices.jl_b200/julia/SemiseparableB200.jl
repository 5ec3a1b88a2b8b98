# SemiseparableB200.jl — Julia glue that routes SemiseparableMatrices.jl's structured QR / solve hot path into
# libssmb200.so (C ABI: include/ssm_b200.h) through `ccall`.  Load it AFTER `using SemiseparableMatrices`:
#
#     using SemiseparableMatrices; include("SemiseparableB200.jl"); using .SemiseparableB200
#
# It only adds more specific methods (Float64 element type; band / fill storage of ANY kind — the README's `Vector` / `Fill` fill and the
# lazy storage `Vcat` constructors produce are converted with `Matrix(...)` before upload) for the generic functions the
# reference already extends, so `qr(A)`, `ldiv!(F, b)`, `F \ b`, `A \ b`, `F.Q' * b`, `F.R \ b` keep their
# names, argument meaning, return types and error behaviour.  There is no CPU fallback: if the library cannot
# be loaded the first call throws.  NOTE: no `julia` binary exists in the build image, so this file is not
# executed by the test-suite; the ctypes twin (ssm_b200/_lib.py) binds the same symbols 1:1 and IS tested, and
# tests/c/replay_julia_calls.c replays the exact `ccall` sequence of every override below (same argument order, strides, nb = 1) on the GPU.
#
# Threads: one library context for the module, every device section runs under DEVLOCK (the C ABI is re-entrant per context, not within one).
# Device copies of live factorisations are found by the IDENTITY of the factorisation's `τ` vector (a mutable object the reference's QR
# carries; AlmostBandedMatrix and BandedPlusSemiseparableMatrix are immutable structs and cannot be WeakKeyDict keys): a `Dict` from
# `objectid(τ)` to the handles, checked against a `WeakRef`, emptied by a finalizer on `τ`.  A QR the table does not know (a `copy`, a
# deserialised one) is uploaded from its own fields (`ssm_abqr_set`).  The library keeps a destroyed context alive until its last object is
# gone, so the order in which Julia runs the atexit hook and the handles' finalizers does not matter.
#
# Overridden reference methods (file:line in the reference's src/):
#   _almostbanded_qr(_, A)                             AlmostBandedMatrix.jl:130-133  (widen + qr!)
#   _almostbanded_qr!(A, τ)                            AlmostBandedMatrix.jl:135-172
#   ldiv!(::QR{<:Any,<:AlmostBandedMatrix}, ::Vector / ::Matrix)   AlmostBandedMatrix.jl:187-199
#   materialize!(MatLdivVec{TriangularLayout{'U',·,AlmostBandedLayout}})   AlmostBandedMatrix.jl:242-256
#   qr(::BandedPlusSemiseparableMatrix)                semiseparableqr.jl:128-131
#   lmul!(::AdjointQ{…QRPackedQ{…BPS}}, ::StridedVector)            semiseparableqr.jl:863-905
#   ldiv!(::UpperTriangular{…BPS}, ::StridedVector)    BandedPlusSemiseparableMatrix.jl:53-70
#   batched extras (no reference equivalent; a batch is a Vector of matrices): qr_batch, solve_batch
module SemiseparableB200

using LinearAlgebra, BandedMatrices, LazyArrays, ArrayLayouts, MatrixFactorizations
using SemiseparableMatrices
import SemiseparableMatrices: AlmostBandedMatrix, AlmostBandedLayout, BandedPlusSemiseparableMatrix, bandpart, fillpart,
                              almostbandwidths, _almostbanded_qr, _almostbanded_qr!
import LinearAlgebra: qr, ldiv!, lmul!, AdjointQ
import MatrixFactorizations: QR, QRPackedQ
import ArrayLayouts: materialize!, MatLdivVec, TriangularLayout, triangulardata
import LazyArrays: arguments

const libssm = get(ENV, "SSM_B200_LIB", joinpath(@__DIR__, "..", "lib", "libssmb200.so"))

const Ctx = Ptr{Cvoid}
const DEVICE = Ref{Cint}(parse(Cint, get(ENV, "SSM_B200_DEVICE", "0")))
const CTX = Ref{Ctx}(C_NULL)
const DEVLOCK = ReentrantLock()          # serialises every device section of this module (one context, one stream, one staging buffer)

struct SSMError <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::SSMError) = print(io, "libssmb200 error ", e.code, ": ", e.msg)

last_error() = unsafe_string(ccall((:ssm_last_error, libssm), Cstring, ()))

# status -> exception, as the reference throws (include/ssm_b200.h "Return value")
function check(rc::Cint)
    rc == 0 && return nothing
    msg = last_error()
    rc == -2 && throw(DimensionMismatch(msg))                       # SSM_E_SHAPE
    rc == -4 && throw(ErrorException(msg))                          # SSM_E_STATE (semiseparableqr.jl:105-107)
    rc < 0 && throw(ArgumentError(msg))
    throw(SSMError(rc, msg))                                        # CUDA runtime failure
end

function context()
    lock(DEVLOCK) do
        if CTX[] == C_NULL
            h = Ref{Ctx}(C_NULL)
            check(ccall((:ssm_ctx_create, libssm), Cint, (Cint, Ptr{Ctx}), DEVICE[], h))
            CTX[] = h[]
            # handles may be finalised after this hook: the library only retires the context then and frees it with its last object
            atexit(() -> lock(DEVLOCK) do
                CTX[] != C_NULL && ccall((:ssm_ctx_destroy, libssm), Cint, (Ctx,), CTX[]); CTX[] = C_NULL
            end)
        end
        CTX[]
    end
end

# Device blocks of destroyed handles (<= 64 MiB each) stay on the library's per-device free list, so qr(A) / ldiv! on one matrix per
# call cost no cudaMalloc / cudaFree; trim_pool() hands them back to the driver (e.g. before another library needs the memory).
trim_pool() = ccall((:ssm_pool_trim, libssm), Int64, (Cint,), DEVICE[])

# ---- thin RAII wrappers around the opaque handles -------------------------------------------------------------------
mutable struct Handle
    p::Ptr{Cvoid}
    destroy::Symbol
    function Handle(p, destroy)
        h = new(p, destroy)
        finalizer(_finalize_handle, h)
        h
    end
end
# Finalizers run on whatever thread the GC runs on: take the device lock without blocking, or try again at the next collection
function _finalize_handle(h)
    h.p == C_NULL && return nothing
    if trylock(DEVLOCK)
        try
            _destroy(h)
        finally
            unlock(DEVLOCK)
        end
    else
        finalizer(_finalize_handle, h)
    end
    nothing
end
function _destroy(h::Handle)
    if h.destroy === :ab;        ccall((:ssm_ab_destroy, libssm), Cint, (Ptr{Cvoid},), h.p)
    elseif h.destroy === :abqr;  ccall((:ssm_abqr_destroy, libssm), Cint, (Ptr{Cvoid},), h.p)
    elseif h.destroy === :bps;   ccall((:ssm_bps_destroy, libssm), Cint, (Ptr{Cvoid},), h.p)
    elseif h.destroy === :bpsqr; ccall((:ssm_bpsqr_destroy, libssm), Cint, (Ptr{Cvoid},), h.p)
    elseif h.destroy === :vec;   ccall((:ssm_vec_destroy, libssm), Cint, (Ptr{Cvoid},), h.p)
    end
    h.p = C_NULL
    nothing
end

function device_vec(b::AbstractVector{Float64}, nb::Integer=1)
    n = length(b) ÷ nb
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ssm_vec_create, libssm), Cint, (Ctx, Cint, Int64, Ptr{Ptr{Cvoid}}), context(), n, nb, h))
    v = Handle(h[], :vec)
    GC.@preserve b check(ccall((:ssm_vec_set, libssm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), v.p, pointer(b), 0))
    v
end
fetch_vec!(b::AbstractVector{Float64}, v::Handle) =
    GC.@preserve b check(ccall((:ssm_vec_get, libssm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), v.p, pointer(b), 0))

# ---- AlmostBandedMatrix ---------------------------------------------------------------------------------------------
const DenseAB = AlmostBandedMatrix{Float64,Matrix{Float64},Matrix{Float64},Matrix{Float64}}

# operands in the reference layout: bands.data is (l+u+1) x n column-major, U is n x r, V is r x n
function upload(bdata::Matrix{Float64}, U::Matrix{Float64}, V::Matrix{Float64}, l::Integer, u::Integer)
    n = size(bdata, 2); r = size(U, 2)
    (size(bdata, 1) == l + u + 1 && size(U, 1) == n && size(V) == (r, n)) ||
        error("Data and fill must be compatible size")                                   # AlmostBandedMatrix.jl:14-16
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ssm_ab_create, libssm), Cint, (Ctx, Cint, Cint, Cint, Cint, Int64, Ptr{Ptr{Cvoid}}), context(), n, l, u, r, 1, h))
    A = Handle(h[], :ab)
    GC.@preserve bdata U V check(ccall((:ssm_ab_set, libssm), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64), A.p, pointer(bdata), 0, pointer(U), 0, pointer(V), 0))
    A
end

# qr on the device, then F.factors / F.τ written straight into reference-layout storage:
#   factors: (2l+u+1) x n band data with bandwidths (l, l+u); U: n x r (overwritten); τ: n
function device_qr!(fdata::Matrix{Float64}, Uout::Matrix{Float64}, τ::Vector{Float64}, A::Handle, ncols::Integer = size(fdata, 2) - 1)
    f = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ssm_ab_qr_cols, libssm), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), A.p, ncols, f))
    F = Handle(f[], :abqr)
    GC.@preserve fdata Uout τ check(ccall((:ssm_abqr_get, libssm), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64), F.p, pointer(fdata), 0, pointer(Uout), 0, pointer(τ), 0))
    F
end

# Device copies of live QR results, so that ldiv!(F, b) does not upload the factors again.  Keyed by the identity of F.τ (see the header).
mutable struct Entry
    key::WeakRef
    operand::Union{Handle,Nothing}
    factor::Handle
end
const FACTORS = Dict{UInt,Entry}()
function remember!(τ::Array{Float64}, operand, factor::Handle)
    lock(DEVLOCK) do
        FACTORS[objectid(τ)] = Entry(WeakRef(τ), operand, factor)
    end
    finalizer(forget!, τ)
    nothing
end
function forget!(τ)
    # finalizers must not block: if the table is busy, run again at the next collection
    if trylock(DEVLOCK)
        try
            e = get(FACTORS, objectid(τ), nothing)
            e !== nothing && (e.key.value === τ || e.key.value === nothing) && delete!(FACTORS, objectid(τ))   # handles go with their own finalizers
        finally
            unlock(DEVLOCK)
        end
    else
        finalizer(forget!, τ)
    end
    nothing
end
function lookup(τ::Array{Float64})
    lock(DEVLOCK) do
        e = get(FACTORS, objectid(τ), nothing)
        (e !== nothing && e.key.value === τ) ? e : nothing
    end
end

# any band / fill storage -> the dense column-major arrays the C ABI takes (README.md:50 builds U::Vector, V::Fill; Vcat constructors give lazy storage)
_mat(X::Matrix{Float64}) = X
_mat(X::AbstractMatrix) = Matrix{Float64}(X)
_mat(X::AbstractVector) = reshape(Vector{Float64}(X), :, 1)
const AnyAB = AlmostBandedMatrix{Float64}

# qr(A) / factorize(A) / A \ b  (AlmostBandedMatrix.jl:125-133).  A is not mutated (test_almostbanded.jl:146).
function _almostbanded_qr(_, A::AnyAB)
    l, u = almostbandwidths(A)
    B, L = bandpart(A), fillpart(A)
    Ua, Va = arguments(L)
    n = size(A, 2)
    (size(A, 1) == n && l >= 0 && u >= 0) || return invoke(_almostbanded_qr, Tuple{Any,Any}, nothing, A)     # rectangular: reference path
    U, V = _mat(Ua), _mat(Va)
    R = AlmostBandedMatrix{Float64}(undef, (n, n), (l, l + u), size(U, 2))                # widened result (:30-38)
    Rb, RL = bandpart(R), fillpart(R)
    RU, RV = arguments(RL)
    copyto!(RV, V)
    τ = zeros(Float64, n)
    lock(DEVLOCK) do
        dA = upload(_mat(B.data), U, V, l, u)
        dF = device_qr!(Rb.data, RU, τ, dA)
        remember!(τ, dA, dF)
    end
    QR(R, τ)
end

# in-place form (:135-172): A already widened to (l, l+u) by the caller, bands u+1..l+u hold the fill values
_almostbanded_qr!(A::DenseAB, τ::Vector{Float64}) = _almostbanded_qr!(A, τ, min(size(A, 1) - 1, size(A, 2)))   # :135-138

# The device sweep forms the widened bands from the fill itself, i.e. it assumes what the widening constructor stores there (:34-36).
# A caller who put anything else into bands u+1 .. l+u (e.g. a partially factored matrix) gets the reference path.
function _widened_is_fill(Bd::Matrix{Float64}, U, V, l::Int, u::Int)
    n = size(Bd, 2)
    @inbounds for j in 1:n, d in (u + 1):(l + u)
        k = j - d
        k >= 1 || continue
        f = 0.0
        for q in axes(U, 2); f += U[k, q] * V[q, j]; end
        Bd[l + u + 1 - d, j] == f || return false
    end
    true
end

# partial form (:139-172): only the first ncols reflectors; the trailing rows / columns come back partially updated
function _almostbanded_qr!(A::DenseAB, τ::Vector{Float64}, ncols::Integer)
    B, L = bandpart(A), fillpart(A)
    l, ū = bandwidths(B)
    u = ū - l
    U, V = arguments(L)
    (u >= 0 && size(A, 1) == size(A, 2) && _widened_is_fill(B.data, U, V, l, u)) ||
        return invoke(_almostbanded_qr!, Tuple{AbstractMatrix,AbstractVector,Any}, A, τ, ncols)
    lock(DEVLOCK) do
        dA = upload(B.data[l+1:end, :], U, V, l, u)                   # rows of bands -l..u of the widened storage
        dF = device_qr!(B.data, U, τ, dA, ncols)
        remember!(τ, dA, dF)
    end
    A, τ
end

# the device copy of F; a factorisation this module did not make (a copy, a deserialised QR) is uploaded from its own fields (ssm_abqr_set)
function device_factor(F::QR{<:Any,<:AnyAB})
    e = lookup(F.τ)
    e !== nothing && return e.factor
    R = F.factors
    l, ū = almostbandwidths(R)
    u = ū - l
    n = size(R, 2)
    u >= 0 || throw(ArgumentError("F.factors must have bandwidths (l, l+u)"))
    Ua, Va = arguments(fillpart(R))
    fdata, U, V = _mat(bandpart(R).data), _mat(Ua), _mat(Va)
    f = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve fdata U V F check(ccall((:ssm_abqr_set, libssm), Cint,
        (Ctx, Cint, Cint, Cint, Cint, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cint, Ptr{Ptr{Cvoid}}),
        context(), n, l, u, size(U, 2), 1, pointer(fdata), 0, pointer(U), 0, pointer(V), 0, pointer(F.τ), 0, -1, f))
    dF = Handle(f[], :abqr)
    remember!(F.τ, nothing, dF)
    dF
end

# ldiv!(F, b) (:187-192) and the matrix right-hand side form (:194-199), in place, returns b
function ldiv!(F::QR{<:Any,<:AnyAB}, b::Vector{Float64})
    length(b) == size(F.factors, 1) || throw(DimensionMismatch("right-hand side has length $(length(b)), matrix has $(size(F.factors,1)) rows"))
    lock(DEVLOCK) do
        dF = device_factor(F)
        v = device_vec(b)
        check(ccall((:ssm_abqr_ldiv, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dF.p, v.p))
        fetch_vec!(b, v)
        _destroy(v)
    end
    b
end
# matrix right-hand side (:194-199): all columns in one pass over the device-resident factors per group of 4 / 8 columns
function ldiv!(F::QR{<:Any,<:AnyAB}, B::Matrix{Float64})
    n, nrhs = size(B)
    n == size(F.factors, 1) || throw(DimensionMismatch("right-hand side has $(n) rows, matrix has $(size(F.factors,1))"))
    lock(DEVLOCK)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    try
        dF = device_factor(F)
        check(ccall((:ssm_mat_create, libssm), Cint, (Ctx, Cint, Cint, Int64, Ptr{Ptr{Cvoid}}), context(), n, nrhs, 1, h))
        GC.@preserve B begin
            check(ccall((:ssm_mat_set, libssm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h[], pointer(B), 0))
            check(ccall((:ssm_abqr_ldiv_mat, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dF.p, h[]))
            check(ccall((:ssm_mat_get, libssm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h[], pointer(B), 0))
        end
    finally
        h[] != C_NULL && ccall((:ssm_mat_destroy, libssm), Cint, (Ptr{Cvoid},), h[])
        unlock(DEVLOCK)
    end
    B
end

# ldiv!(UpperTriangular(R), b) / UnitUpperTriangular on a FACTORED object (:215-256)
for (UNIT, unitflag) in (('N', 0), ('U', 1))
    @eval function materialize!(M::MatLdivVec{TriangularLayout{'U',$UNIT,AlmostBandedLayout},<:Any,<:Any,<:Vector{Float64}})
        R, x = M.A, M.B
        A = triangulardata(R)
        (A isa AnyAB && size(A, 1) == size(A, 2)) || return invoke(materialize!, Tuple{MatLdivVec{TriangularLayout{'U',$UNIT,AlmostBandedLayout}}}, M)
        l, u = almostbandwidths(A)
        U, V = arguments(fillpart(A))
        lock(DEVLOCK) do
            dA = upload(_mat(bandpart(A).data), _mat(U), _mat(V), l, u)
            v = device_vec(x)
            check(ccall((:ssm_ab_tri_ldiv, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint), dA.p, v.p, 1, $unitflag))
            fetch_vec!(x, v)
            _destroy(v); _destroy(dA)
        end
        x
    end
end

# A \ b for ONE system (README.md:96-97, n = 10^6 and beyond): the chunked n-axis path returns the same solution without the
# strictly sequential sweep (solution parity, not factor parity — DESIGN.md section 6).  Measured on B200 for a single system:
# n = 1024: 0.11 ms chunked against 0.44 ms for the sweep; n = 65,536: 0.19 against 27.7 ms.  Smaller systems keep qr + ldiv!.
const CHUNKED_MIN_N = Ref(1 << 10)
function Base.:\(A::AnyAB, b::Vector{Float64})
    n = size(A, 2)
    (size(A, 1) == n && n >= CHUNKED_MIN_N[]) || return ldiv!(qr(A), copy(b))
    length(b) == n || throw(DimensionMismatch("right-hand side has length $(length(b)), matrix has $(n) rows"))
    l, u = almostbandwidths(A)
    Ua, Va = arguments(fillpart(A))
    U, V = _mat(Ua), _mat(Va)
    lock(DEVLOCK)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ssm_abc_create, libssm), Cint, (Ctx, Int64, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}), context(), n, l, u, size(U, 2), h)
    if rc == -3                                                      # no compiled chunked kernel for this (l+u, r): sequential path
        unlock(DEVLOCK)
        return ldiv!(qr(A), copy(b))
    end
    x = similar(b)
    try
        check(rc)
        bd = _mat(bandpart(A).data)
        GC.@preserve bd U V b x begin
            check(ccall((:ssm_abc_set, libssm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), h[], pointer(bd), pointer(U), pointer(V)))
            check(ccall((:ssm_abc_solve, libssm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint), h[], pointer(b), pointer(x), 0))
        end
    finally
        h[] != C_NULL && ccall((:ssm_abc_destroy, libssm), Cint, (Ptr{Cvoid},), h[])
        unlock(DEVLOCK)
    end
    x
end

# ---- adaptive / incremental QR on the device (AlmostBandedMatrix.jl:139-172 on the trailing view, :357-394) ----------------------------
# No reference method of these names exists — the adaptive-QR driver lives in InfiniteLinearAlgebra and calls `_almostbanded_qr!(view, τ, ncols)`;
# these are the device-resident equivalents of its three steps, returning the usual QR object each time.
"""
    partial_qr(A, ncols) -> QR        first `ncols` reflectors only (`_almostbanded_qr!(A, τ, ncols)`, :139-172); the rest of `F.factors` is the
                                      partially updated trailing matrix
    continue_qr!(F, ncols) -> F       more reflectors of the same operand, in place on the device (`ssm_abqr_continue`)
    grow_qr!(F, A) -> F               `A` = the operand after it has grown (`resizedata!`, :357-394; its old block unchanged): the device copy of
                                      the factorisation grows with it (`ssm_ab_resize` + `ssm_ab_set_rows` + `ssm_abqr_resize`); then `continue_qr!`
"""
function partial_qr(A::AnyAB, ncols::Integer)
    l, u = almostbandwidths(A)
    Ua, Va = arguments(fillpart(A))
    n = size(A, 2)
    U, V = _mat(Ua), _mat(Va)
    R = AlmostBandedMatrix{Float64}(undef, (n, n), (l, l + u), size(U, 2))
    RU, RV = arguments(fillpart(R))
    copyto!(RV, V)
    τ = zeros(Float64, n)
    lock(DEVLOCK) do
        dA = upload(_mat(bandpart(A).data), U, V, l, u)
        dF = device_qr!(bandpart(R).data, RU, τ, dA, ncols)
        remember!(τ, dA, dF)
    end
    QR(R, τ)
end
function _refresh!(F::QR, dF::Handle)
    R = F.factors
    RU, _ = arguments(fillpart(R))
    fd = bandpart(R).data
    GC.@preserve fd RU F check(ccall((:ssm_abqr_get, libssm), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64), dF.p, pointer(fd), 0, pointer(RU), 0, pointer(F.τ), 0))
    F
end
function continue_qr!(F::QR{<:Any,<:DenseAB}, ncols::Integer)
    lock(DEVLOCK) do
        e = lookup(F.τ)
        (e === nothing || e.operand === nothing) && error("continue_qr!: the factorisation has no device-resident operand; start from partial_qr(A, ncols)")
        check(ccall((:ssm_abqr_continue, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), e.factor.p, e.operand.p, ncols))
        _refresh!(F, e.factor)
    end
end
function grow_qr!(F::QR{<:Any,<:DenseAB}, A::AnyAB)
    l, u = almostbandwidths(A)
    n0, n1 = size(F.factors, 2), size(A, 2)
    Ua, Va = arguments(fillpart(A))
    bd, U, V = _mat(bandpart(A).data), _mat(Ua), _mat(Va)
    τ = zeros(Float64, n1)
    R = AlmostBandedMatrix{Float64}(undef, (n1, n1), (l, l + u), size(U, 2))
    copyto!(arguments(fillpart(R))[2], V)
    lock(DEVLOCK) do
        e = lookup(F.τ)
        (e === nothing || e.operand === nothing) && error("grow_qr!: the factorisation has no device-resident operand; start from partial_qr(A, ncols)")
        g = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ssm_ab_resize, libssm), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), e.operand.p, n1, g))
        dA = Handle(g[], :ab)
        GC.@preserve bd U V check(ccall((:ssm_ab_set_rows, libssm), Cint,
            (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
            dA.p, max(n0 - u, 0), n1, pointer(bd), 0, pointer(U), 0, pointer(V), 0))       # rows n0-u .. n1-1 (0-based): everything that changed
        check(ccall((:ssm_abqr_resize, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.factor.p, dA.p))
        G = QR(R, τ)
        _refresh!(G, e.factor)
        remember!(τ, dA, e.factor)
        delete!(FACTORS, objectid(F.τ))
        G
    end
end

# ---- batched entry points (no reference equivalent; same-shaped systems stacked along the last axis) -----------------
"""
    solve_batch(bands::Array{Float64,3}, U::Array{Float64,3}, V::Array{Float64,3}, b::Matrix{Float64}, (l,u)) -> x

`x[:,s] = AlmostBandedMatrix(bands[:,:,s], U[:,:,s]*V[:,:,s]) \\ b[:,s]` for every system s, through the chunk-pipelined
host path `ssm_ab_solve_host` (H2D + pack + qr + ldiv! + D2H).
"""
function solve_batch(bands::Array{Float64,3}, U::Array{Float64,3}, V::Array{Float64,3}, b::Matrix{Float64}, (l, u))
    w, n, nb = size(bands)
    r = size(U, 2)
    (w == l + u + 1 && size(U) == (n, r, nb) && size(V) == (r, n, nb) && size(b) == (n, nb)) ||
        throw(DimensionMismatch("Data and fill must be compatible size"))
    x = similar(b)
    GC.@preserve bands U V b x check(ccall((:ssm_ab_solve_host, libssm), Cint,
        (Ctx, Cint, Cint, Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        context(), n, l, u, r, nb, pointer(bands), pointer(U), pointer(V), pointer(b), pointer(x)))
    x
end

# ---- BandedPlusSemiseparableMatrix ----------------------------------------------------------------------------------
# BPS factor objects are found through their τ vector (Q) and through the band storage of F.factors (R = UpperTriangular(F.factors) carries no τ);
# factors of unknown origin keep the reference's own methods
bps_lookup(x::Array{Float64}) = lookup(x)

function upload(A::BandedPlusSemiseparableMatrix{Float64})
    n = size(A, 1); l, m = bandwidths(A.B); r = size(A.U, 2); p = size(A.W, 2)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ssm_bps_create, libssm), Cint, (Ctx, Cint, Cint, Cint, Cint, Cint, Int64, Ptr{Ptr{Cvoid}}), context(), n, l, m, r, p, 1, h))
    dA = Handle(h[], :bps)
    Bd = A.B.data
    GC.@preserve A check(ccall((:ssm_bps_set, libssm), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
        dA.p, pointer(Bd), 0, pointer(A.U), pointer(A.V), 0, pointer(A.W), pointer(A.S), 0))
    dA
end

# qr(A) (semiseparableqr.jl:128-131): QR(BandedPlusSemiseparableMatrix(B (l,l+m), (U,Vnew), ([α β],[S AᵀU])), τ)
function qr(A::BandedPlusSemiseparableMatrix{Float64})
    n = size(A, 1); l, m = bandwidths(A.B); r = size(A.U, 2); p = size(A.W, 2)
    B = BandedMatrix{Float64}(undef, (n, n), (l, l + m))
    U = Matrix{Float64}(undef, n, r); V = similar(U)
    W = Matrix{Float64}(undef, n, p + r); S = similar(W)
    τ = zeros(Float64, n)
    lock(DEVLOCK) do
        dA = upload(A)
        f = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ssm_bps_qr, libssm), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), dA.p, f))
        dF = Handle(f[], :bpsqr)
        GC.@preserve B U V W S τ check(ccall((:ssm_bpsqr_get, libssm), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64),
            dF.p, pointer(B.data), 0, pointer(U), pointer(V), 0, pointer(W), pointer(S), 0, pointer(τ), 0))
        remember!(τ, dA, dF)
        remember!(B.data, dA, dF)
    end
    QR(BandedPlusSemiseparableMatrix(B, (U, V), (W, S)), τ)
end

# lmul!(F.Q', b) (semiseparableqr.jl:863-905).  Q' is found through its τ; a Q of unknown origin keeps the reference's own method.
function lmul!(adjQ::AdjointQ{<:Any,<:QRPackedQ{<:Any,<:BandedPlusSemiseparableMatrix{Float64}}}, b::Vector{Float64})
    e = bps_lookup(parent(adjQ).τ)
    e === nothing && return invoke(lmul!, Tuple{AdjointQ{<:Any,<:QRPackedQ{<:Any,<:BandedPlusSemiseparableMatrix}},StridedVector}, adjQ, b)
    lock(DEVLOCK) do
        v = device_vec(b)
        check(ccall((:ssm_bpsqr_lmul_qt, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.factor.p, v.p))
        fetch_vec!(b, v)
        _destroy(v)
    end
    b
end

# ldiv!(F.R, b) (BandedPlusSemiseparableMatrix.jl:53-70): a factored object -> its device-resident factors; any other BandedPlusSemiseparableMatrix
# is uploaded as an operand (shapes beyond the compiled kernels keep the reference's method)
function ldiv!(R::UpperTriangular{<:Any,<:BandedPlusSemiseparableMatrix{Float64}}, b::Vector{Float64})
    F = parent(R)
    e = bps_lookup(F.B.data)
    lock(DEVLOCK) do
        if e !== nothing
            v = device_vec(b)
            check(ccall((:ssm_bpsqr_upper_ldiv, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.factor.p, v.p))
            fetch_vec!(b, v); _destroy(v)
        else
            n = size(F, 1); l, m = bandwidths(F.B)
            h = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:ssm_bps_create, libssm), Cint, (Ctx, Cint, Cint, Cint, Cint, Cint, Int64, Ptr{Ptr{Cvoid}}), context(), n, l, m, size(F.U, 2), size(F.W, 2), 1, h)
            if rc == -3                                           # SSM_E_UNSUPPORTED
                invoke(ldiv!, Tuple{UpperTriangular{<:Any,<:BandedPlusSemiseparableMatrix},StridedVector}, R, b)
            else
                check(rc)
                ccall((:ssm_bps_destroy, libssm), Cint, (Ptr{Cvoid},), h[])
                v = device_vec(b)
                dA = upload(F)
                check(ccall((:ssm_bps_upper_ldiv, libssm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), dA.p, v.p))
                fetch_vec!(b, v)
                _destroy(v); _destroy(dA)
            end
        end
    end
    b
end

# rq_mul(F.factors, τ) (semiseparableqr.jl:282-287): R*Q of a symmetric BPS factorisation produced by the B200 path
function SemiseparableMatrices.rq_mul(R::BandedPlusSemiseparableMatrix{Float64}, τ::Vector{Float64})
    e = bps_lookup(τ)
    e === nothing && return invoke(SemiseparableMatrices.rq_mul, Tuple{BandedPlusSemiseparableMatrix,Vector}, R, τ)
    n = size(R, 1); l = bandwidth(R.B, 1); r = size(R.U, 2)
    B = BandedMatrix{Float64}(undef, (n, n), (l, l))
    U = Matrix{Float64}(undef, n, r); V = similar(U); W = similar(U); S = similar(U)
    lock(DEVLOCK) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ssm_bpsqr_rq_mul, libssm), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), e.factor.p, h))
        dQ = Handle(h[], :bps)
        GC.@preserve B U V W S check(ccall((:ssm_bps_get, libssm), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64),
            dQ.p, pointer(B.data), 0, pointer(U), pointer(V), 0, pointer(W), pointer(S), 0))
        _destroy(dQ)
    end
    BandedPlusSemiseparableMatrix(B, (U, V), (W, S))
end

end # module
