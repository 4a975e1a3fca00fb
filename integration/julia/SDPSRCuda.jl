# SDPSRCuda.jl -- the reference-side binding a maintainer would add to use libsdpsr_cuda.so
# from SDPSymmetryReduction.jl.  NOT EXECUTED in this repository (the build image has no Julia);
# it is a thin transliteration of sdpsymmetryreduction.jl_b200/{binding,api}.py, which IS executed
# by the tests.  The C ABI (include/sdpsr.h) is the single source of truth.
#
#   using SDPSymmetryReduction, SDPSRCuda
#   P    = admissible_subspace(CuPartition, C, A, b)     # plug-in point: src/partitions.jl:109-116
#   blkD = blockDiagonalize(P)                            # src/compat.jl:26-68
module SDPSRCuda

import SDPSymmetryReduction as SR
using LinearAlgebra, SparseArrays, Random
import Krylov

const LIB = get(ENV, "SDPSR_LIB", "libsdpsr_cuda")

struct SdpsrError <: Exception
    code::Cint
    msg::String
end

mutable struct Ctx
    h::Ptr{Cvoid}
    n::Int
    function Ctx(n::Integer; device::Integer=0, flags::Integer=0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:sdpsr_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Int64, Cint, UInt32), ref, n, device, flags)
        st == 0 || throw(SdpsrError(st, unsafe_string(ccall((:sdpsr_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        c = new(ref[], n)
        finalizer(x -> ccall((:sdpsr_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end

function check(c::Ctx, st::Cint)
    st == 0 && return
    msg = unsafe_string(ccall((:sdpsr_last_error, LIB), Cstring, (Ptr{Cvoid},), c.h))
    st == -5 && throw(InexactError(:CuPartition, UInt16, msg))                    # SDPSR_E_LABEL_OVERFLOW
    st == -6 && throw(SR.InvalidDecompositionField(Float64, ComplexF64))          # SDPSR_E_NOT_SYMMETRIC
    throw(SdpsrError(st, msg))
end

# --- AbstractPartition back-end (src/abstract_part.jl:1-18) -----------------------------------
mutable struct CuPartition <: SR.AbstractPartition
    ctx::Ctx
    nparts::Int
end
SR.dim(p::CuPartition) = p.nparts
Base.size(p::CuPartition) = (p.ctx.n, p.ctx.n)
Base.size(p::CuPartition, i::Integer) = p.ctx.n

"CuPartition(M): Partition{T}(M) on the device (src/partitions.jl:24-42)."
function CuPartition(M::AbstractMatrix{<:AbstractFloat})
    c = Ctx(size(M, 1)); d = Ref{Int64}(0)
    Md = Matrix{Float64}(M)
    GC.@preserve Md check(c, ccall((:sdpsr_refine_values, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint, Ref{Int64}), c.h, Md, sqrt(eps()), 0, d))
    return CuPartition(c, d[])
end
function CuPartition(M::AbstractMatrix{<:Integer})
    c = Ctx(size(M, 1)); d = Ref{Int64}(0)
    Mi = Matrix{Int64}(M)
    GC.@preserve Mi check(c, ccall((:sdpsr_partition_set_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Int64}), c.h, Mi, 8, d))
    return CuPartition(c, d[])
end

"Host Partition{T} with the canonical labels (P.matrix), like the reference returns."
function SR.Partition{T}(p::CuPartition) where {T<:Union{UInt8,UInt16,UInt32,UInt64}}
    L = Matrix{T}(undef, size(p))
    GC.@preserve L check(p.ctx, ccall((:sdpsr_partition_get_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint), p.ctx.h, L, sizeof(T)))
    return SR.Partition{T}(p.nparts, L)
end

function Base.fill!(M::AbstractMatrix{Float64}, p::CuPartition; values::AbstractVector)
    @assert length(values) == SR.dim(p)
    v = Vector{Float64}(values)
    GC.@preserve v M begin
        check(p.ctx, ccall((:sdpsr_fill, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), p.ctx.h, v, length(v)))
        check(p.ctx, ccall((:sdpsr_get_matrix, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), p.ctx.h, 0, M))
    end
    return M
end

function SR.refine!(p::CuPartition, q::SR.Partition)
    d = Ref{Int64}(0); L = Matrix{Int64}(q.matrix)
    GC.@preserve L check(p.ctx, ccall((:sdpsr_refine_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Int64}), p.ctx.h, L, 8, d))
    p.nparts = d[]
    return p
end

function set_constraints!(c::Ctx, A::SparseMatrixCSC{Float64})
    At = sparse(transpose(A))                # CSC of A' == CSR of A
    rp = Vector{Int64}(At.colptr); ci = Vector{Int64}(At.rowval); vv = At.nzval
    GC.@preserve rp ci vv check(c, ccall((:sdpsr_set_constraints_csr, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint), c.h, size(A, 1), rp, ci, vv, 1))
end
function set_constraints!(c::Ctx, A::AbstractMatrix{Float64})
    Ad = Matrix{Float64}(A)
    GC.@preserve Ad check(c, ccall((:sdpsr_set_constraints_dense, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float64}), c.h, size(A, 1), Ad))
end

"""
admissible_subspace(CuPartition, C, A, b): the specialised driver.  The two initial elements are
computed by the reference's OWN code (qr / Krylov.craig, src/partitions.jl:124-142), so the initial
partition is reference-exact; the loop runs on the device.  rand() is called at exactly the
reference's points with the reference's lengths (src/partitions.jl:159,167).
"""
function SR.admissible_subspace(::Type{CuPartition}, C::AbstractVector{T}, A::AbstractMatrix{T},
        b::AbstractVector{T}; verbose::Bool=false, atol=Base.rtoldefault(real(T))) where {T<:AbstractFloat}
    n = isqrt(length(C)); @assert n^2 == length(C)
    tmp = Vector{T}(undef, length(C))
    projL! = let A′ = A', A′qr = qr(A′)
        (tmp, v) -> SR.project_colspace!(tmp, v, A′, Afact=A′qr)
    end
    CL = let c = Vector(C)
        c .-= projL!(tmp, c); c = SR._clamp_round!(c, atol=atol); c = SR._symmetrize!(c, n); reshape(c, n, n)
    end
    X0 = let (X, _) = Krylov.craig(A, b)
        X = SR._symmetrize!(X, n); X = (tmp = projL!(tmp, X); copyto!(X, tmp))
        X = SR._clamp_round!(X, atol=atol); reshape(X, n, n)
    end
    c = Ctx(n); set_constraints!(c, A); d = Ref{Int64}(0)
    rv(M) = GC.@preserve M check(c, ccall((:sdpsr_refine_values, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Float64, Cint, Ref{Int64}), c.h, M, atol, 0, d))
    rv(Matrix{Float64}(CL)); rv(Matrix{Float64}(X0))                     # :145-146
    cur = d[]; maxdim = (n^2 + n) ÷ 2
    fillr(k) = (r = rand(Float64, k); GC.@preserve r check(c, ccall((:sdpsr_fill, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, r, k)))
    while cur < maxdim                                                     # :154
        fillr(cur)                                                          # :159
        check(c, ccall((:sdpsr_project_round_refine, LIB), Cint, (Ptr{Cvoid}, Float64, Ref{Int64}), c.h, atol, d))
        d[] != cur && fillr(d[])                                            # :166-168
        check(c, ccall((:sdpsr_square_round_refine, LIB), Cint, (Ptr{Cvoid}, Float64, Ref{Int64}), c.h, atol, d))
        cur == d[] && break                                                 # :180-182
        cur = d[]
    end
    return CuPartition(c, d[])
end

"""
    square_digits!(c::Ctx, k)

Number of int8 digits per entry the INT8 tensor-path square uses for `mul!(X², X, X)` (symmetric X):
`0` = default (FP64-grade: 7 digits of 8 bits), `2..7` = coarser coefficients, `8` = eight 7-bit digits.
The integer products are exact whatever `k` is, so the classes found do not depend on it.
"""
square_digits!(c::Ctx, k::Integer) = check(c, ccall((:sdpsr_set_square_slices, LIB), Cint, (Ptr{Cvoid}, Cint), c.h, k))

"Drop-in with the reference's signature: returns a host Partition{UInt16} (src/partitions.jl:77-85)."
admissible_subspace_cuda(C, A, b; kw...) = SR.Partition{UInt16}(SR.admissible_subspace(CuPartition, C, A, b; kw...))

"blockDiagonalize(P::CuPartition): src/compat.jl:46-68; scalar steps reuse the reference's own functions."
function SR.blockDiagonalize(P::CuPartition, verbose=true; epsilon=Base.rtoldefault(Float64), complex=false)
    complex && error("complex path: use the reference implementation on SR.Partition{UInt32}(P)")
    c = P.ctx; n = c.n; dimP = SR.dim(P)
    r1 = rand(Float64, dimP); vals = Vector{Float64}(undef, n)              # src/eigen_decomposition.jl:242
    GC.@preserve r1 vals check(c, ccall((:sdpsr_eig, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), c.h, r1, dimP, vals))
    eigdec = SR.EigenDecomposition(vals, zeros(0, 0); atol=epsilon)         # cluster boundaries only
    ptrs = Vector{Int64}(eigdec.ptrs .- 1); ne = length(ptrs) - 1
    r2 = rand(Float64, dimP); norms = Matrix{Float64}(undef, ne, ne)        # :259
    GC.@preserve r2 ptrs norms check(c, ccall((:sdpsr_block_norms, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Float64}), c.h, r2, dimP, ptrs, ne + 1, norms))
    thr = SR.otsu_threshold(norms, atol=epsilon)
    K = SR.IntDisjointSets(ne)
    for i in 1:ne, j in (i+1):ne
        norms[i, j] ≥ thr && SR.union!(K, i, j)
    end
    SR.__isconsistent(K) || throw(SR.NumericalInconsistency("eigen_decomposition", "inconsistent K"))
    kroot = Int64[SR.find_root!(K, i) - 1 for i in 1:ne]
    r3 = rand(Float64, dimP); sizes = Vector{Int64}(undef, ne); nblk = Ref{Int64}(0)   # :306
    GC.@preserve r3 ptrs kroot sizes check(c, ccall((:sdpsr_irreducible, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Float64, Ptr{Int64}, Ref{Int64}),
        c.h, r3, dimP, ptrs, ne + 1, kroot, epsilon, sizes, nblk))
    resize!(sizes, nblk[])
    sum(s -> (s + 1) * s ÷ 2, sizes) == dimP || throw(DimensionMismatch("Decomposition failed"))   # src/diagonalize.jl:1-11
    sq = sum(abs2, sizes); out = Vector{Float64}(undef, dimP * sq)
    GC.@preserve out check(c, ccall((:sdpsr_basis_image, LIB), Cint,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Int64), c.h, 1e-12 * n, out, length(out)))
    blks = Vector{Vector{Matrix{Float64}}}(undef, dimP); off = 0
    for i in 1:dimP
        blks[i] = Matrix{Float64}[]
        for s in sizes
            push!(blks[i], reshape(out[off+1:off+s*s], s, s)); off += s * s
        end
    end
    return (blkSizes=Vector{Int}(sizes), blks=blks)
end

end # module
