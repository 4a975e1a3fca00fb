# SDPSRCuda.jl -- the reference-side binding a maintainer would add to use libsdpsr_cuda.so
# from SDPSymmetryReduction.jl.  NOT EXECUTED in this repository (the build image has no Julia);
# the same call sequence IS executed on the GPU box by two other clients of the same C ABI:
#   * tests/cabi_client.c           (plain C; tests/test_gpu_cabi_client.py holds its results against the oracle)
#   * sdpsymmetryreduction.jl_b200/{binding,api}.py   (every other GPU test)
# The C ABI (include/sdpsr.h) is the single source of truth.
#
#   using SDPSymmetryReduction, SDPSRCuda
#   P    = admissible_subspace(CuPartition, C, A, b)     # plug-in point: src/partitions.jl:109-116
#   blkD = blockDiagonalize(P)                            # src/compat.jl:26-68
#   blkD = blockDiagonalize(P; complex=true)              # src/compat.jl:46-68 with T = ComplexF64
module SDPSRCuda

import SDPSymmetryReduction as SR
using LinearAlgebra, SparseArrays, Random
import Krylov
import DataStructures: IntDisjointSets, union!, find_root!

const LIB = get(ENV, "SDPSR_LIB", "libsdpsr_cuda")
const E_KRYLOV = Cint(-12)      # SDPSR_E_KRYLOV: the module path does not apply -> dense path, same draws

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

lasterr(c::Ctx) = unsafe_string(ccall((:sdpsr_last_error, LIB), Cstring, (Ptr{Cvoid},), c.h))

function check(c::Ctx, st::Cint)
    st == 0 && return
    msg = lasterr(c)
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
# (`typeof(P)(XY)`, the constructor form the generic code calls at src/partitions.jl:214, is the pair above)

"Host Partition{T} with the canonical labels (P.matrix), like the reference returns."
function SR.Partition{T}(p::CuPartition) where {T<:Union{UInt8,UInt16,UInt32,UInt64}}
    L = Matrix{T}(undef, size(p))
    GC.@preserve L check(p.ctx, ccall((:sdpsr_partition_get_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint), p.ctx.h, L, sizeof(T)))
    return SR.Partition{T}(p.nparts, L)
end

"deepcopy (src/compat.jl:54, src/partitions.jl:203): a second context holding the same partition."
function Base.deepcopy(p::CuPartition)
    L = Matrix{UInt32}(undef, size(p))
    GC.@preserve L check(p.ctx, ccall((:sdpsr_partition_get_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint), p.ctx.h, L, 4))
    c = Ctx(p.ctx.n); d = Ref{Int64}(0)
    GC.@preserve L check(c, ccall((:sdpsr_partition_set_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Int64}), c.h, L, 4, d))
    return CuPartition(c, d[])
end
Base.:(==)(p::CuPartition, q::CuPartition) = SR.Partition{UInt32}(p) == SR.Partition{UInt32}(q)

function Base.fill!(M::AbstractMatrix{Float64}, p::CuPartition; values::AbstractVector)
    @assert length(values) == SR.dim(p)
    v = Vector{Float64}(values)
    GC.@preserve v M begin
        check(p.ctx, ccall((:sdpsr_fill, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), p.ctx.h, v, length(v)))
        check(p.ctx, ccall((:sdpsr_get_matrix, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), p.ctx.h, 0, M))
    end
    return M
end

function refine_labels!(p::CuPartition, L::Matrix{T}) where {T<:Union{UInt8,UInt16,UInt32,Int64}}
    d = Ref{Int64}(0)
    GC.@preserve L check(p.ctx, ccall((:sdpsr_refine_labels, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Int64}), p.ctx.h, L, sizeof(T), d))
    p.nparts = d[]
    return p
end
"refine!(p, q) (src/partitions.jl:62-66) against a host partition or another device partition."
SR.refine!(p::CuPartition, q::SR.Partition) = refine_labels!(p, Matrix{Int64}(q.matrix))
SR.refine!(p::CuPartition, q::CuPartition) = refine_labels!(p, SR.Partition{UInt32}(q).matrix)

"_constraints(P) (src/diagonalize.jl:42-50; the per-back-end override of test/partitions_set.jl:92)."
function SR._constraints(p::CuPartition)
    c = p.ctx; d = SR.dim(p); z = Ref{Int64}(0)
    check(c, ccall((:sdpsr_partition_zero_count, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), c.h, z))
    total = c.n^2 - z[]
    ptr = Vector{Int64}(undef, d + 1); idx = Vector{UInt32}(undef, max(total, 1))
    GC.@preserve ptr idx check(c, ccall((:sdpsr_partition_constraints, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int64}, Ptr{UInt32}, Int64, Cint), c.h, ptr, idx, total, 1))
    return [idx[ptr[i]+1:ptr[i+1]] for i in 1:d]
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
`0` = default (FP64-grade: 7 digits of 8 bits, 54 magnitude bits), `8` = eight 7-bit digits.  `2..6` is an
experiment knob: the integer products stay exact, but the random coefficients are then quantised to
8(k-1)+6 bits, which weakens the probability-one argument of the randomised closure.
"""
square_digits!(c::Ctx, k::Integer) = check(c, ccall((:sdpsr_set_square_slices, LIB), Cint, (Ptr{Cvoid}, Cint), c.h, k))

"""
    labels_async!(out, P::CuPartition);  wait_labels(P)

Start the export of `P.matrix` into `out` (`Matrix{UInt8|UInt16|UInt32|UInt64}`, N x N; page-locked memory for a real
overlap) on the context's copy stream and return at once, so that `blockDiagonalize(P)` can run meanwhile;
`wait_labels(P)` returns once `out` is complete (throws for a label that does not fit, like the reference's
`InexactError`).
"""
labels_async!(out::Matrix{T}, p::CuPartition) where {T<:Union{UInt8,UInt16,UInt32,UInt64}} =
    GC.@preserve out check(p.ctx, ccall((:sdpsr_partition_get_labels_async, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint), p.ctx.h, out, sizeof(T)))
wait_labels(p::CuPartition) = check(p.ctx, ccall((:sdpsr_partition_labels_wait, LIB), Cint, (Ptr{Cvoid},), p.ctx.h))

"Drop-in with the reference's signature: returns a host Partition{UInt16} (src/partitions.jl:77-85)."
admissible_subspace_cuda(C, A, b; kw...) = SR.Partition{UInt16}(SR.admissible_subspace(CuPartition, C, A, b; kw...))

# --- desymmetrize (src/partitions.jl:197-223): X*Y products on the device ------------------------------
function SR.desymmetrize(P::CuPartition; verbose=false, atol=Base.rtoldefault(Float64))
    P = deepcopy(P); c = P.ctx; cur = SR.dim(P); d = Ref{Int64}(0); it = 0
    while true
        it += 1
        rx = rand(Float64, cur); ry = rand(Float64, cur)                    # :210-211
        GC.@preserve rx ry check(c, ccall((:sdpsr_product_round_refine, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Float64, Ref{Int64}), c.h, rx, ry, cur, atol, d))
        d[] == cur && break
        cur = d[]
    end
    verbose && @info "desymmetization converged in $it iterations"
    P.nparts = cur
    return P
end

# --- scalar steps shared by every path: Otsu threshold + union-find (src/eigen_decomposition.jl:205-219) ----
function isomorphism_classes(norms::Matrix{Float64}, atol)
    ne = size(norms, 1)
    thr = SR.otsu_threshold(norms, atol=atol)
    K = IntDisjointSets(ne)
    for i in 1:ne, j in (i+1):ne
        norms[i, j] ≥ thr && union!(K, i, j)
    end
    SR.__isconsistent(K) || throw(SR.NumericalInconsistency("eigen_decomposition",
        "the K-partition seems inconsistent with eigenspaces. Decrease `atol`, or simply try again."))
    return Int64[find_root!(K, i) - 1 for i in 1:ne]
end

"diagonalize through the module path (csrc/krylov.cu); returns block sizes or `nothing` when it does not apply."
function diagonalize_module!(c::Ctx, dimP::Int, r1, r2, r3, atol)
    maxmod = min(2dimP + 16, c.n)
    vals = Vector{Float64}(undef, maxmod); mult = Vector{Int64}(undef, maxmod); ne = Ref{Int64}(0)
    st = GC.@preserve r1 vals mult ccall((:sdpsr_eig_krylov, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Float64, Ptr{Float64}, Ptr{Int64}, Ref{Int64}),
        c.h, r1, dimP, maxmod, atol, vals, mult, ne)
    st == E_KRYLOV && return nothing
    check(c, st)
    norms = Matrix{Float64}(undef, ne[], ne[])
    st = GC.@preserve r2 norms ccall((:sdpsr_block_norms_krylov, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), c.h, r2, dimP, norms)
    st == E_KRYLOV && return nothing
    check(c, st)
    kroot = try
        isomorphism_classes(norms, atol)
    catch e
        e isa SR.NumericalInconsistency || rethrow()
        return nothing                                # let the reference's own statistic decide (dense path)
    end
    sizes = Vector{Int64}(undef, ne[]); nblk = Ref{Int64}(0)
    st = GC.@preserve r3 kroot sizes ccall((:sdpsr_irreducible_krylov, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Float64, Ptr{Int64}, Ref{Int64}),
        c.h, r3, dimP, kroot, atol, sizes, nblk)
    st == E_KRYLOV && return nothing
    check(c, st)
    resize!(sizes, nblk[])
    sum(s -> (s + 1) * s ÷ 2, sizes) == dimP || return nothing
    return sizes
end

"the reference's algorithm step by step: cuSOLVER syevd + Q'AQ on the FP64 DMMA GEMM"
function diagonalize_dense!(c::Ctx, dimP::Int, r1, r2, r3, atol)
    n = c.n; vals = Vector{Float64}(undef, n)
    GC.@preserve r1 vals check(c, ccall((:sdpsr_eig, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}), c.h, r1, dimP, vals))
    eigdec = SR.EigenDecomposition(vals, zeros(0, 0); atol=atol)          # cluster boundaries only
    ptrs = Vector{Int64}(eigdec.ptrs .- 1); ne = length(ptrs) - 1
    norms = Matrix{Float64}(undef, ne, ne)
    GC.@preserve r2 ptrs norms check(c, ccall((:sdpsr_block_norms, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Float64}), c.h, r2, dimP, ptrs, ne + 1, norms))
    kroot = isomorphism_classes(norms, atol)
    sizes = Vector{Int64}(undef, ne); nblk = Ref{Int64}(0)
    GC.@preserve r3 ptrs kroot sizes check(c, ccall((:sdpsr_irreducible, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Float64, Ptr{Int64}, Ref{Int64}),
        c.h, r3, dimP, ptrs, ne + 1, kroot, atol, sizes, nblk))
    return resize!(sizes, nblk[])
end

function qhat(c::Ctx, sizes)
    S = sum(sizes); buf = Matrix{Float64}(undef, c.n, S)
    GC.@preserve buf check(c, ccall((:sdpsr_get_qhat, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, buf, length(buf)))
    out = Matrix{Float64}[]; o = 0
    for s in sizes
        push!(out, buf[:, o+1:o+s]); o += s
    end
    return out
end

"""
diagonalize(Float64, P::CuPartition) (src/diagonalize.jl:25-40).  The three coefficient vectors are drawn
first, in the reference's order (:242, :259, :306), so that the dense fallback consumes the SAME draws.
`eig = :auto | :module | :syevd`.
"""
function SR.diagonalize(::Type{Float64}, P::CuPartition; verbose=false, atol=1e-12 * size(P, 1), eig::Symbol=:auto,
        fetch::Bool=true)
    c = P.ctx; dimP = SR.dim(P)
    r1 = rand(Float64, dimP); r2 = rand(Float64, dimP); r3 = rand(Float64, dimP)
    sizes = eig == :syevd ? nothing : diagonalize_module!(c, dimP, r1, r2, r3, atol)
    if sizes === nothing
        eig == :module && throw(SR.NumericalInconsistency("diagonalize", "the module path does not apply"))
        sizes = diagonalize_dense!(c, dimP, r1, r2, r3, atol)
    end
    return fetch ? qhat(c, sizes) : sizes
end

function blocks_from(out::Vector{T}, dimP, sizes) where {T}
    blks = Vector{Vector{Matrix{T}}}(undef, dimP); off = 0
    for i in 1:dimP
        blks[i] = Matrix{T}[]
        for s in sizes
            push!(blks[i], reshape(out[off+1:off+s*s], s, s)); off += s * s
        end
    end
    return blks
end

"blockDiagonalize(P::CuPartition) (src/compat.jl:26-68); scalar steps reuse the reference's own functions."
function SR.blockDiagonalize(P::CuPartition, verbose=true; epsilon=Base.rtoldefault(Float64), complex=false,
        eig::Symbol=:auto)
    complex && return blockDiagonalize_complex(P, verbose; epsilon=epsilon)
    c = P.ctx; n = c.n; dimP = SR.dim(P)
    sizes = SR.diagonalize(Float64, P; verbose=verbose, atol=epsilon, eig=eig, fetch=false)
    sum(s -> (s + 1) * s ÷ 2, sizes) == dimP || throw(DimensionMismatch("Decomposition failed"))   # src/diagonalize.jl:1-11
    sq = sum(abs2, sizes); out = Vector{Float64}(undef, dimP * sq)
    GC.@preserve out check(c, ccall((:sdpsr_basis_image, LIB), Cint,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Int64), c.h, 1e-12 * n, out, length(out)))
    return (blkSizes=Vector{Int}(sizes), blks=blocks_from(out, dimP, sizes))
end

# --- complex path (src/diagonalize.jl:25-40, src/compat.jl:46-68 with T = ComplexF64) ----------------------
function diagonalize_complex!(Pd::CuPartition, atol)
    c = Pd.ctx; n = c.n; dimP = SR.dim(Pd)
    r1 = rand(ComplexF64, dimP); vals = Vector{ComplexF64}(undef, n)
    GC.@preserve r1 vals check(c, ccall((:sdpsr_eig_complex, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Ptr{ComplexF64}), c.h, r1, dimP, vals))
    eigdec = SR.EigenDecomposition(vals, zeros(ComplexF64, 0, 0); atol=atol)
    ptrs = Vector{Int64}(eigdec.ptrs .- 1); ne = length(ptrs) - 1
    r2 = rand(ComplexF64, dimP); norms = Matrix{Float64}(undef, ne, ne)
    GC.@preserve r2 ptrs norms check(c, ccall((:sdpsr_block_norms_complex, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Ptr{Int64}, Int64, Ptr{Float64}), c.h, r2, dimP, ptrs, ne + 1, norms))
    kroot = isomorphism_classes(norms, atol)
    r3 = rand(ComplexF64, dimP); sizes = Vector{Int64}(undef, ne); nblk = Ref{Int64}(0)
    GC.@preserve r3 ptrs kroot sizes check(c, ccall((:sdpsr_irreducible_complex, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Float64, Ptr{Int64}, Ref{Int64}),
        c.h, r3, dimP, ptrs, ne + 1, kroot, atol, sizes, nblk))
    return resize!(sizes, nblk[])
end

function SR.diagonalize(::Type{ComplexF64}, P::CuPartition; verbose=false, atol=1e-12 * size(P, 1))
    Pd = SR.desymmetrize(P; verbose=verbose)                                # default atol (src/diagonalize.jl:26-28)
    sizes = diagonalize_complex!(Pd, atol)
    c = Pd.ctx; S = sum(sizes); buf = Matrix{ComplexF64}(undef, c.n, S)
    GC.@preserve buf check(c, ccall((:sdpsr_get_qhat_complex, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Int64), c.h, buf, length(buf)))
    out = Matrix{ComplexF64}[]; o = 0
    for s in sizes
        push!(out, buf[:, o+1:o+s]); o += s
    end
    return out
end

function blockDiagonalize_complex(P::CuPartition, verbose=true; epsilon=Base.rtoldefault(Float64))
    Pd = SR.desymmetrize(P; verbose=verbose)                                # inside diagonalize (:26-28)
    sizes = diagonalize_complex!(Pd, epsilon)
    P2 = SR.desymmetrize(P; verbose=verbose, atol=epsilon)                  # src/compat.jl:54-57
    sum(abs2, sizes) == SR.dim(P2) || throw(DimensionMismatch("Decomposition failed"))
    P2 == Pd || throw(SR.NumericalInconsistency("desymmetrize", "did not reproduce its own partition; try again"))
    c = Pd.ctx; dimP = SR.dim(Pd); sq = sum(abs2, sizes); out = Vector{ComplexF64}(undef, dimP * sq)
    GC.@preserve out check(c, ccall((:sdpsr_basis_image_complex, LIB), Cint,
        (Ptr{Cvoid}, Float64, Ptr{ComplexF64}, Int64), c.h, 1e-12 * c.n, out, length(out)))
    return (blkSizes=Vector{Int}(sizes), blks=blocks_from(out, dimP, sizes))
end

end # module
