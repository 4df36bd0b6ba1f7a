# AD4SMB200.jl -- drop-in B200 backend for AD4SM.jl's `makeϕrKt` path.
#
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia there).  It is the binding a maintainer adds next to AD4SM.jl; the same
# C ABI (include/ad4sm_b200.h) is exercised end to end by the Python/ctypes mirror (ad4sm.jl_b200/_capi.py,
# Elements.py) and the GPU parity tests, so what this file has to get right is only the argument marshalling below.
#
#     using AD4SM; include("AD4SMB200.jl"); using .AD4SMB200      # overrides the methods listed below
#     ϕ, r, Kt = makeϕrKt(elems, u)                                # unchanged call, now on the GPU
#
# Methods re-defined (same signatures and return types as the reference):
#   Elements.makeϕrKt(elems::AbstractVector{<:AbstractCElem}, u::AbstractMatrix)       src/elasticelements.jl:356-358
#   Elements.makeϕrKt(Φ::Vector{<:adiff.D2}, elems::Vector, u)                         src/elements.toolkit.jl:169-203
#   Elements.makeϕrKt(elems::Vector{<:CPElem}, u::Array, d::Array)                     src/phasefieldelements.jl:228-238
#   Elements.makeϕrKt_d(elems, u, d) / makeϕrKt_d(elems, u, d, ϕmax)                   src/phasefieldelements.jl:239-262
# `Solvers.solvestep!` builds Φ with a serial comprehension and calls the 3-argument assembly
# (src/solvers.jl:213-216, 257-260; the direct calls are commented out at :212, :256): to route the Newton loop
# through this backend restore those two lines to `makeϕrKt(elems, u)`.
module AD4SMB200

using AD4SM
using AD4SM.Elements: AbstractCElem, CEElem, CPElem
using AD4SM.Materials
using AD4SM: adiff
using SparseArrays

const lib = get(ENV, "AD4SM_B200_LIB", joinpath(@__DIR__, "..", "lib", "libad4sm_b200.so"))

# ---- include/ad4sm_b200.h ---------------------------------------------------------------------------
const ELEM_CE, ELEM_CP, ELEM_CAS = Int32(0), Int32(1), Int32(2)
const MAT_HOOKE, MAT_HOOKE2D, MAT_NEOHOOKE, MAT_MOONEYRIVLIN, MAT_OGDEN, MAT_PHASEFIELD = Int32.(1:6)
const F_SMALL, F_PLANE_STRESS, F_DHN, F_PF_GRAD_INTENDED, F_NH2D_EXT = UInt32(1), UInt32(2), UInt32(4), UInt32(8), UInt32(16)

struct Batch                      # struct ad4sm_batch
  elem_kind::Int32; dim::Int32; n_gauss::Int32; n_elem_nodes::Int32
  mat_kind::Int32; mat_base::Int32; mat_flags::UInt32; batch_flags::Int32
  mat_params::NTuple{8,Float64}
  n_elems::Int64
  nodes::Ptr{Int64}; gradN::Ptr{Float64}; wgt::Ptr{Float64}; Nval::Ptr{Float64}; orig_index::Ptr{Int64}
  mesh::Ptr{Cvoid}               # NULL, or an ad4sm_mesh: ∇N / wgt / N constructed on the device (ABI version 2)
end
struct Mesh                       # struct ad4sm_mesh
  shape::Int32; n_gp_1d::Int32
  gp_xi::NTuple{3,Float64}; gp_w::NTuple{3,Float64}
  coord_dim::Int32; reserved::Int32
  n_nodes::Int64
  coords::Ptr{Float64}
end
const SHAPE_TRIA3, SHAPE_QUAD4, SHAPE_TET4, SHAPE_HEX8, SHAPE_WDG6 = Int32.(1:5)

check(rc) = rc == 0 || error("libad4sm_b200: " * unsafe_string(ccall((:ad4sm_last_error, lib), Cstring, ())))

# ---- material descriptors (elasticmaterials.jl:25-120, phasefieldmaterials.jl:5-16) ------------------
pad8(x...) = ntuple(i -> i <= length(x) ? Float64(x[i]) : 0.0, 8)
matdesc(m::Hooke)                   = (MAT_HOOKE,        Int32(0), m.small ? F_SMALL : UInt32(0), (m.E, m.ν))
matdesc(m::Hooke2D{T,P}) where {T,P} = (MAT_HOOKE2D,      Int32(0), (m.small ? F_SMALL : UInt32(0)) | (P == :plane_stress ? F_PLANE_STRESS : UInt32(0)), (m.E, m.ν))
matdesc(m::NeoHooke)                = (MAT_NEOHOOKE,     Int32(0), UInt32(0), (m.C1, m.K))
matdesc(m::MooneyRivlin)            = (MAT_MOONEYRIVLIN, Int32(0), UInt32(0), (m.C1, m.C2, m.K))
matdesc(m::Ogden)                   = (MAT_OGDEN,        Int32(0), UInt32(0), (m.α, m.μ, m.K))
function matdesc(m::PhaseField{M,AT}; intended_grad=false, nh2d=false) where {M,AT}
  kind, _, flags, p = matdesc(m.mat)
  flags |= (AT == :DHn ? F_DHN : UInt32(0)) | (intended_grad ? F_PF_GRAD_INTENDED : UInt32(0)) | (nh2d ? F_NH2D_EXT : UInt32(0))
  (MAT_PHASEFIELD, kind, flags, (ntuple(i -> i <= length(p) ? Float64(p[i]) : 0.0, 4)..., m.l0, m.Gc, m.n))
end

# ---- SoA packing of a homogeneous run of elements (elements.jl:71-77, 104-111) ------------------------
struct Packed
  batch::Batch
  keep::Vector{Any}              # arrays the Batch points into
end
function pack(elems::AbstractVector{E}, idx::Vector{Int}; kw...) where {D,P,M,T,N,E<:AbstractCElem{D,P,M,T,N}}
  ne = length(idx)
  nodes = Matrix{Int64}(undef, N, ne)                 # [ne][N] row-major == N x ne column-major
  gradN = Array{Float64}(undef, N, P, D, ne)          # [ne][D][P][N]
  wgt   = Matrix{Float64}(undef, P, ne)
  iscp  = E <: CPElem
  Nval  = iscp ? Array{Float64}(undef, N, P, ne) : Float64[]
  for (k, ii) in enumerate(idx)
    e = elems[ii]
    nodes[:, k] .= e.nodes
    for d in 1:D, p in 1:P
      gradN[:, p, d, k] .= e.∇N[d][p]
      iscp && d == 1 && (Nval[:, p, k] .= e.N[p])
    end
    wgt[:, k] .= e.wgt
  end
  kind, base, flags, params = elems[idx[1]].mat isa PhaseField ? matdesc(elems[idx[1]].mat; kw...) : matdesc(elems[idx[1]].mat)
  oi = Int64.(idx .- 1)
  b = Batch(iscp ? ELEM_CP : ELEM_CE, D, P, N, kind, base, flags, 0, pad8(params...), ne,
            pointer(nodes), pointer(gradN), pointer(wgt), iscp ? pointer(Nval) : C_NULL, pointer(oi), C_NULL)
  Packed(b, Any[nodes, gradN, wgt, Nval, oi])
end

# ---- CAS: the axisymmetric element the reference constructs but never defines (elasticelements.jl:251, 284 call
# `CAS(nodes, N, Nx, Ny, X0, wgt, V, mat)`; no struct, no getϕ -- SURVEY §9-7).  Defining it inside AD4SM.Elements makes
# `ASTria` / `ASQuad` work as written; its energy exists only in this backend (3x3 F with the hoop stretch 1 + (N·u_r)/X0
# and the 3-D materials; an extension, see include/ad4sm_b200.h AD4SM_ELEM_CAS).
@eval AD4SM.Elements begin
  struct CAS{P,M,T,N} <: AbstractCElem{2,P,M,T,N}
    nodes::Vector{Int}
    N::NTuple{P,Vector{T}}
    Nx::NTuple{P,Vector{T}}
    Ny::NTuple{P,Vector{T}}
    X0::NTuple{P,T}
    wgt::NTuple{P,T}
    V::T
    mat::M
  end
  CAS(nodes, N, Nx, Ny, X0, wgt, V, mat::M) where M =
    CAS{length(wgt),M,typeof(float(V)),length(nodes)}(collect(Int, nodes), Tuple(N), Tuple(Nx), Tuple(Ny), Tuple(X0), Tuple(wgt), float(V), mat)
end
const CAS = AD4SM.Elements.CAS
function pack(elems::AbstractVector{E}, idx::Vector{Int}; kw...) where {P,M,T,N,E<:CAS{P,M,T,N}}
  ne = length(idx)
  nodes = Matrix{Int64}(undef, N, ne)
  gradN = Array{Float64}(undef, N, P, 2, ne)          # [ne][2][P][N] = (Nx, Ny)
  wgt   = Matrix{Float64}(undef, P, ne)
  hoop  = Array{Float64}(undef, N, P, ne)             # the C ABI's Nval for CAS: N[gp][a] / X0[gp]
  for (k, ii) in enumerate(idx)
    e = elems[ii]
    nodes[:, k] .= e.nodes
    for p in 1:P
      gradN[:, p, 1, k] .= e.Nx[p]
      gradN[:, p, 2, k] .= e.Ny[p]
      hoop[:, p, k]     .= e.N[p] ./ e.X0[p]
    end
    wgt[:, k] .= e.wgt
  end
  kind, base, flags, params = matdesc(elems[idx[1]].mat)
  oi = Int64.(idx .- 1)
  b = Batch(ELEM_CAS, 2, P, N, kind, base, flags, 0, pad8(params...), ne,
            pointer(nodes), pointer(gradN), pointer(wgt), pointer(hoop), pointer(oi), C_NULL)
  Packed(b, Any[nodes, gradN, wgt, hoop, oi])
end

# ---- batched constructors on the device (ad4sm_construct; elasticelements.jl:35-197, phasefieldelements.jl:9-145) ---------
# `construct(shape, conn, p0; GP, phasefield)`: what `[Hex08(conn[:,e], p0[conn[:,e]]; GP) for e in ...]` computes, for all
# elements in two kernels.  conn: N x ne node ids (1-based), p0: dim x nNodes coordinates.  Returns (∇N[N,P,D,ne], wgt[P,ne],
# N[N,P,ne] or nothing, V[ne]); `elements(...)` wraps them into the reference's element structs.
const GP2 = ((-0.577350269189626, 1.0), (0.577350269189626, 1.0))
shape_dims(shape, n) = shape == SHAPE_TRIA3 ? (2, 3, 1) : shape == SHAPE_QUAD4 ? (2, 4, n^2) : shape == SHAPE_TET4 ? (3, 4, 1) :
                       shape == SHAPE_HEX8 ? (3, 8, n^3) : (3, 6, 6)
function construct(shape::Int32, conn::Matrix{Int64}, p0::Matrix{Float64}; GP=GP2, phasefield=false)
  n = shape in (SHAPE_QUAD4, SHAPE_HEX8) ? length(GP) : 0
  D, N, P = shape_dims(shape, n)
  size(conn, 1) == N || error("connectivity must be $N x ne")
  ne = size(conn, 2)
  xi = ntuple(i -> i <= n ? Float64(GP[i][1]) : 0.0, 3); w = ntuple(i -> i <= n ? Float64(GP[i][2]) : 0.0, 3)
  gradN = Array{Float64}(undef, N, P, D, ne); wgt = Matrix{Float64}(undef, P, ne); V = Vector{Float64}(undef, ne)
  Nval = phasefield ? Array{Float64}(undef, N, P, ne) : nothing
  GC.@preserve p0 begin
    m = Ref(Mesh(shape, n, xi, w, size(p0, 1), 0, size(p0, 2), pointer(p0)))
    check(ccall((:ad4sm_construct, lib), Cint,
                (Ref{Mesh}, Int64, Ptr{Int64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                m, ne, conn, -1, gradN, wgt, Nval === nothing ? C_NULL : Nval, V))
  end
  (gradN, wgt, Nval, V)
end

# ---- plan cache: one plan per (elems object, size(u)) ---------------------------------------------------
# Keyed WEAKLY on the `elems` vector (a strong key would keep every mesh -- and ~12 GB of HBM per 1M-element plan --
# alive for ever: the finalizer could never run).  A hit is validated against length(elems) and size(u), so a vector that
# was push!-ed to, or is evaluated with another size(u,1), gets a fresh plan instead of a stale one.  Mutating an
# element or a material IN PLACE is not detectable (the reference's element structs are immutable; `elems[i] = other`
# on the same vector is the one case the caller must announce with `release!(elems)`).
mutable struct Plan
  h::Ptr{Cvoid}
  ne::Int
  usize::Tuple{Int,Int}
  groups::Vector{Vector{Int}}      # per batch: positions in `elems` (history marshalling)
  gauss::Vector{Int}               # per batch: Gauss points per element
end
const plans = WeakKeyDict{Any,Plan}()

destroy!(p::Plan) = (p.h != C_NULL && ccall((:ad4sm_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), p.h); p.h = C_NULL; nothing)
"""release!(elems): drop (and free on the device) the plan cached for `elems`."""
release!(elems) = (haskey(plans, elems) && destroy!(pop!(plans, elems)); nothing)

function plan_for(elems, u; kw...)
  hit = get(plans, elems, nothing)
  if hit !== nothing
    hit.ne == length(elems) && hit.usize == size(u) && return hit
    release!(elems)
  end
  groups = Dict{Any,Vector{Int}}()                     # one batch per (element type, material)
  order = Any[]
  for (i, e) in enumerate(elems)
    k = (typeof(e), e.mat)
    haskey(groups, k) || push!(order, k)
    push!(get!(groups, k, Int[]), i)
  end
  idxs = [groups[k] for k in order]                    # deterministic batch order: first appearance in `elems`
  packed = [pack(elems, idx; kw...) for idx in idxs]
  batches = [p.batch for p in packed]
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve packed batches begin
    check(ccall((:ad4sm_plan_create, lib), Cint, (Ptr{Batch}, Int32, Int64, Int32, Int32, Ref{Ptr{Cvoid}}),
                batches, length(batches), size(u, 2), size(u, 1), -1, h))
  end
  p = Plan(h[], length(elems), size(u), idxs, [Int(b.n_gauss) for b in batches])
  finalizer(destroy!, p)
  plans[elems] = p
  p
end

dense(u::Matrix{Float64}) = u                          # already what the C ABI reads: no copy
dense(u::AbstractMatrix)  = Matrix{Float64}(u)

function fetch(p::Plan, n, nnz)
  colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnz); nzval = Vector{Float64}(undef, nnz)
  check(ccall((:ad4sm_fetch_csc, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), p.h, colptr, rowval, nzval))
  SparseMatrixCSC{Float64,Int64}(n, n, colptr, rowval, nzval)   # already sorted, duplicates summed, zeros dropped
end

# ---- the overridden methods -----------------------------------------------------------------------------
function AD4SM.Elements.makeϕrKt(elems::AbstractVector{<:AbstractCElem}, u::AbstractMatrix{Float64})
  @assert length(elems) > 0                                        # elasticelements.jl:210
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(u))
  uc = dense(u)                                                    # dense column-major D x nNodes (no copy for a Matrix{Float64})
  GC.@preserve uc check(ccall((:ad4sm_eval, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ref{Int64}), p.h, uc, ϕ, r, nnz))
  (ϕ[], r, fetch(p, length(u), nnz[]))
end

# Assembly only (elements.toolkit.jl:169-203; the method Solvers.solvestep! calls, solvers.jl:213-216): a
# Vector{D2{N,M,Float64}} is a dense array of (v | g | h) records, passed as it lies in memory.
function AD4SM.Elements.makeϕrKt(Φ::Vector{adiff.D2{N,M,Float64}}, elems::Vector, u::AbstractMatrix{Float64}) where {N,M}
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(u))
  GC.@preserve Φ check(ccall((:ad4sm_assemble_duals, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Float64}, Ptr{Float64}, Ref{Int64}),
                             p.h, pointer(Φ), 1 + N + M, ϕ, r, nnz))
  (ϕ[], r, fetch(p, length(u), nnz[]))
end

function AD4SM.Elements.makeϕr(Φ::Vector{T}, elems::Vector, u::AbstractMatrix{Float64}) where {T<:Union{adiff.D1,adiff.D2}}   # toolkit:204-215
  p = plan_for(elems, u); ϕ = Ref(0.0); r = Vector{Float64}(undef, length(u))
  GC.@preserve Φ check(ccall((:ad4sm_assemble_duals_phi_r, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Float64}, Ptr{Float64}),
                             p.h, pointer(Φ), sizeof(T) ÷ 8, ϕ, r))
  (ϕ[], r)
end

function AD4SM.Elements.makeϕrKt(elems::Vector{<:CPElem}, u::Array{Float64}, d::Array{Float64})
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(u))
  check(ccall((:ad4sm_eval_pf_u, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ref{Int64}),
              p.h, u, vec(d), ϕ, r, nnz))
  (ϕ[], r, fetch(p, length(u), nnz[]))
end

function AD4SM.Elements.makeϕrKt_d(elems::Vector{<:CPElem}, u::Array{Float64}, d::Array{Float64})
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(d))
  check(ccall((:ad4sm_eval_pf_d, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ref{Float64}, Ptr{Float64}, Ref{Int64}),
              p.h, u, vec(d), C_NULL, ϕ, r, nnz))
  (ϕ[], r, fetch(p, length(d), nnz[]))
end

# ϕmax[ii][gp] is mutated in place by the reference (phasefieldelements.jl:180): a Tuple (ϕmax⁺, ϕmax⁻) for the Hooke-type
# bases (phasefieldmaterials.jl:75-93, 116-135), a scalar for PhaseField{NeoHooke} (:94-115, the second slot is unused).
# Marshalled as one [ne_b][P][2] array per batch -- the library reads hist[b] for EVERY batch -- and scattered back through
# the batch's positions in `elems`.
hist_get(x::Tuple) = (Float64(x[1]), Float64(x[2]))
hist_get(x::Real)  = (Float64(x), 0.0)
hist_set(old::Tuple, a, b) = (a, b)
hist_set(old::Real, a, b)  = a
function AD4SM.Elements.makeϕrKt_d(elems::Vector{<:CPElem}, u::Array{Float64}, d::Array{Float64}, ϕmax::Array)
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(d))
  hs = [Array{Float64}(undef, 2, p.gauss[b], length(idx)) for (b, idx) in enumerate(p.groups)]
  for (b, idx) in enumerate(p.groups), (k, e) in enumerate(idx), gp in 1:p.gauss[b]
    hs[b][1, gp, k], hs[b][2, gp, k] = hist_get(ϕmax[e][gp])
  end
  hp = [pointer(h) for h in hs]
  GC.@preserve hs check(ccall((:ad4sm_eval_pf_d, lib), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ref{Float64}, Ptr{Float64}, Ref{Int64}), p.h, u, vec(d), hp, ϕ, r, nnz))
  for (b, idx) in enumerate(p.groups), (k, e) in enumerate(idx), gp in 1:p.gauss[b]
    ϕmax[e][gp] = hist_set(ϕmax[e][gp], hs[b][1, gp, k], hs[b][2, gp, k])
  end
  (ϕ[], r, fetch(p, length(d), nnz[]))
end

# ---- Newton-step glue on the device (solvers.jl:183-184, 211-224, 257-291; include/ad4sm_b200.h ad4sm_newton_*) --------
# In `Solvers.solvestep!` replace, for nEqs == 0,
#     res = fi[ifreeu]-fe[ifreeu];  res .-= Kt[ifreeu,icnstu]*Δucnst;  updt = lu(Kt[ifreeu,ifreeu])\res          (:220-223)
# by
#     makeϕrKt!(elems, uold);  res, Kff = newton_reduce(elems, uold, bfreeu, 0; fe=fe, Δucnst=Δucnst);  updt = lu(Kff)\res
# so that the full Kt never crosses PCIe and no sparse slicing runs on the host.
struct NewtonSplit
  nfree::Int; ncnst::Int; colptr::Vector{Int64}; rowval::Vector{Int64}
end
const splits = WeakKeyDict{Any,Tuple{BitVector,NewtonSplit}}()
function newton_split(elems, u, bfree)
  p = plan_for(elems, u); bf = BitVector(vec(bfree))
  hit = get(splits, elems, nothing)
  hit !== nothing && hit[1] == bf && return hit[2]
  nf = Ref{Int64}(0); nc = Ref{Int64}(0); nz = Ref{Int64}(0); b8 = UInt8.(bf)
  check(ccall((:ad4sm_newton_setup, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ref{Int64}, Ref{Int64}, Ref{Int64}), p.h, 0, b8, nf, nc, nz))
  cp = Vector{Int64}(undef, nf[] + 1); rv = Vector{Int64}(undef, nz[])
  check(ccall((:ad4sm_newton_pattern, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}), p.h, 0, cp, rv))
  s = NewtonSplit(nf[], nc[], cp, rv); splits[elems] = (bf, s); s
end
"""evaluate on the device without fetching Kt: ϕ, r"""
function makeϕrKt!(elems, u::Matrix{Float64})
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(u))
  check(ccall((:ad4sm_eval, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ref{Int64}), p.h, u, ϕ, r, nnz))
  (ϕ[], r)
end
"""kind 0: predictor residual (:220-221), 1: corrector residual (:262).  Returns (res, Kt[ifree,ifree], norm(res), norm(fi[icnst]))."""
function newton_reduce(elems, u, bfree, kind; fe=nothing, Δucnst=nothing)
  p = plan_for(elems, u); s = newton_split(elems, u, bfree)
  res = Vector{Float64}(undef, s.nfree); nz = Vector{Float64}(undef, length(s.rowval)); norms = zeros(2)
  check(ccall((:ad4sm_newton_reduce, lib), Cint,
              (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              p.h, 0, kind, fe === nothing ? C_NULL : pointer(fe), Δucnst === nothing ? C_NULL : pointer(Δucnst), res, nz, norms))
  (res, SparseMatrixCSC{Float64,Int64}(s.nfree, s.nfree, copy(s.colptr), copy(s.rowval), nz), norms[1], norms[2])
end

# ---- mesh-partitioned evaluation, one Julia process per GPU (ad4sm_dist_*: NCCL inside the library) ----------------------
# The reference has no multi-device path.  Every rank holds its own elements `own` (LOCAL node ids) and the connectivity
# `halo_nodes` (N x nHalo, local ids) of the foreign elements that touch nodes it owns; `peers` lists, per neighbour, the
# own slots (0-based) whose staged rows that neighbour needs and the range of HALO slots it fills (see
# include/ad4sm_b200.h: ad4sm_dist_peer; ad4sm.jl_b200/dist.py: build_partition computes exactly these from a global
# connectivity and an element -> rank map).  `bcast(bytes_or_nothing)` must hand rank 0's 128-byte NCCL id to every rank
# (MPI.bcast, a shared file, Distributed.remotecall_fetch, ...).
struct DistPeer                    # struct ad4sm_dist_peer
  rank::Int32; n_send::Int64; send_slots::Ptr{Int64}; recv_first_slot::Int64; n_recv::Int64
end
mutable struct DistPlan
  plan::Plan
  comm::Ptr{Cvoid}
  keep::Vector{Any}
end
function dist_plan(own::AbstractVector{<:AbstractCElem}, halo_nodes::Matrix{Int64}, own_index::Vector{Int64}, halo_index::Vector{Int64},
                   n_local_nodes::Int, dofs_per_node::Int, nranks::Int, rank::Int, device::Int,
                   peers::Vector{Tuple{Int,Vector{Int64},Int,Int}}, bcast)
  packed = pack(own, collect(1:length(own)))
  oi = copy(own_index)                                 # global positions in `elems`: fixes the summation order across ranks
  b  = packed.batch
  ob = Batch(b.elem_kind, b.dim, b.n_gauss, b.n_elem_nodes, b.mat_kind, b.mat_base, b.mat_flags, 0, b.mat_params, b.n_elems,
             b.nodes, b.gradN, b.wgt, b.Nval, pointer(oi), C_NULL)
  hb = Batch(b.elem_kind, b.dim, b.n_gauss, b.n_elem_nodes, b.mat_kind, b.mat_base, b.mat_flags, 1 #= AD4SM_BATCH_HALO =#, b.mat_params,
             size(halo_nodes, 2), pointer(halo_nodes), C_NULL, C_NULL, C_NULL, pointer(halo_index), C_NULL)
  batches = size(halo_nodes, 2) > 0 ? [ob, hb] : [ob]
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve packed oi halo_nodes halo_index batches begin
    check(ccall((:ad4sm_plan_create, lib), Cint, (Ptr{Batch}, Int32, Int64, Int32, Int32, Ref{Ptr{Cvoid}}),
                batches, length(batches), n_local_nodes, dofs_per_node, device, h))
  end
  plan = Plan(h[], length(own), (dofs_per_node, n_local_nodes), [collect(1:length(own))], [Int(b.n_gauss)])
  finalizer(destroy!, plan)
  id = Vector{UInt8}(undef, 128)
  rank == 0 && check(ccall((:ad4sm_dist_unique_id, lib), Cint, (Ptr{UInt8},), id))
  id = bcast(rank == 0 ? id : nothing)
  comm = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ad4sm_dist_init, lib), Cint, (Int32, Int32, Ptr{UInt8}, Int32, Ref{Ptr{Cvoid}}), nranks, rank, id, device, comm))
  n_own = length(own)
  dp = [DistPeer(pr, length(send), pointer(send), n_own + first, cnt) for (pr, send, first, cnt) in peers]
  GC.@preserve peers dp check(ccall((:ad4sm_dist_attach, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{DistPeer}), plan.h, comm[], length(dp), dp))
  DistPlan(plan, comm[], Any[packed, oi, halo_nodes, halo_index, peers])
end
"""One partitioned evaluation: `u` is this rank's D x nLocalNodes block.  Returns the GLOBAL ϕ and this rank's r; the rows
of Kt this rank owns stay on the device (`fetch(dp.plan, ...)` / `newton_reduce` read them) and are bitwise those of a
single-GPU evaluation of the whole mesh."""
function dist_makeϕrKt(dp::DistPlan, u::Matrix{Float64})
  ϕ = Ref(0.0); r = Vector{Float64}(undef, length(u))
  check(ccall((:ad4sm_dist_eval, lib), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
              dp.plan.h, 0, u, C_NULL, ϕ, r))
  (ϕ[], r)
end
dist_close!(dp::DistPlan) = (dp.comm != C_NULL && ccall((:ad4sm_dist_destroy, lib), Cvoid, (Ptr{Cvoid},), dp.comm); dp.comm = C_NULL; nothing)

end # module
