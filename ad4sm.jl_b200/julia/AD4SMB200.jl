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
const ELEM_CE, ELEM_CP = Int32(0), Int32(1)
const MAT_HOOKE, MAT_HOOKE2D, MAT_NEOHOOKE, MAT_MOONEYRIVLIN, MAT_OGDEN, MAT_PHASEFIELD = Int32.(1:6)
const F_SMALL, F_PLANE_STRESS, F_DHN, F_PF_GRAD_INTENDED, F_NH2D_EXT = UInt32(1), UInt32(2), UInt32(4), UInt32(8), UInt32(16)

struct Batch                      # struct ad4sm_batch
  elem_kind::Int32; dim::Int32; n_gauss::Int32; n_elem_nodes::Int32
  mat_kind::Int32; mat_base::Int32; mat_flags::UInt32; batch_flags::Int32
  mat_params::NTuple{8,Float64}
  n_elems::Int64
  nodes::Ptr{Int64}; gradN::Ptr{Float64}; wgt::Ptr{Float64}; Nval::Ptr{Float64}; orig_index::Ptr{Int64}
end

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
            pointer(nodes), pointer(gradN), pointer(wgt), iscp ? pointer(Nval) : C_NULL, pointer(oi))
  Packed(b, Any[nodes, gradN, wgt, Nval, oi])
end

# ---- plan cache: one plan per (elems object, size(u)) ---------------------------------------------------
mutable struct Plan
  h::Ptr{Cvoid}
  ndof::Int
  nnodes::Int
end
const plans = IdDict{Any,Plan}()

function plan_for(elems, u; kw...)
  get!(plans, elems) do
    groups = Dict{Any,Vector{Int}}()                     # one batch per (element type, material)
    for (i, e) in enumerate(elems)
      push!(get!(groups, (typeof(e), e.mat), Int[]), i)
    end
    packed = [pack(elems, idx; kw...) for idx in values(groups)]
    batches = [p.batch for p in packed]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve packed batches begin
      check(ccall((:ad4sm_plan_create, lib), Cint, (Ptr{Batch}, Int32, Int64, Int32, Int32, Ref{Ptr{Cvoid}}),
                  batches, length(batches), size(u, 2), size(u, 1), -1, h))
    end
    p = Plan(h[], length(u), size(u, 2))
    finalizer(q -> ccall((:ad4sm_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), q.h), p)
    p
  end
end

function fetch(p::Plan, n, nnz)
  colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nnz); nzval = Vector{Float64}(undef, nnz)
  check(ccall((:ad4sm_fetch_csc, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), p.h, colptr, rowval, nzval))
  SparseMatrixCSC{Float64,Int64}(n, n, colptr, rowval, nzval)   # already sorted, duplicates summed, zeros dropped
end

# ---- the overridden methods -----------------------------------------------------------------------------
function AD4SM.Elements.makeϕrKt(elems::AbstractVector{<:AbstractCElem}, u::AbstractMatrix{Float64})
  @assert length(elems) > 0                                        # elasticelements.jl:210
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(u))
  uc = Matrix{Float64}(u)                                          # dense column-major D x nNodes
  check(ccall((:ad4sm_eval, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ref{Int64}), p.h, uc, ϕ, r, nnz))
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

# ϕmax[ii][gp] is a Tuple (ϕmax⁺, ϕmax⁻) mutated in place by the reference (phasefieldelements.jl:180): marshal it as
# one [ne][P][2] array per batch (single homogeneous batch shown; mixed meshes scatter back by `orig_index`).
function AD4SM.Elements.makeϕrKt_d(elems::Vector{<:CPElem}, u::Array{Float64}, d::Array{Float64}, ϕmax::Array)
  p = plan_for(elems, u); ϕ = Ref(0.0); nnz = Ref{Int64}(0); r = Vector{Float64}(undef, length(d))
  P = length(ϕmax[1])
  h = Array{Float64}(undef, 2, P, length(elems))
  for (e, he) in enumerate(ϕmax), gp in 1:P
    h[1, gp, e], h[2, gp, e] = he[gp]
  end
  hp = [pointer(h)]
  GC.@preserve h check(ccall((:ad4sm_eval_pf_d, lib), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ref{Float64}, Ptr{Float64}, Ref{Int64}), p.h, u, vec(d), hp, ϕ, r, nnz))
  for e in eachindex(ϕmax), gp in 1:P
    ϕmax[e][gp] = (h[1, gp, e], h[2, gp, e])
  end
  (ϕ[], r, fetch(p, length(d), nnz[]))
end

end # module
