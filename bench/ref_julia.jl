# ref_julia.jl -- the UNMODIFIED reference `makeϕrKt(elems, u)`, timed or dumped.
# Shipped, not run: Julia is not installed in the build container (BASELINE.md §4).
#
#   julia -t auto bench/ref_julia.jl 100 [u.f64]             time config C2 (Hex8 NeoHooke n^3 block; mesh, material and u
#                                                            follow bench.py's generator, ad4sm.jl_b200/mesh.py)
#   julia -t auto bench/ref_julia.jl --dump <case dir>       read meta.txt / coords.f64 / conn.i64 / u.f64 written by
#                                                            tests/golden/make_julia_inputs.py, evaluate makeϕrKt on exactly
#                                                            those bits and write phi.f64, r.f64, colptr.i64, rowval.i64,
#                                                            nzval.f64 (raw little-endian) next to them: the golden vectors
#                                                            tests/test_julia_golden.py consumes
using AD4SM, Random, Printf, SparseArrays

readvec(T, path) = collect(reinterpret(T, read(path)))

function material(words)
  k = words[1]; x = [tryparse(Float64, w) for w in words[2:end]]; v = Float64[y for y in x if y !== nothing]
  k == "neohooke"     && return Materials.NeoHooke(v[1], v[2])
  k == "mooneyrivlin" && return Materials.MooneyRivlin(v[1], v[2], v[3])
  k == "hooke"        && return Materials.Hooke(v[1], v[2]; small = ("small" in words))
  k == "hooke2d"      && return Materials.Hooke2D(v[1], v[2]; small = ("small" in words), plane_stress = ("plane_stress" in words))
  error("unknown material line: " * join(words, " "))
end

function dump_case(dir)
  meta = Dict{String,Vector{String}}()
  for l in eachline(joinpath(dir, "meta.txt"))
    w = split(strip(l)); isempty(w) || (meta[String(w[1])] = String.(w[2:end]))
  end
  D = parse(Int, meta["dim"][1]); nn = parse(Int, meta["nodes"][1]); ne = parse(Int, meta["elems"][1]); N = parse(Int, meta["nodes_per_elem"][1])
  coords = reshape(readvec(Float64, joinpath(dir, "coords.f64")), D, nn)
  conn   = reshape(readvec(Int64, joinpath(dir, "conn.i64")), N, ne)
  u      = reshape(readvec(Float64, joinpath(dir, "u.f64")), D, nn)
  mat    = material(meta["material"])
  ctor   = Dict("hex8" => Elements.Hex08, "tet4" => Elements.Tet04, "quad4" => Elements.Quad)[meta["kind"][1]]
  elems  = [begin
              nd = conn[:, e]
              ctor(nd, [coords[:, i] for i in nd]; mat = mat)      # the reference's own constructors (elasticelements.jl:35-162)
            end for e in 1:ne]
  ϕ, r, Kt = makeϕrKt(elems, u)
  write(joinpath(dir, "phi.f64"), Float64[ϕ]); write(joinpath(dir, "r.f64"), Vector{Float64}(r))
  write(joinpath(dir, "colptr.i64"), Vector{Int64}(Kt.colptr)); write(joinpath(dir, "rowval.i64"), Vector{Int64}(Kt.rowval))
  write(joinpath(dir, "nzval.f64"), Vector{Float64}(Kt.nzval))
  @printf("%s: ne=%d nnz=%d phi=%.15e (julia %s, %d threads)\n", dir, ne, nnz(Kt), ϕ, VERSION, Threads.nthreads())
end

if length(ARGS) >= 2 && ARGS[1] == "--dump"
  foreach(dump_case, ARGS[2:end])
else
  n  = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 40
  h  = 1.0 / n
  id(i, j, k) = 1 + i + (n + 1) * (j + (n + 1) * k)
  nodes = [[Float64(i) * h, Float64(j) * h, Float64(k) * h] for k in 0:n for j in 0:n for i in 0:n]
  mat   = Materials.NeoHooke(1000.0, 5000.0)
  elems = [begin
             nd = [id(i, j, k), id(i + 1, j, k), id(i + 1, j + 1, k), id(i, j + 1, k),
                   id(i, j, k + 1), id(i + 1, j, k + 1), id(i + 1, j + 1, k + 1), id(i, j + 1, k + 1)]
             Elements.Hex08(nd, nodes[nd]; mat = mat)
           end for k in 0:n-1 for j in 0:n-1 for i in 0:n-1]
  u = length(ARGS) >= 2 ? reshape(readvec(Float64, ARGS[2]), 3, :) :
                          0.02 * h * (2 .* rand(MersenneTwister(20261023), 3, length(nodes)) .- 1)
  makeϕrKt(elems[1:min(end, 64)], u)                      # compile
  ts = [(@elapsed makeϕrKt(elems, u)) for _ in 1:3]
  @printf("{\"impl\": \"reference-julia\", \"threads\": %d, \"elements\": %d, \"value\": %.6e, \"unit\": \"elements/s\", \"best_s\": %.4f}\n",
          Threads.nthreads(), length(elems), length(elems) / minimum(ts), minimum(ts))
end
