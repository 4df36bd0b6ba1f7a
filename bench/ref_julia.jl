# ref_julia.jl -- times the UNMODIFIED reference `makeϕrKt(elems, u)` on config C2 (Hex8 NeoHooke n^3 block).
# Shipped, not run: Julia is not installed in the build container (BASELINE.md §4).  The mesh, material and u follow
# bench.py's generator (ad4sm.jl_b200/mesh.py: nodes x-fastest, h = 1/n, u = 0.02 h U(-1,1)); for a bit-identical
# comparison write u from Python (`np.asfortranarray(u).tofile("u.f64")`) and pass the file as the second argument.
#
#     julia -t auto bench/ref_julia.jl 100 [u.f64]
using AD4SM, Random, Printf

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
u = length(ARGS) >= 2 ? reshape(reinterpret(Float64, read(ARGS[2])), 3, :) |> collect :
                        0.02 * h * (2 .* rand(MersenneTwister(20261023), 3, length(nodes)) .- 1)
makeϕrKt(elems[1:min(end, 64)], u)                      # compile
ts = [(@elapsed makeϕrKt(elems, u)) for _ in 1:3]
@printf("{\"impl\": \"reference-julia\", \"threads\": %d, \"elements\": %d, \"value\": %.6e, \"unit\": \"elements/s\", \"best_s\": %.4f}\n",
        Threads.nthreads(), length(elems), length(elems) / minimum(ts), minimum(ts))
