#!/usr/bin/env python
"""Writes the INPUTS of the Julia-side golden cases: tests/golden/julia/<case>/{meta.txt, coords.f64, conn.i64, u.f64}
(raw little-endian, column-major like Julia reads them).  Someone with Julia then runs, per case,

    julia -t auto bench/ref_julia.jl --dump tests/golden/julia/<case>

which evaluates the UNMODIFIED reference `makeϕrKt(elems, u)` on exactly these bits and writes phi.f64, r.f64,
colptr.i64, rowval.i64, nzval.i64 next to them.  tests/test_julia_golden.py picks the outputs up when they exist and
compares both the oracle (CPU) and the CUDA path (GPU) with them -- the one route from "parity unpinned" to pinned."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ad4sm_b200 import mesh  # noqa: E402

CASES = {
    # name: (kind, dims, material line, seed)
    "hex8_neohooke_6x5x4": ("hex8", (6, 5, 4), "neohooke 1000.0 5000.0", 101),
    "hex8_mooney_4x4x3": ("hex8", (4, 4, 3), "mooneyrivlin 800.0 200.0 5000.0", 102),
    "tet4_mooney_4x3x3": ("tet4", (4, 3, 3), "mooneyrivlin 800.0 200.0 5000.0", 103),
    "quad4_hooke2d_pe_12x9": ("quad4", (12, 9), "hooke2d 210000.0 0.3 plane_strain small", 104),
    "hex8_neohooke_56x56x56": ("hex8", (56, 56, 56), "neohooke 1000.0 5000.0", 21),   # = tests/test_gpu_fullsize C2@56^3 inputs
}


def build(kind, dims, seed):
    gen = {"hex8": mesh.hex_block, "tet4": mesh.tet_kuhn, "quad4": mesh.quad_grid}[kind]
    coords, conn = gen(*dims, jitter=0.1, seed=seed)
    D = coords.shape[1]
    u = mesh.random_u(len(coords), D, 1.0 / max(dims), a=0.02, seed=seed + 100)
    return coords, conn, u


def main(only=None):
    for name, (kind, dims, mat, seed) in CASES.items():
        if only and name not in only:
            continue
        if name.endswith("56x56x56") and not only:
            continue   # 40 MB of inputs: generated on request only (python make_julia_inputs.py hex8_neohooke_56x56x56)
        coords, conn, u = build(kind, dims, seed)
        d = os.path.join(HERE, "julia", name)
        os.makedirs(d, exist_ok=True)
        np.ascontiguousarray(coords.T.reshape(-1, order="F")).astype("<f8").tofile(os.path.join(d, "coords.f64"))   # D x nNodes column-major
        np.ascontiguousarray(conn.T.reshape(-1, order="F")).astype("<i8").tofile(os.path.join(d, "conn.i64"))      # N x ne column-major
        np.asfortranarray(u).reshape(-1, order="F").astype("<f8").tofile(os.path.join(d, "u.f64"))                  # D x nNodes column-major
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"kind {kind}\ndim {coords.shape[1]}\nnodes {len(coords)}\nelems {len(conn)}\nnodes_per_elem {conn.shape[1]}\n"
                    f"material {mat}\ndims {' '.join(map(str, dims))}\nseed {seed}\n")
        print(name, "->", d)


if __name__ == "__main__":
    main(sys.argv[1:])
