"""Golden vectors produced by the REAL reference (Julia), when someone has dropped them into tests/golden/julia/<case>/
(see the README there; `julia bench/ref_julia.jl --dump <case>` writes them).  None can be produced in the build image --
no Julia -- so these tests skip until the files exist; with them, the oracle (CPU) and the CUDA path (GPU) are both held to
the reference's own output: pattern bit-exact, every stored value to relative 1e-12."""
import glob
import os

import numpy as np
import pytest

from ad4sm_b200 import Elements as El, Materials as Ma

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(d for d in glob.glob(os.path.join(HERE, "golden", "julia", "*")) if os.path.isfile(os.path.join(d, "meta.txt")))
OUT = ("phi.f64", "r.f64", "colptr.i64", "rowval.i64", "nzval.f64")


def _load(d):
    meta = {}
    for line in open(os.path.join(d, "meta.txt")):
        w = line.split()
        if w:
            meta[w[0]] = w[1:]
    D, nn, ne, N = (int(meta[k][0]) for k in ("dim", "nodes", "elems", "nodes_per_elem"))
    coords = np.fromfile(os.path.join(d, "coords.f64"), "<f8").reshape((D, nn), order="F").T.copy()
    conn = np.fromfile(os.path.join(d, "conn.i64"), "<i8").reshape((N, ne), order="F").T.copy()
    u = np.fromfile(os.path.join(d, "u.f64"), "<f8").reshape((D, nn), order="F")
    m = meta["material"]
    v = [float(x) for x in m[1:] if x[0].isdigit() or x[0] == "-"]
    mat = {"neohooke": lambda: Ma.NeoHooke(v[0], v[1]), "mooneyrivlin": lambda: Ma.MooneyRivlin(v[0], v[1], v[2]),
           "hooke": lambda: Ma.Hooke(v[0], v[1], small="small" in m),
           "hooke2d": lambda: Ma.Hooke2D(v[0], v[1], small="small" in m, plane_stress="plane_stress" in m)}[m[0]]()
    ctor = {"hex8": El.Hex08_batch, "tet4": El.Tet04_batch, "quad4": El.Quad_batch}[meta["kind"][0]]
    return ctor(conn, coords, mat), np.asfortranarray(u)


def _julia(d):
    if not all(os.path.isfile(os.path.join(d, f)) for f in OUT):
        pytest.skip("no Julia-produced outputs in this case directory (tests/golden/julia/README.md)")
    f = lambda n, t: np.fromfile(os.path.join(d, n), t)  # noqa: E731
    return float(f("phi.f64", "<f8")[0]), f("r.f64", "<f8"), (f("colptr.i64", "<i8"), f("rowval.i64", "<i8"), f("nzval.f64", "<f8"))


def test_inputs_are_what_the_generator_writes():
    """the committed inputs are reproducible from tests/golden/make_julia_inputs.py (same bits)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mji", os.path.join(HERE, "golden", "make_julia_inputs.py"))
    mji = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mji)
    assert CASES, "tests/golden/julia holds no case"
    for d in CASES:
        kind, dims, _, seed = mji.CASES[os.path.basename(d)]
        coords, conn, u = mji.build(kind, dims, seed)
        b, u2 = _load(d)
        assert np.array_equal(u2, u) and np.array_equal(b.nodes, conn)


@pytest.mark.parametrize("d", CASES, ids=[os.path.basename(c) for c in CASES])
def test_oracle_against_julia(d):
    from oracle import cxx_binding as X
    from oracle.check import assert_result_parity
    ref = _julia(d)
    b, u = _load(d)
    phi, r, (cp, rv, nz) = X.make_phi_r_Kt([b], u)
    assert_result_parity(phi, r, cp, rv, nz, ref, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("d", CASES, ids=[os.path.basename(c) for c in CASES])
def test_cuda_path_against_julia(d):
    from oracle.check import assert_result_parity
    ref = _julia(d)
    b, u = _load(d)
    phi, r, Kt = El.Plan(b, u.shape[1], u.shape[0]).makeϕrKt(u)
    assert_result_parity(phi, r, Kt.colptr, Kt.rowval, Kt.nzval, ref, 1e-12)
