"""Shared builders for the parity tests: seeded jittered meshes through the PRODUCT's constructors, and the
same SoA arrays evaluated by the oracle (C++ restatement of the reference algorithm)."""
import numpy as np

from ad4sm_b200 import Elements as El, Materials as Ma, mesh
from oracle import cxx_binding as X

from oracle.check import assert_result_parity

TOL = 1e-12  # north_star: values to relative 1e-12 in fp64


def nh():
    return Ma.NeoHooke(1000.0, 5000.0)


def hex_case(n=(5, 4, 3), mat=None, seed=1, amp=0.02, ctor=None, dpn=3):
    coords, conn = mesh.hex_block(*n, jitter=0.1, seed=seed)
    b = (ctor or El.Hex08_batch)(conn, coords, mat or nh())
    u = mesh.random_u(len(coords), dpn, 1.0 / max(n), a=amp, seed=seed + 100)
    return b, u


def quad_case(n=(9, 7), mat=None, seed=2, amp=0.02, ctor=None, dpn=2):
    coords, conn = mesh.quad_grid(*n, jitter=0.1, seed=seed)
    b = (ctor or El.Quad_batch)(conn, coords, mat or Ma.Hooke2D(210000.0, 0.3, plane_stress=False))
    u = mesh.random_u(len(coords), dpn, 1.0 / max(n), a=amp, seed=seed + 100)
    return b, u


def tet_case(n=(4, 3, 3), mat=None, seed=3, amp=0.02, ctor=None):
    coords, conn = mesh.tet_kuhn(*n, jitter=0.1, seed=seed)
    b = (ctor or El.Tet04_batch)(conn, coords, mat or Ma.MooneyRivlin(800.0, 200.0, 5000.0))
    u = mesh.random_u(len(coords), 3, 1.0 / max(n), a=amp, seed=seed + 100)
    return b, u


def tri_case(n=(6, 5), mat=None, seed=4, amp=0.02, ctor=None):
    coords, conn = mesh.tri_grid(*n, jitter=0.1, seed=seed)
    b = (ctor or El.Tria_batch)(conn, coords, mat or Ma.Hooke2D(210000.0, 0.3))
    u = mesh.random_u(len(coords), 2, 1.0 / max(n), a=amp, seed=seed + 100)
    return b, u


def random_d(n_nodes, seed=7, hi=0.5):
    return np.random.default_rng(seed).uniform(0.0, hi, n_nodes)


def oracle_eval(batches, u, mode=X.M_ELASTIC, d=None, hist=None, **kw):
    """Reference-algorithm result for the same arrays: ϕ, r, (colptr, rowval, nzval)."""
    if not isinstance(batches, (list, tuple)):
        batches = [batches]
    return X.make_phi_r_Kt(batches, u, mode=mode, d=d, hist=hist, **kw)


def assert_parity(got, ref, tol=TOL, stats=None):
    """(ϕ, r, Kt) from the product vs the oracle -- SURVEY.md §8(c), north_star "relative 1e-12":
      (1) colptr / rowval identical integer for integer;
      (2) every stored value: |a - b| <= tol * max(|a|, |b|, 0.01 * sqrt(|K_ii| |K_jj|)) -- the per-entry relative bound,
          with the entry's magnitude floored at 1 % of its natural scale for entries that are themselves the result of
          cancellation (argued and measured in oracle/check.py).  `stats`, if given, receives how many entries needed
          the floor and the largest plain relative difference;
      (3) r per entry to tol of the residual scale, ϕ relative tol."""
    phi, r, Kt = got
    n_floor, max_rel = assert_result_parity(phi, r, Kt.colptr, Kt.rowval, Kt.nzval, ref, tol)
    if stats is not None:
        stats["n_floor"] = n_floor
        stats["max_rel"] = max_rel


def csc_dense(Kt):
    return Kt.to_scipy().toarray()
