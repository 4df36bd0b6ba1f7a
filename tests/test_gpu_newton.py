"""Newton-step glue on the device (solvers.jl:183-184, 211-224, 257-291) against plain host slicing of the
ORACLE's Kt -- what the reference does: Kt[ifree,ifree] pattern bit-exact and values per entry, res for the predictor and
the corrector, Kt[ifree,icnst]*Δucnst, the two norms; and the whole solvestep loop on the GPU path against the same loop
on the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from ad4sm_b200 import Elements as El, Materials as Ma, Solvers, _capi as capi, mesh
from oracle import cxx_binding as X
from oracle.check import assert_values_entrywise

import _cases as K

pytestmark = pytest.mark.gpu


def _oracle_csc(ref, n):
    cp, rv, nz = ref[2]
    return sp.csc_matrix((nz, rv - 1, cp - 1), shape=(n, n))


def _check_glue(plan, b, u, bfree, ref, field=capi.FIELD_U, rng=None):
    rng = rng or np.random.default_rng(0)
    n = len(ref[1])
    bf = np.asarray(bfree, dtype=bool).reshape(-1, order="F")
    ifree, icnst = np.flatnonzero(bf), np.flatnonzero(~bf)
    nfree, ncnst, nnz = plan.newton_setup(bfree, field)
    assert (nfree, ncnst) == (len(ifree), len(icnst))
    Kt = _oracle_csc(ref, n)
    Kff0 = Kt[ifree][:, ifree].tocsc()
    Kff0.sort_indices()
    assert nnz == Kff0.nnz
    fi = ref[1]
    fe = rng.uniform(-1, 1, n) * np.abs(fi).max()
    du = rng.uniform(-1, 1, ncnst) * 1e-3
    # predictor
    res, Kff, nres, ncn = plan.newton_reduce(0, fe=fe, ducnst=du, field=field)
    assert np.array_equal(Kff.colptr - 1, Kff0.indptr) and np.array_equal(Kff.rowval - 1, Kff0.indices), "Kt[ifree,ifree] pattern"
    assert_values_entrywise(Kff.nzval, Kff0.data, Kff.colptr, Kff.rowval, 1e-12)
    fc0 = Kt[ifree][:, icnst] @ du
    res0 = fi[ifree] - fe[ifree] - fc0
    scale = max(np.abs(fi).max(), np.abs(fe).max(), np.abs(fc0).max() if ncnst else 0.0)
    assert np.abs(res - res0).max() <= 1e-12 * scale
    assert abs(nres - np.linalg.norm(res0)) <= 1e-12 * np.linalg.norm(res0)
    assert abs(ncn - np.linalg.norm(fi[icnst])) <= 1e-12 * max(np.linalg.norm(fi[icnst]), 1e-300)
    # corrector, default fe = zeros, values not requested
    res, none, nres, _ = plan.newton_reduce(1, want_K=False, field=field)
    assert none is None
    assert np.abs(res + fi[ifree]).max() <= 1e-12 * np.abs(fi).max()
    assert abs(nres - np.linalg.norm(fi[ifree])) <= 1e-12 * np.linalg.norm(fi[ifree])
    # the values are a pure gather of the full nzval: bit-identical to slicing the product's own Kt
    _, Kff2, _, _ = plan.newton_reduce(1, field=field)
    mine = plan.fetch_csc(field).to_scipy()[ifree][:, ifree].tocsc()
    mine.sort_indices()
    assert np.array_equal(Kff2.nzval, mine.data)


def test_newton_glue_hex8():
    b, u = K.hex_case((6, 5, 4), seed=41)
    plan = El.Plan(b, u.shape[1], 3)
    plan.makeϕrKt(u, fetch=False)
    ref = K.oracle_eval(b, u)
    rng = np.random.default_rng(1)
    bfree = np.ones(u.shape, dtype=bool)
    bfree[:, :42] = False                       # the bottom plane of 7 x 6 nodes is clamped
    bfree[2, -42:] = False                      # the top plane is displacement-driven in z
    bfree[rng.integers(0, 3, 20), rng.integers(0, u.shape[1], 20)] = False   # and a few scattered dofs
    _check_glue(plan, b, u, bfree, ref, rng=rng)
    # all free / re-setup on the same plan
    _check_glue(plan, b, u, np.ones(u.shape, dtype=bool), ref, rng=rng)


def test_newton_glue_quad4_and_extra_u_row():
    b, u = K.quad_case((9, 7), seed=42)
    plan = El.Plan(b, u.shape[1], 2)
    plan.makeϕrKt(u, fetch=False)
    bfree = np.ones(u.shape, dtype=bool)
    bfree[:, :10] = False
    _check_glue(plan, b, u, bfree, K.oracle_eval(b, u))


def test_newton_glue_phasefield_d_field():
    coords, conn = mesh.quad_grid(8, 6, jitter=0.1, seed=43)
    pf = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
    b = El.QuadP_batch(conn, coords, pf)
    u = mesh.random_u(len(coords), 2, 1 / 8, a=0.3, seed=44)
    d = K.random_d(u.shape[1], seed=45)
    plan = El.Plan(b, u.shape[1], 2)
    plan.makeϕrKt_d(u, d, fetch=False)
    ref = K.oracle_eval(b, u, mode=X.M_PF_D, d=d)
    bfree = np.ones(u.shape[1], dtype=bool)
    bfree[::7] = False
    _check_glue(plan, b, u, bfree, ref, field=capi.FIELD_D)


def test_newton_glue_errors():
    b, u = K.hex_case((3, 3, 2), seed=46)
    plan = El.Plan(b, u.shape[1], 3)
    with pytest.raises(capi.AD4SMError):
        plan._newton = {capi.FIELD_U: (1, 1, 1)}
        plan.newton_reduce(0)                       # before setup
    plan.newton_setup(np.ones(u.shape, dtype=bool))
    with pytest.raises(capi.AD4SMError):
        plan.newton_reduce(1)                       # before any evaluation
    with pytest.raises(ValueError):
        plan.newton_setup(np.ones(5, dtype=bool))


def test_solvestep_matches_host_sliced_oracle_loop():
    """One load step of a clamped Hex8 NeoHooke block, top face pushed down by 5 %: the loop on the GPU path (device glue)
    and the identical loop on the oracle with host slicing must take the same iterations and land on the same u."""
    n = (5, 4, 4)
    coords, conn = mesh.hex_block(*n, jitter=0.1, seed=47)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    nn = len(coords)
    top, bot = coords[:, 2] > coords[:, 2].max() - 1e-9, coords[:, 2] < 1e-9
    bfree = np.ones((3, nn), dtype=bool)
    bfree[:, bot] = False
    bfree[:, top] = False
    uold = np.zeros((3, nn), order="F")

    def run(evaluate):
        unew = np.zeros((3, nn), order="F")
        unew[2, top] = -0.05 * coords[:, 2].max()
        fe = np.zeros(3 * nn)
        plan = El.Plan(b, nn, 3)
        out = Solvers.solvestep(plan, uold, unew, bfree, fe=fe, dTol=1e-10, evaluate=evaluate)
        return out, unew, fe

    def oracle_eval(u2):
        phi, r, (cp, rv, nz) = X.make_phi_r_Kt([b], u2)
        return phi, r, sp.csc_matrix((nz, rv - 1, cp - 1), shape=(u2.size, u2.size))

    (f1, nr1, it1), u1, fe1 = run(None)
    (f2, nr2, it2), u2, fe2 = run(oracle_eval)
    assert not f1 and not f2 and it1 == it2 and it1 >= 3
    assert np.abs(u1 - u2).max() <= 1e-10 * np.abs(u2).max()
    assert np.abs(fe1 - fe2).max() <= 1e-8 * np.abs(fe2).max()
    assert np.abs(u1[:, ~(top | bot)]).max() > 1e-3      # the interior really moved
