"""CPU: the host mirror of Solvers.solvestep! (solvers.jl:172-300, nEqs == 0) driven by the ORACLE through its test hook
-- the control flow, norms and update rule -- without a GPU."""
import numpy as np
import scipy.sparse as sp

from ad4sm_b200 import Elements as El, Materials as Ma, Solvers, mesh
from oracle import cxx_binding as X


def _problem(n=(4, 3, 3)):
    coords, conn = mesh.hex_block(*n, jitter=0.1, seed=3)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    nn = len(coords)
    top, bot = coords[:, 2] > coords[:, 2].max() - 1e-9, coords[:, 2] < 1e-9
    bfree = np.ones((3, nn), dtype=bool)
    bfree[:, bot] = False
    bfree[:, top] = False

    def ev(u2):
        phi, r, (cp, rv, nz) = X.make_phi_r_Kt([b], u2)
        return phi, r, sp.csc_matrix((nz, rv - 1, cp - 1), shape=(u2.size, u2.size))
    return coords, top, bfree, ev


def test_solvestep_converges_and_reports_reactions():
    coords, top, bfree, ev = _problem()
    nn = len(coords)
    uold = np.zeros((3, nn), order="F")
    unew = np.zeros((3, nn), order="F")
    unew[2, top] = -0.04 * coords[:, 2].max()
    fe = np.zeros(3 * nn)
    bfailed, normru, it = Solvers.solvestep(None, uold, unew, bfree, fe=fe, dTol=1e-9, evaluate=ev)
    assert not bfailed and normru < 1e-9 and 2 <= it <= 8
    # at convergence fe = fi (solvers.jl:281): zero on the free dofs, reactions on the constrained ones
    free = bfree.reshape(-1, order="F")
    assert np.abs(fe[free]).max() <= 1e-6 * np.abs(fe[~free]).max()
    assert abs(fe.reshape(3, nn, order="F")[2].sum()) <= 1e-8 * np.abs(fe).max()   # global equilibrium in z
    assert np.abs(unew[2, top] + 0.04 * coords[:, 2].max()).max() == 0.0           # prescribed values untouched


def test_solvestep_without_predictor_and_maxiter_failure():
    coords, top, bfree, ev = _problem()
    nn = len(coords)
    uold = np.zeros((3, nn), order="F")
    unew = np.zeros((3, nn), order="F")
    unew[2, top] = -0.04 * coords[:, 2].max()
    bfailed, _, it = Solvers.solvestep(None, uold, unew.copy(order="F"), bfree, dTol=1e-9, bpredict=False, evaluate=ev)
    assert not bfailed and it >= 3
    bfailed, _, it = Solvers.solvestep(None, uold, unew.copy(order="F"), bfree, dTol=1e-14, maxiter=1, bpredict=False, evaluate=ev)
    assert bfailed and it == 2
    # step-size limiter (:287-292): a tiny maxupdt cannot converge in 3 iterations
    bfailed, _, _ = Solvers.solvestep(None, uold, unew.copy(order="F"), bfree, dTol=1e-9, maxiter=3, maxupdt=1e-6, evaluate=ev)
    assert bfailed
