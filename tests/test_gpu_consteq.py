"""Constraint-equation catalogue on the device (ad4sm_consteq_eval, Solvers.ConstEq / Solvers.makeϕrKt)
against the literal restatement of the reference's `makeϕrKt(eqns, u, λ)` (solvers.jl:54-73), which evaluates the
closures a user would write with D2 duals (oracle/consteq.py).  Values to 1e-12 of the array scale."""
import numpy as np
import pytest

from ad4sm_b200 import Solvers as S, _capi as capi
from oracle import consteq as OC

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _close(a, b, what):
    a = a.toarray() if hasattr(a, "toarray") else np.asarray(a)
    scale = max(np.abs(b).max(), 1e-300)
    assert a.shape == b.shape, what
    assert np.abs(a - b).max() <= TOL * scale, f"{what}: {np.abs(a - b).max():.3e} at scale {scale:.3e}"


def _problem(seed=5, D=3, n_nodes=40):
    rng = np.random.default_rng(seed)
    u = np.asfortranarray(rng.uniform(-0.1, 0.1, (D, n_nodes)))
    X = rng.uniform(0, 1, (n_nodes, D))
    dof = lambda node, comp: comp + D * node + 1   # 1-based linear index of u[comp, node] (0-based arguments)
    prod, ref = [], []
    # periodic pairs with a prescribed jump: u_a - u_b - jump
    for k in range(6):
        a, b, jump = dof(k, k % D), dof(n_nodes - 1 - k, k % D), 0.01 * (k + 1)
        prod.append(S.ConstEq.periodic(a, b, jump))
        ref.append(OC.ConstEq(lambda v, j=jump: v[0] - v[1] - j, [a, b]))
    # ties: slave = Σ w master
    for k in range(4):
        sl, ms, w = dof(10 + k, 1), [dof(20 + k, 1), dof(21 + k, 1), dof(22 + k, 1)], np.array([0.2, 0.5, 0.3])
        prod.append(S.ConstEq.tie(sl, ms, w))
        ref.append(OC.ConstEq(lambda v, w=w: v[0] - (w[0] * v[1] + w[1] * v[2] + w[2] * v[3]), [sl, *ms]))
    # prescribed average of a component over a node set
    nodes = [3, 7, 11, 19, 23]
    prod.append(S.ConstEq.average([dof(n, 0) for n in nodes], 0.02))
    ref.append(OC.ConstEq(lambda v: (v[0] + v[1] + v[2] + v[3] + v[4]) * (1.0 / 5) - 0.02, [dof(n, 0) for n in nodes]))
    # rigid links |x_a + u_a - x_b - u_b|² = L²
    for (a, b) in [(2, 30), (5, 31), (8, 8 + 1)]:
        da, db = [dof(a, c) for c in range(D)], [dof(b, c) for c in range(D)]
        L = 0.9 * np.linalg.norm(X[a] - X[b])
        prod.append(S.ConstEq.distance(da, db, X[a], X[b], L))

        def f(v, xa=X[a], xb=X[b], L=L, D=D):
            s = 0.0
            for c in range(D):
                t = (xa[c] - xb[c]) + v[c] - v[D + c]
                s = s + t * t
            return s - L * L
        ref.append(OC.ConstEq(f, [*da, *db]))
    # a general quadratic form with a linear part
    idx = [dof(12, 0), dof(13, 2), dof(14, 1), dof(15, 0)]
    A = rng.uniform(-1, 1, (4, 4))
    Q, c = A + A.T, rng.uniform(-1, 1, 4)
    prod.append(S.ConstEq.quadratic(idx, Q, c, 0.3))

    def g(v, Q=Q, c=c):
        s = 0.0
        for i in range(4):
            for j in range(4):
                s = s + (0.5 * Q[i, j]) * v[i] * v[j]
            s = s + c[i] * v[i]
        return s - 0.3
    ref.append(OC.ConstEq(g, idx))
    lam = rng.uniform(-2, 2, len(prod))
    return prod, ref, u, lam


def test_catalogue_matches_reference_closures():
    prod, ref, u, lam = _problem()
    v, r, K = S.makeϕrKt(prod, u, lam)
    v0, r0, K0 = OC.make_phi_r_Kt(ref, u, lam)
    _close(v, v0, "veqs")
    _close(r, r0, "reqs")
    _close(K, K0, "Keqs")
    Kd = K.toarray()
    assert np.array_equal(Kd, Kd.T)
    assert r.shape == (u.size, len(prod)) and K.shape == (u.size, u.size)


def test_linear_only_and_empty():
    prod, ref, u, lam = _problem(seed=6)
    lin = [(p, q) for p, q in zip(prod, ref) if p.Q is None]
    v, r, K = S.makeϕrKt([p for p, _ in lin], u, lam[:len(lin)])
    v0, r0, K0 = OC.make_phi_r_Kt([q for _, q in lin], u, lam[:len(lin)])
    _close(v, v0, "veqs")
    _close(r, r0, "reqs")
    assert K.nnz == 0 and not K0.any()
    v, r, K = S.makeϕrKt([], u, np.zeros(0))
    assert v.shape == (0,) and r.shape == (u.size, 0) and K.shape == (u.size, u.size)


def test_2d_and_overlapping_equations():
    """Two equations sharing dofs: Keqs accumulates in equation order like the reference's `+=`."""
    rng = np.random.default_rng(9)
    u = np.asfortranarray(rng.uniform(-0.1, 0.1, (2, 12)))
    X = rng.uniform(0, 1, (12, 2))
    d = lambda n, c: c + 2 * n + 1
    prod = [S.ConstEq.distance([d(1, 0), d(1, 1)], [d(2, 0), d(2, 1)], X[1], X[2], 0.5),
            S.ConstEq.distance([d(2, 0), d(2, 1)], [d(3, 0), d(3, 1)], X[2], X[3], 0.4)]

    def mk(a, b, L):
        return lambda v: ((X[a][0] - X[b][0]) + v[0] - v[2]) * ((X[a][0] - X[b][0]) + v[0] - v[2]) + \
                         ((X[a][1] - X[b][1]) + v[1] - v[3]) * ((X[a][1] - X[b][1]) + v[1] - v[3]) - L * L
    ref = [OC.ConstEq(mk(1, 2, 0.5), [d(1, 0), d(1, 1), d(2, 0), d(2, 1)]),
           OC.ConstEq(mk(2, 3, 0.4), [d(2, 0), d(2, 1), d(3, 0), d(3, 1)])]
    lam = np.array([1.5, -0.7])
    v, r, K = S.makeϕrKt(prod, u, lam)
    v0, r0, K0 = OC.make_phi_r_Kt(ref, u, lam)
    _close(v, v0, "veqs")
    _close(r, r0, "reqs")
    _close(K, K0, "Keqs")


def test_errors():
    u = np.zeros((3, 4), order="F")
    with pytest.raises(ValueError):
        S.ConstEq.linear([1, 2], [1.0])
    with pytest.raises(ValueError):
        S.ConstEq.quadratic([1, 2], np.array([[1.0, 2.0], [0.0, 1.0]]))
    with pytest.raises(capi.AD4SMError, match="out of range"):
        S.makeϕrKt([S.ConstEq.periodic(1, 13)], u, np.zeros(1))
    with pytest.raises(ValueError):
        S.makeϕrKt([S.ConstEq.periodic(1, 2)], u, np.zeros(2))
