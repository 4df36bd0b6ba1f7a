"""Pins for the ORACLE itself (CPU only).

The reference ships no tests or golden vectors and Julia is not installed, so the oracle cannot be checked
against reference outputs ("parity unpinned by reference tests").  What pins it instead:
  1. the two independent restatements agree (literal Python duals, oracle/*.py  ==  C++ duals, oracle/cxx);
  2. analytic known-answer tests that follow from the reference's own formulas;
  3. independent differentiation: central differences of the *energy only* (ϕ -> r) and of r (-> Kt), and a
     40-digit mpmath gradient/Hessian of one element energy formula;
  4. the assembly semantics of `dropzeros(sparse(II,JJ,Kt,N,N))` (elements.toolkit.jl:202): 1-based CSC, rows
     ascending, duplicates left-folded in `elems` order, stored exact zeros removed;
  5. the committed golden vectors under tests/golden/ (generated from the literal Python restatement by
     tests/golden/make_golden.py).
"""
import glob
import os

import numpy as np
import pytest

from oracle import adiff, cxx_binding as X, elements as E, materials as M

HERE = os.path.dirname(os.path.abspath(__file__))
NH = M.NeoHooke(1000.0, 5000.0)
MR = M.MooneyRivlin(800.0, 200.0, 5000.0)


# ---------------------------------------------------------------------------------------------------
# tiny meshes built with the ORACLE's own constructors (literal restatements of Hex08/Quad/Tet04/...)
# ---------------------------------------------------------------------------------------------------
def _grid(n, D, jitter, seed):
    rng = np.random.default_rng(seed)
    ax = [np.arange(k + 1) for k in n]
    G = np.meshgrid(*ax, indexing="ij")
    ids = sum(g * int(np.prod([k + 1 for k in n[:i]])) for i, g in enumerate(G))
    h = 1.0 / max(n)
    X0 = np.zeros((ids.size, D))
    for i, g in enumerate(G):
        X0[ids.ravel(), i] = g.ravel() * h
    X0 += rng.uniform(-jitter * h, jitter * h, X0.shape)
    return ids, X0


def hex_mesh(n=(2, 2, 1), mat=NH, jitter=0.1, seed=0, ctor=E.Hex08):
    ids, X0 = _grid(n, 3, jitter, seed)
    elems = []
    for k in range(n[2]):
        for j in range(n[1]):
            for i in range(n[0]):
                nd = np.array([ids[i, j, k], ids[i + 1, j, k], ids[i + 1, j + 1, k], ids[i, j + 1, k],
                               ids[i, j, k + 1], ids[i + 1, j, k + 1], ids[i + 1, j + 1, k + 1], ids[i, j + 1, k + 1]]) + 1
                elems.append(ctor(nd, X0[nd - 1], mat))
    return elems, X0


def quad_mesh(n=(3, 2), mat=None, jitter=0.1, seed=0, ctor=E.Quad, shift=None):
    mat = mat or M.Hooke2D(210000.0, 0.3, plane_stress=False)
    ids, X0 = _grid(n, 2, jitter, seed)
    if shift is not None:   # axisymmetric meshes: away from the axis r = 0
        X0 = X0 + np.asarray(shift)
    elems = []
    for j in range(n[1]):
        for i in range(n[0]):
            nd = np.array([ids[i, j], ids[i + 1, j], ids[i + 1, j + 1], ids[i, j + 1]]) + 1
            elems.append(ctor(nd, X0[nd - 1], mat))
    return elems, X0


def tri_mesh(n=(3, 2), mat=None, jitter=0.1, seed=0, ctor=E.Tria, shift=None):
    """every quad of the grid split along its (0,2) diagonal into two counter-clockwise triangles"""
    mat = mat or M.Hooke2D(210000.0, 0.3, plane_stress=False)
    ids, X0 = _grid(n, 2, jitter, seed)
    if shift is not None:
        X0 = X0 + np.asarray(shift)
    elems = []
    for j in range(n[1]):
        for i in range(n[0]):
            q = np.array([ids[i, j], ids[i + 1, j], ids[i + 1, j + 1], ids[i, j + 1]]) + 1
            for nd in (q[[0, 1, 2]], q[[0, 2, 3]]):
                elems.append(ctor(nd, X0[nd - 1], mat))
    return elems, X0


def tet_mesh(mat=MR, jitter=0.1, seed=0, ctor=E.Tet04):
    """one jittered cube split into 6 positively oriented Kuhn tets"""
    ids, X0 = _grid((1, 1, 1), 3, jitter, seed)
    v = lambda i, j, k: ids[i, j, k] + 1  # noqa: E731
    import itertools
    elems = []
    for perm in itertools.permutations(range(3)):
        p = [0, 0, 0]
        path = [v(*p)]
        for ax in perm:
            p[ax] = 1
            path.append(v(*p))
        nd = np.array(path)
        A = np.ones((4, 4))
        A[:, 1:] = X0[nd - 1]
        if np.linalg.det(A) < 0:
            nd = nd[[0, 2, 1, 3]]
        elems.append(ctor(nd, X0[nd - 1], mat))
    return elems, X0


class _Batch:
    """SoA view of a homogeneous element list, in the attribute layout oracle/cxx_binding.py consumes."""

    def __init__(self, elems):
        e0 = elems[0]
        self.kind, self.D, self.P, self.N, self.mat = e0.kind, e0.D, e0.P, e0.N, e0.mat
        self.nodes = np.stack([e.nodes for e in elems]).astype(np.int64)
        self.gradN = np.stack([e.gradN for e in elems])
        self.wgt = np.stack([e.wgt for e in elems])
        self.Nval = np.stack([e.Nval for e in elems]) if e0.kind != "CE" else None
        self.X0 = np.stack([e.X0 for e in elems]) if e0.kind == "CAS" else None


def _u(X0, dpn, amp=0.02, seed=5):
    return np.asfortranarray(amp * np.random.default_rng(seed).uniform(-1, 1, (dpn, len(X0))))


def _dense(csc, n):
    cp, rv, nz = csc
    A = np.zeros((n, n))
    for c in range(n):
        for s in range(cp[c] - 1, cp[c + 1] - 1):
            A[rv[s] - 1, c] = nz[s]
    return A


# ---------------------------------------------------------------------------------------------------
# 1. Python restatement == C++ restatement
# ---------------------------------------------------------------------------------------------------
CROSS = [
    ("hex8-neohooke", lambda: hex_mesh()),
    ("hex8-neohooke-incomp", lambda: hex_mesh(mat=M.NeoHooke(1000.0, -1.0))),
    ("hex8-mooney", lambda: hex_mesh(mat=MR)),
    ("hex8-hooke-gl", lambda: hex_mesh(mat=M.Hooke(210000.0, 0.3))),
    ("hex8-hooke-small", lambda: hex_mesh(mat=M.Hooke(210000.0, 0.3, small=True))),
    ("hex8r-neohooke", lambda: hex_mesh(ctor=E.Hex08R)),
    ("tet4-mooney", lambda: tet_mesh()),
    ("quad4-hooke2d-pe", lambda: quad_mesh()),
    ("quad4-hooke2d-ps-gl", lambda: quad_mesh(mat=M.Hooke2D(210000.0, 0.3, small=False))),
    ("quad4-neohooke", lambda: quad_mesh(mat=NH)),
    ("quad4-mooney-ps", lambda: quad_mesh(mat=M.MooneyRivlin(800.0, 200.0))),
]


@pytest.mark.parametrize("name,make", CROSS, ids=[c[0] for c in CROSS])
def test_python_and_cxx_restatements_agree(name, make):
    elems, X0 = make()
    u = _u(X0, elems[0].D)
    phi, r, (cp, rv, nz) = E.make_phi_r_Kt(elems, u)
    phi2, r2, (cp2, rv2, nz2) = X.make_phi_r_Kt([_Batch(elems)], u)
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
    assert abs(phi - phi2) <= 1e-14 * max(1.0, abs(phi))
    assert np.abs(r - r2).max() <= 1e-13 * np.abs(r).max()
    assert np.abs(nz - nz2).max() <= 1e-13 * np.abs(nz).max()


@pytest.mark.parametrize("make", [lambda: hex_mesh(), lambda: quad_mesh(mat=NH)], ids=["hex8", "quad4"])
def test_element_dual_record_layout(make):
    """Input format of the assembly-only entry points (ad4sm_assemble_duals): one record (v | g[n] | h[n(n+1)/2]) per element
    -- the field order of `struct D2{N,M,T}` (adiff.jl:38-42) with h packed like adiff.jl:26,84.  The literal Python duals
    carry exactly those three fields; `Elements.pack_duals` must build the same record from the C++ oracle's dense Hessian."""
    from ad4sm_b200 import Elements as El
    elems, X0 = make()
    u = _u(X0, elems[0].D)
    Phi = E.element_duals(elems, u)
    rec_py = np.stack([np.concatenate([[p.v], p.g, p.h]) for p in Phi])          # D2 field order: v, g, h
    b = _Batch(elems)
    val, g, H = X.elem_duals(b.kind, b.D, b.P, b.N, b.mat, b.nodes, b.gradN, b.wgt, b.Nval, X.M_ELASTIC, u)
    rec = El.pack_duals(val, g, H)
    n = b.N * b.D
    assert rec.shape == rec_py.shape == (len(elems), 1 + n + n * (n + 1) // 2)
    assert np.abs(rec - rec_py).max() <= 1e-13 * np.abs(rec_py).max()
    for i, j in ((0, 0), (1, 4), (4, 1), (n - 1, n - 1), (2, n - 1)):               # entry (i,j) of the Hessian sits at packed_index
        assert rec[0, 1 + n + adiff.packed_index(i, j)] == H[0, i, j]


def test_python_and_cxx_agree_phasefield():
    pf = M.PhaseField(0.01, 100.0, M.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
    elems, X0 = quad_mesh((2, 2), mat=pf, ctor=E.QuadP)
    u = _u(X0, 2, amp=0.05)
    d = np.random.default_rng(1).uniform(0, 0.5, len(X0))
    a = E.make_phi_r_Kt_pf_u(elems, u, d)
    b = X.make_phi_r_Kt([_Batch(elems)], u, mode=X.M_PF_U, d=d)
    assert np.array_equal(a[2][1], b[2][1])
    assert np.abs(a[2][2] - b[2][2]).max() <= 1e-13 * np.abs(a[2][2]).max()
    for gt in ("reference", "intended"):
        h1 = [[(0.0, 0.0)] * e.P for e in elems]
        h2 = np.zeros((len(elems), elems[0].P, 2))
        a = E.make_phi_r_Kt_pf_d(elems, u, d, h1, grad_term=gt)
        b = X.make_phi_r_Kt([_Batch(elems)], u, mode=X.M_PF_D, d=d, hist=[h2], grad_term=gt)
        assert np.array_equal(a[2][1], b[2][1])
        assert np.abs(a[1] - b[1]).max() <= 1e-13 * np.abs(a[1]).max()
        assert np.abs(a[2][2] - b[2][2]).max() <= 1e-13 * np.abs(a[2][2]).max()
        assert np.abs(np.array(h1, dtype=float) - h2).max() <= 1e-13 * max(1.0, np.abs(h2).max())


# ---------------------------------------------------------------------------------------------------
# 2. analytic known answers
# ---------------------------------------------------------------------------------------------------
STRESS_FREE = [NH, MR, M.Hooke(210000.0, 0.3), M.Hooke(210000.0, 0.3, small=True)]


@pytest.mark.parametrize("mat", STRESS_FREE, ids=lambda m: type(m).__name__)
def test_rigid_translation_gives_zero_energy_and_residual(mat):
    elems, X0 = hex_mesh(mat=mat)
    u = np.asfortranarray(np.tile(np.array([[0.3], [-0.2], [0.1]]), (1, len(X0))))
    phi, r, csc = X.make_phi_r_Kt([_Batch(elems)], u)
    ks = np.abs(csc[2]).max()
    assert abs(phi) <= 1e-12 and np.abs(r).max() <= 1e-12 * ks


@pytest.mark.parametrize("mat", [NH, MR, M.Hooke(210000.0, 0.3)], ids=lambda m: type(m).__name__)
def test_finite_rigid_rotation_is_stress_free(mat):
    elems, X0 = hex_mesh(mat=mat)
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    u = np.asfortranarray(((X0 @ R.T) - X0).T)
    phi, r, csc = X.make_phi_r_Kt([_Batch(elems)], u)
    ks = np.abs(csc[2]).max()
    assert abs(phi) <= 1e-9 and np.abs(r).max() <= 1e-10 * ks


@pytest.mark.parametrize("mat", STRESS_FREE + [M.NeoHooke(1000.0, -1.0)], ids=lambda m: repr(m))
def test_tangent_annihilates_translations_at_any_u(mat):
    """ϕ depends on ∇u only, so Kt·t = 0 for a rigid translation t, at any u and for any material."""
    elems, X0 = hex_mesh(mat=mat)
    u = _u(X0, 3)
    _, _, csc = X.make_phi_r_Kt([_Batch(elems)], u)
    K = _dense(csc, u.size)
    for j in range(3):
        t = np.zeros((3, len(X0)))
        t[j] = 1.0
        assert np.abs(K @ t.reshape(-1, order="F")).max() <= 1e-11 * np.abs(K).max()


def test_homogeneous_stretch_closed_form_neohooke():
    """F = diag(λ) on the unit cube: ϕ = V·ψ(F), ψ = C1 (I1 J^(-2/3) - 3) + K (J-1)²  (elasticmaterials.jl:186-197)."""
    lam = np.array([1.10, 0.95, 1.05])
    p0 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1.0]])
    el = E.Hex08(np.arange(1, 9), p0, NH)
    u = np.asfortranarray((p0 * (lam - 1.0)).T)
    phi, r, _ = X.make_phi_r_Kt([_Batch([el])], u)
    J = lam.prod()
    psi = NH.C1 * ((lam ** 2).sum() * J ** (-2.0 / 3.0) - 3.0) + NH.K * (J - 1.0) ** 2
    assert abs(phi - psi) <= 1e-12 * psi
    # consistent nodal forces of the (diagonal) first Piola-Kirchhoff stress: face resultants P_ii on unit faces
    h = 1e-6
    P = np.zeros(3)
    for i in range(3):
        lp, lm = lam.copy(), lam.copy()
        lp[i] += h
        lm[i] -= h
        f = lambda l: NH.C1 * ((l ** 2).sum() * l.prod() ** (-2.0 / 3.0) - 3.0) + NH.K * (l.prod() - 1.0) ** 2  # noqa: E731
        P[i] = (f(lp) - f(lm)) / (2 * h)
    R = r.reshape(3, 8, order="F")
    for i in range(3):
        plus = p0[:, i] == 1.0
        assert abs(R[i, plus].sum() - P[i]) <= 1e-6 * abs(P).max()
        assert abs(R[i].sum()) <= 1e-9 * abs(P).max()


def test_linear_hooke2d_is_quadratic_form():
    """Hooke2D small strain (the default, elasticmaterials.jl:70): Kt independent of u, r = Kt u, ϕ = u·r/2."""
    elems, X0 = quad_mesh(mat=M.Hooke2D(210000.0, 0.3))
    b = [_Batch(elems)]
    u1, u2 = _u(X0, 2, seed=1), _u(X0, 2, seed=2)
    phi1, r1, c1 = X.make_phi_r_Kt(b, u1)
    _, _, c2 = X.make_phi_r_Kt(b, u2)
    K1, K2 = _dense(c1, u1.size), _dense(c2, u2.size)
    assert np.abs(K1 - K2).max() <= 1e-12 * np.abs(K1).max()
    uf = u1.reshape(-1, order="F")
    assert np.abs(K1 @ uf - r1).max() <= 1e-11 * np.abs(r1).max()
    assert abs(0.5 * uf @ r1 - phi1) <= 1e-11 * abs(phi1)


@pytest.mark.parametrize("name,make", CROSS, ids=[c[0] for c in CROSS])
def test_tangent_is_bitwise_symmetric(name, make):
    """Both triangles come from one packed Hessian slot (adiff.jl:160) and the same summation order."""
    elems, X0 = make()
    u = _u(X0, elems[0].D)
    _, _, csc = X.make_phi_r_Kt([_Batch(elems)], u)
    K = _dense(csc, u.size)
    assert np.array_equal(K, K.T)


# ---------------------------------------------------------------------------------------------------
# 3. independent differentiation
# ---------------------------------------------------------------------------------------------------
def _energy(elems, u):
    """value-only evaluation through the literal Python restatement (plain floats, no duals)."""
    return sum(float(E.getphi_elem(e, u[:, e.nodes - 1])) for e in elems)


@pytest.mark.parametrize("name,make", [CROSS[0], CROSS[2], CROSS[3], CROSS[6], CROSS[9]], ids=lambda x: x if isinstance(x, str) else "")
def test_residual_and_tangent_match_central_differences(name, make):
    elems, X0 = make()
    D = elems[0].D
    u = _u(X0, D)
    b = [_Batch(elems)]
    phi, r, csc = X.make_phi_r_Kt(b, u)
    assert abs(phi - _energy(elems, u)) <= 1e-13 * max(1.0, abs(phi))
    K = _dense(csc, u.size)
    rng = np.random.default_rng(3)
    h = 1e-6
    for _ in range(6):
        k = int(rng.integers(u.size))
        du = np.zeros(u.size)
        du[k] = h
        du = du.reshape(u.shape, order="F")
        # 4th-order central difference of the energy -> r[k]
        e = lambda s: _energy(elems, u + s * du)  # noqa: E731
        fd = (-e(2) + 8 * e(1) - 8 * e(-1) + e(-2)) / (12 * h)
        assert abs(fd - r[k]) <= 2e-6 * np.abs(r).max(), (fd, r[k])
        # central difference of r -> column k of Kt
        rp = X.make_phi_r_Kt(b, u + du)[1]
        rm = X.make_phi_r_Kt(b, u - du)[1]
        col = (rp - rm) / (2 * h)
        assert np.abs(col - K[:, k]).max() <= 1e-6 * np.abs(K).max()


def test_tet4_neohooke_hessian_against_mpmath():
    """Gradient and Hessian of one Tet4 NeoHooke element energy by 40-digit numerical differentiation (mpmath) of the
    energy FORMULA (elasticmaterials.jl:186-197 with materials.jl:111-113), independent of any dual arithmetic."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    elems, X0 = tet_mesh(mat=NH)
    el = elems[0]
    u_all = _u(X0, 3)
    uvals = u_all[:, el.nodes - 1]
    G = [[mp.mpf(float(el.gradN[k, 0, a])) for a in range(4)] for k in range(3)]
    w, C1, K = mp.mpf(float(el.wgt[0])), mp.mpf(NH.C1), mp.mpf(NH.K)

    def energy(*x):  # x in element dof order (a-1)*3 + j  (adiff.jl:58-62)
        F = [[(1 if j == k else 0) + sum(G[k][a] * x[a * 3 + j] for a in range(4)) for k in range(3)] for j in range(3)]
        C = [[sum(F[i][j] * F[i][k] for i in range(3)) for k in range(3)] for j in range(3)]
        I1 = C[0][0] + C[1][1] + C[2][2]
        I3 = (C[0][0] * C[1][1] * C[2][2] + 2 * C[0][1] * C[0][2] * C[1][2] - C[0][0] * C[1][2] ** 2
              - C[1][1] * C[0][2] ** 2 - C[2][2] * C[0][1] ** 2)
        J = mp.sqrt(I3)
        return w * (C1 * (I1 * J ** (mp.mpf(-2) / 3) - 3) + K * (J - 1) ** 2)

    x0 = [mp.mpf(float(uvals[j, a])) for a in range(4) for j in range(3)]
    n = 12
    g = np.zeros(n)
    H = np.zeros((n, n))
    for i in range(n):
        o = [0] * n
        o[i] = 1
        g[i] = float(mp.diff(energy, x0, tuple(o)))
        for j in range(i, n):
            o = [0] * n
            o[i] += 1
            o[j] += 1
            H[i, j] = H[j, i] = float(mp.diff(energy, x0, tuple(o)))
    val, gg, HH = X.elem_duals("CE", 3, 1, 4, NH, el.nodes[None], el.gradN[None], el.wgt[None], None, X.M_ELASTIC,
                               np.asfortranarray(u_all))
    assert abs(val[0] - float(energy(*x0))) <= 1e-13 * abs(val[0])
    assert np.abs(gg[0] - g).max() <= 1e-12 * np.abs(g).max()
    assert np.abs(HH[0] - H).max() <= 1e-12 * np.abs(H).max()


# ---------------------------------------------------------------------------------------------------
# 4. assembly semantics: dropzeros(sparse(II, JJ, Kt, N, N))   (elements.toolkit.jl:169-203)
# ---------------------------------------------------------------------------------------------------
def test_assembly_left_fold_in_elems_order():
    """Three 1-node 'elements' on the same dof: (1e16 + 1) - 1e16 differs from 1e16 - 1e16 + 1 in fp64, so the
    summation order is observable."""
    nodes = np.array([[1], [1], [1]], dtype=np.int64)
    val = np.array([1.0, 2.0, 3.0])
    g = np.array([[1e16], [1.0], [-1e16]])
    H = g.reshape(3, 1, 1).copy()
    phi, r, (cp, rv, nz) = X.assemble(nodes, 1, val, g, H, 1)
    assert phi == 6.0
    assert r[0] == (1e16 + 1.0) - 1e16 == 0.0  # left fold: the 1.0 is absorbed
    assert len(nz) == 0, "an exact zero must be dropped (dropzeros)"
    g2 = g[[0, 2, 1]]
    phi, r, (cp, rv, nz) = X.assemble(nodes, 1, val, g2, g2.reshape(3, 1, 1).copy(), 1)
    assert r[0] == 1.0 and list(nz) == [1.0] and list(cp) == [1, 2] and list(rv) == [1]


def test_assembly_csc_layout_and_dropzeros():
    # two 2-node bars (dpn = 1) sharing node 2; the second one has an exactly cancelling off-diagonal
    nodes = np.array([[1, 2], [2, 3]], dtype=np.int64)
    val = np.zeros(2)
    g = np.array([[1.0, -1.0], [2.0, -2.0]])
    H = np.array([[[2.0, -2.0], [-2.0, 2.0]], [[5.0, 0.0], [0.0, 5.0]]])
    phi, r, (cp, rv, nz) = X.assemble(nodes, 1, val, g, H, 3)
    assert list(r) == [1.0, 1.0, -2.0]
    assert list(cp) == [1, 3, 5, 6]  # column 3 holds only the diagonal: the (2,3)/(3,2) zeros are dropped
    assert list(rv) == [1, 2, 1, 2, 3]
    assert list(nz) == [2.0, -2.0, -2.0, 7.0, 5.0]
    # the literal Python restatement agrees
    cp2, rv2, nz2 = E.coo_to_csc_leftfold(np.array([1, 2, 1, 2, 2, 3, 2, 3]), np.array([1, 1, 2, 2, 2, 2, 3, 3]),
                                          np.array([2.0, -2.0, -2.0, 2.0, 5.0, 0.0, 0.0, 5.0]), 3)
    assert list(cp2) == list(cp) and list(rv2) == list(rv) and list(nz2) == list(nz)


def test_dof_numbering_is_node_major():
    """DOF id = (node-1)*size(u,1) + comp (toolkit:183,187): r is flat, column-major over u."""
    elems, X0 = quad_mesh((1, 1))
    u = _u(X0, 2)
    _, r, _ = X.make_phi_r_Kt([_Batch(elems)], u)
    el = elems[0]
    dual = E.getphi_elem(el, adiff.seed_D2(u[:, el.nodes - 1]))
    g = adiff.grad(dual)  # element dof order: (a-1)*D + j
    for a, nd in enumerate(el.nodes):
        for j in range(2):
            assert abs(r[(nd - 1) * 2 + j] - g[a * 2 + j]) <= 1e-14 * np.abs(g).max()


# ---------------------------------------------------------------------------------------------------
# 5. golden vectors (generated by tests/golden/make_golden.py from the literal Python restatement)
# ---------------------------------------------------------------------------------------------------
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cxx_oracle_reproduces_golden_vectors(path):
    import _golden
    case = _golden.load(path)
    got = _golden.run_oracle(case)
    _golden.compare(got, case, tol=1e-13)


def test_golden_vectors_exist():
    assert len(GOLDEN) >= 8, "run tests/golden/make_golden.py"


# ---------------------------------------------------------------------------------------------------
# 6. CAS (axisymmetric extension): the reference only has the constructors; the energy is pinned by
#    closed forms and independent differentiation
# ---------------------------------------------------------------------------------------------------
def _cas_case(ctor=E.ASQuad, mat=NH, n=(3, 2), seed=41):
    mesh_fn = quad_mesh if ctor is E.ASQuad else tri_mesh
    return mesh_fn(n, mat, seed=seed, ctor=ctor, shift=(0.7, 0.0))


@pytest.mark.parametrize("ctor", [E.ASQuad, E.ASTria], ids=["asquad", "astria"])
@pytest.mark.parametrize("mat", [NH, MR, M.Hooke(210000.0, 0.3)], ids=["neohooke", "mooney", "hooke"])
def test_cas_homogeneous_expansion_closed_form(ctor, mat):
    """u_r = ε r, u_z = γ z  =>  F = diag(1+ε, 1+γ, 1+ε) everywhere (the hoop stretch equals the radial one), so
    ϕ = ϕ(F) · Σ wgt exactly (both shape-function sets reproduce linear fields) and Σ wgt = 2π ∫ r dA."""
    elems, X0 = _cas_case(ctor, mat)
    eps, gam = 0.03, -0.02
    u = np.asfortranarray(np.stack([eps * X0[:, 0], gam * X0[:, 1]]))
    phi, r, _ = E.make_phi_r_Kt(elems, u)
    F = np.diag([1 + eps, 1 + gam, 1 + eps]).astype(object)
    V = sum(float(e.wgt.sum()) for e in elems)
    assert abs(phi - M.getphi(F, mat) * V) <= 1e-12 * abs(phi)


def test_cas_volume_of_revolution():
    """wgt = detJ·2π·r (elasticelements.jl:249, 279): for a straight-sided quad mesh Σ wgt = 2π ∫ r dA is integrated exactly
    by the 2x2 rule; compare with the shoelace/Pappus value of the mesh outline."""
    elems, X0 = quad_mesh((3, 2), NH, jitter=0.0, seed=1, ctor=E.ASQuad, shift=(0.7, 0.0))
    V = sum(float(e.wgt.sum()) for e in elems)
    r0, r1, h = X0[:, 0].min(), X0[:, 0].max(), X0[:, 1].max() - X0[:, 1].min()
    assert abs(V - np.pi * (r1 ** 2 - r0 ** 2) * h) <= 1e-13 * V


@pytest.mark.parametrize("ctor", [E.ASQuad, E.ASTria], ids=["asquad", "astria"])
def test_cas_axial_translation_and_central_differences(ctor):
    """a rigid axial translation costs nothing (ϕ = 0, r = 0, Kt·t = 0 at any u); r and Kt are the derivatives of ϕ
    (4th-order central differences), Kt symmetric."""
    elems, X0 = _cas_case(ctor, NH)
    nn = len(X0)
    t = np.zeros((2, nn))
    t[1] = 1.0
    phi, r, _ = E.make_phi_r_Kt(elems, 0.37 * t)
    assert abs(phi) <= 1e-12 and np.abs(r).max() <= 1e-9
    u = _u(X0, 2, seed=43)
    phi, r, csc = E.make_phi_r_Kt(elems, u)
    Kd = _dense(csc, 2 * nn)
    assert np.abs(Kd - Kd.T).max() == 0.0
    assert np.abs(Kd @ t.reshape(-1, order="F")).max() <= 1e-9 * np.abs(Kd).max()
    h = 1e-4
    rng = np.random.default_rng(44)
    for _ in range(3):
        dvec = rng.uniform(-1, 1, u.shape)
        f = lambda s: E.make_phi_r_Kt(elems, u + s * h * dvec)  # noqa: E731
        (pm2, rm2, _), (pm1, rm1, _), (pp1, rp1, _), (pp2, rp2, _) = f(-2), f(-1), f(1), f(2)
        dphi = (pm2 - 8 * pm1 + 8 * pp1 - pp2) / (12 * h)
        assert abs(dphi - r @ dvec.reshape(-1, order="F")) <= 1e-9 * max(1.0, abs(dphi))
        dr = (rm2 - 8 * rm1 + 8 * rp1 - rp2) / (12 * h)
        assert np.abs(dr - Kd @ dvec.reshape(-1, order="F")).max() <= 1e-8 * np.abs(dr).max()


def test_cas_python_and_cxx_restatements_agree():
    for ctor, mat in ((E.ASQuad, NH), (E.ASQuad, M.Hooke(210000.0, 0.3, small=True)), (E.ASTria, MR)):
        elems, X0 = _cas_case(ctor, mat, seed=45)
        u = _u(X0, 2, seed=46)
        phi, r, (cp, rv, nz) = E.make_phi_r_Kt(elems, u)
        phi2, r2, (cp2, rv2, nz2) = X.make_phi_r_Kt([_Batch(elems)], u)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
        assert abs(phi - phi2) <= 1e-13 * abs(phi) and np.abs(r - r2).max() <= 1e-13 * np.abs(r).max()
        assert np.abs(nz - nz2).max() <= 1e-13 * np.abs(nz).max()
