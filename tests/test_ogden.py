"""Ogden (principal-stretch) material.  The reference defines only the scalar energy (elasticmaterials.jl:148-172):
its `Ogden` can neither be constructed (:114-120) nor differentiated (`svdvals` on duals, :162), so r and Kt have no
reference value -- PARITY UNPINNED.  What is checked instead: the product's ψ equals the reference energy formula
(the oracle's literal restatement of it), and P = ∂ψ/∂F, A = ∂²ψ/∂F² equal 30-digit numerical derivatives (mpmath)
of that formula -- at generic F, at F with two equal stretches, and at F = I / F = R (all stretches equal), where
eigenvector-based formulas are singular unless written as divided differences."""
import numpy as np
import pytest

from oracle import cxx_binding as X, materials as M

import _hostmath as H

mp = pytest.importorskip("mpmath")


def _energy_mp(Fm, D, alpha, mu, K):
    """ψ(F) by the reference formula, in mpmath; F is a DxD mp matrix"""
    if D == 3:
        C = Fm.T * Fm
    else:
        F33 = 1 / (Fm[0, 0] * Fm[1, 1] - Fm[1, 0] * Fm[0, 1]) if K < 0 else mp.mpf(1)
        F3 = mp.zeros(3, 3)
        for i in range(2):
            for j in range(2):
                F3[i, j] = Fm[i, j]
        F3[2, 2] = F33
        C = F3.T * F3
    ev = mp.eigsy(C, eigvals_only=True)
    lam = [mp.sqrt(ev[i]) for i in range(3)]
    S = sum(l ** alpha for l in lam)
    if K < 0:
        return mu / alpha * (S - 3)
    J = lam[0] * lam[1] * lam[2]
    return mu / alpha * (S / J ** (mp.mpf(alpha) / 3) - 3) + K * (J - 1) ** 2


def _derivs(F, D, alpha, mu, K):
    mp.mp.dps = 30
    DD = D * D
    x0 = [mp.mpf(float(F[I % D, I // D])) for I in range(DD)]  # linear (column-major) index like the reference

    def f(*x):
        Fm = mp.matrix(D, D)
        for I in range(DD):
            Fm[I % D, I // D] = x[I]
        return _energy_mp(Fm, D, alpha, mu, K)

    P = np.zeros(DD)
    A = np.zeros((DD, DD))
    for i in range(DD):
        o = [0] * DD
        o[i] = 1
        P[i] = float(mp.diff(f, x0, tuple(o), h=mp.mpf(10) ** -8))
        for j in range(i, DD):
            o = [0] * DD
            o[i] += 1
            o[j] += 1
            A[i, j] = A[j, i] = float(mp.diff(f, x0, tuple(o), h=mp.mpf(10) ** -8))
    return float(f(*x0)), P, A


def _rot(th, D):
    if D == 2:
        return np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    a, b = th, 0.7 * th
    R1 = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    R2 = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    return R1 @ R2


def _cases(D):
    rng = np.random.default_rng(5 + D)
    out = [("generic", np.eye(D) + 0.15 * rng.uniform(-1, 1, (D, D))), ("identity", np.eye(D)), ("rotation", _rot(0.4, D))]
    lam = np.array([1.2, 1.2, 0.8][:D]) if D == 3 else np.array([1.1, 1.1])
    out.append(("two-equal-stretches", _rot(0.3, D) @ np.diag(lam) @ _rot(-0.5, D).T))
    out.append(("nearly-equal", _rot(0.3, D) @ np.diag(lam * (1 + 1e-7 * np.arange(D))) @ _rot(-0.5, D).T))
    return out


@pytest.mark.parametrize("K", [5000.0, -1.0], ids=["compressible", "K<0"])
@pytest.mark.parametrize("D", [3, 2])
def test_ogden_tangent_against_mpmath(D, K):
    alpha, mu = 3.0, 1000.0
    mat = M.Ogden(alpha, mu, K)
    desc, params = X.mat_desc(mat)
    for name, F in _cases(D):
        psi0, P0, A0 = _derivs(F, D, alpha, mu, K)
        psi, P, A = H.tangent(D, desc, params, F)
        assert abs(psi - float(M.getphi(F, mat))) <= 1e-12 * max(1.0, abs(psi)), name  # the reference's energy formula
        assert abs(psi - psi0) <= 1e-12 * max(1.0, abs(psi0)), name
        DD = D * D
        Afull = np.zeros((DD, DD))
        for J in range(DD):
            for I in range(J + 1):
                Afull[I, J] = Afull[J, I] = A[J * (J + 1) // 2 + I]
        sP, sA = max(np.abs(P0).max(), mu), np.abs(A0).max()
        assert np.abs(P - P0).max() <= 1e-9 * sP, (name, np.abs(P - P0).max() / sP)
        assert np.abs(Afull - A0).max() <= 1e-8 * sA, (name, np.abs(Afull - A0).max() / sA)


def test_ogden_alpha2_is_neohookean_limit():
    """α = 2: Σλ² = I1, so Ogden(2, μ, K) == NeoHooke(μ/2, K) -- ties the spectral tangent to the invariant-based one."""
    rng = np.random.default_rng(1)
    og, nh = M.Ogden(2.0, 1000.0, 5000.0), M.NeoHooke(500.0, 5000.0)
    d1, p1 = X.mat_desc(og)
    d2, p2 = X.mat_desc(nh)
    for _ in range(5):
        F = np.eye(3) + 0.2 * rng.uniform(-1, 1, (3, 3))
        a, b = H.tangent(3, d1, p1, F), H.tangent(3, d2, p2, F)
        assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0])
        assert np.abs(a[1] - b[1]).max() <= 1e-11 * np.abs(b[1]).max()
        assert np.abs(a[2] - b[2]).max() <= 1e-11 * np.abs(b[2]).max()


@pytest.mark.parametrize("K", [5000.0, -1.0], ids=["compressible", "K<0"])
@pytest.mark.parametrize("alpha", [3.0, 2.0, 4.0])
def test_ogden_element_against_matrix_sqrt_oracle(alpha, K):
    """Two independent derivations of the SAME energy formula, on whole Hex8 / Tet4 elements:
      product (materials.cuh, compiled for the host): spectral tangent with divided differences;
      C++ oracle: D2{9,45} dual arithmetic through Σλ^α = tr(U^α), U = (FᵀF)^(1/2) by a Newton iteration (no
      eigen-decomposition, oracle/cxx/ad4sm_oracle.cpp: phi_ogden3).
    ϕ_e, r_e and K_e must agree to 1e-12 -- including the undeformed element, where all stretches coincide."""
    from oracle import elements as E
    rng = np.random.default_rng(11)
    mat = M.Ogden(alpha, 1000.0, K)
    desc, params = X.mat_desc(mat)
    for ctor, N in ((E.Hex08, 8), (E.Tet04, 4)):
        if N == 8:
            X0 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float).T
        else:
            X0 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float).T
        p0 = X0.T + 0.1 * rng.uniform(-1, 1, (N, 3))
        el = ctor(list(range(1, N + 1)), [q for q in p0], mat)
        for amp in (0.05, 0.0):
            u = amp * rng.uniform(-1, 1, (3, N))
            phi, re, Ke = H.element_u(3, N, desc, params, el.gradN, el.wgt, u.T)
            v, g, Hs = X.elem_duals("CE", 3, el.P, N, mat, np.arange(1, N + 1)[None, :], el.gradN[None], el.wgt[None], None,
                                    X.M_ELASTIC, np.asfortranarray(u))
            sK = np.abs(Hs[0]).max()
            assert abs(phi - v[0]) <= 1e-12 * max(1e-3, abs(v[0])), (N, amp)
            assert np.abs(re - g[0]).max() <= 1e-12 * max(np.abs(g[0]).max(), 1e-3 * sK), (N, amp)
            assert np.abs(Ke - Hs[0]).max() <= 1e-12 * sK, (N, amp, np.abs(Ke - Hs[0]).max() / sK)
