"""The product's device math (csrc/materials.cuh, csrc/element.cuh), compiled for the host, against the
oracle's literal dual-number restatement of the reference.  CPU only."""
import numpy as np
import pytest

from oracle import adiff, cxx_binding as X, elements as E, materials as M

import _hostmath as H

TOL = 1e-12


def _rel(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() / (s if s > 0 else 1.0)


def _randF(rng, D, amp=0.15):
    return np.eye(D) + amp * rng.uniform(-1, 1, (D, D))


MATS3 = [
    M.Hooke(210000.0, 0.3), M.Hooke(210000.0, 0.3, small=True),
    M.NeoHooke(1000.0, 5000.0), M.NeoHooke(1000.0, -1.0),
    M.MooneyRivlin(800.0, 200.0, 5000.0), M.MooneyRivlin(800.0, 200.0),
]
MATS2 = [
    M.Hooke2D(210000.0, 0.3), M.Hooke2D(210000.0, 0.3, plane_stress=False),
    M.Hooke2D(210000.0, 0.3, small=False), M.Hooke2D(210000.0, 0.3, small=False, plane_stress=False),
    M.NeoHooke(1000.0, 5000.0), M.NeoHooke(1000.0, -1.0),
    M.MooneyRivlin(800.0, 200.0, 5000.0), M.MooneyRivlin(800.0, 200.0),
]


@pytest.mark.parametrize("mat", MATS3 + MATS2, ids=lambda m: repr(m))
def test_material_tangent_matches_oracle_duals(mat):
    """(ψ, ∂ψ/∂F, ∂²ψ/∂F²) == payload of getϕ(adiff.D2(F), mat)  (elasticelements.jl:346)."""
    D = 2 if mat in MATS2 and not (mat in MATS3 and not isinstance(mat, (M.NeoHooke, M.MooneyRivlin))) else 3
    for D in ((2, 3) if isinstance(mat, (M.NeoHooke, M.MooneyRivlin)) else ((2,) if isinstance(mat, M.Hooke2D) else (3,))):
        rng = np.random.default_rng(7 + D)
        for _ in range(5):
            F = _randF(rng, D)
            ref = M.getphi(adiff.seed_D2(F), mat)
            desc, params = X.mat_desc(mat)
            psi, P, A = H.tangent(D, desc, params, F)
            assert abs(psi - ref.v) <= TOL * max(1.0, abs(ref.v))
            assert _rel(P, ref.g) < TOL
            assert _rel(A, ref.h) < TOL


PF3 = [
    M.PhaseField(0.01, 100.0, M.Hooke(210000.0, 0.3, small=True), 2),
    M.PhaseField(0.01, 100.0, M.Hooke(210000.0, 0.3), 2),
    M.PhaseField(0.01, 100.0, M.Hooke(210000.0, 0.3, small=True), 2, AT="DHn"),
    M.PhaseField(0.01, 100.0, M.NeoHooke(1000.0, 5000.0), 2),
    M.PhaseField(0.01, 100.0, M.NeoHooke(1000.0, 5000.0), 1),
]
PF2 = [
    M.PhaseField(0.01, 100.0, M.Hooke2D(210000.0, 0.3, plane_stress=False), 2),
    M.PhaseField(0.01, 100.0, M.Hooke2D(210000.0, 0.3), 2),
    M.PhaseField(0.01, 100.0, M.Hooke2D(210000.0, 0.3), 3, AT="DHn"),
    M.PhaseField(0.01, 100.0, M.Hooke2D(210000.0, 0.3, plane_stress=False, small=False), 2, AT="DHn"),
    M.PhaseField(0.01, 100.0, M.NeoHooke(1000.0, 5000.0), 2),
]


@pytest.mark.parametrize("mat,D", [(m, 3) for m in PF3] + [(m, 2) for m in PF2], ids=lambda m: repr(m))
def test_phasefield_u_tangent_matches_oracle_duals(mat, D):
    rng = np.random.default_rng(11)
    for trial in range(8):
        F = _randF(rng, D)
        if trial % 2:  # compressive branch
            F = F - 0.3 * np.eye(D)
        d = rng.uniform(0, 0.8)
        gd = rng.uniform(-1, 1, D)
        ext = D == 2 and isinstance(mat.mat, M.NeoHooke)
        ref = M.getphi_pf(adiff.seed_D2(F), d, list(gd), mat, nh2d_extension=ext) if ext else \
            M.getphi_pf(adiff.seed_D2(F), d, list(gd), mat)
        desc, params = X.mat_desc(mat, nh2d_extension=ext)
        psi, P, A = H.tangent(D, desc, params, F, pf=(d, float(gd @ gd)))
        assert abs(psi - ref.v) <= TOL * max(1.0, abs(ref.v))
        assert _rel(P, ref.g) < TOL
        assert _rel(A, ref.h) < TOL


def _one_elem(ctor, D, N, mat, rng, jitter=0.1):
    ref = {
        (2, 3): [[0, 0], [1, 0], [0, 1]],
        (2, 4): [[0, 0], [1, 0], [1, 1], [0, 1]],
        (3, 4): [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
        (3, 6): [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]],
        (3, 8): [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
    }[(D, N)]
    p0 = np.array(ref, float) + jitter * rng.uniform(-1, 1, (N, D))
    return ctor(list(range(1, N + 1)), [p for p in p0], mat)


ELASTIC_CASES = [
    (E.Hex08, 3, 8, M.NeoHooke(1000.0, 5000.0)),
    (E.Hex08, 3, 8, M.MooneyRivlin(800.0, 200.0, 5000.0)),
    (E.Hex08, 3, 8, M.Hooke(210000.0, 0.3)),
    (E.Hex08R, 3, 8, M.NeoHooke(1000.0, 5000.0)),
    (E.Tet04, 3, 4, M.MooneyRivlin(800.0, 200.0, 5000.0)),
    (E.Tet04, 3, 4, M.Hooke(210000.0, 0.3, small=True)),
    (E.Quad, 2, 4, M.Hooke2D(210000.0, 0.3, plane_stress=False)),
    (E.Quad, 2, 4, M.Hooke2D(210000.0, 0.3, small=False)),
    (E.Quad, 2, 4, M.NeoHooke(1000.0, 5000.0)),
    (E.Quad, 2, 4, M.MooneyRivlin(800.0, 200.0)),
    (E.QuadR, 2, 4, M.NeoHooke(1000.0, -1.0)),
    (E.Tria, 2, 3, M.Hooke2D(210000.0, 0.3)),
    (E.Tria, 2, 3, M.NeoHooke(1000.0, 5000.0)),
]


@pytest.mark.parametrize("ctor,D,N,mat", ELASTIC_CASES, ids=lambda x: getattr(x, "__name__", repr(x)))
def test_element_tangent_matches_oracle(ctor, D, N, mat):
    """phase 1 + phase 2 of K1 for one element == getϕ(elem, adiff.D2(u_e)) (elasticelements.jl:216-218)."""
    rng = np.random.default_rng(3)
    el = _one_elem(ctor, D, N, mat, rng)
    u = 0.05 * rng.uniform(-1, 1, (D, N))
    ref = E.getphi_elem(el, adiff.seed_D2(u))
    desc, params = X.mat_desc(mat)
    phi, re, K = H.element_u(D, N, desc, params, el.gradN, el.wgt, u.T)
    Href = adiff.hess(ref)
    assert abs(phi - ref.v) <= TOL * max(1.0, abs(ref.v))
    assert _rel(re, adiff.grad(ref)) < TOL
    assert _rel(K, Href) < TOL
    assert np.array_equal(K, K.T)


PF_ELEM_CASES = [
    (E.QuadP, 2, 4, PF2[0]), (E.QuadP, 2, 4, PF2[1]), (E.QuadP, 2, 4, PF2[2]), (E.QuadP, 2, 4, PF2[4]),
    (E.TriaP, 2, 3, PF2[0]), (E.Hex08P, 3, 8, PF3[0]), (E.Hex08P, 3, 8, PF3[3]), (E.Hex08P, 3, 8, PF3[2]),
    (E.Tet04P, 3, 4, PF3[3]), (E.Wdg06P, 3, 6, PF3[0]), (E.Wdg06P, 3, 6, PF3[3]),
]


@pytest.mark.parametrize("ctor,D,N,mat", PF_ELEM_CASES, ids=lambda x: getattr(x, "__name__", repr(x)))
def test_phasefield_element_u_phase(ctor, D, N, mat):
    rng = np.random.default_rng(5)
    el = _one_elem(ctor, D, N, mat, rng)
    u = 0.05 * rng.uniform(-1, 1, (D, N))
    d = rng.uniform(0, 0.6, N)
    ext = D == 2 and isinstance(mat.mat, M.NeoHooke)
    ref = E.getphi_elem_pf(el, adiff.seed_D2(u), d, nh2d_extension=ext)
    desc, params = X.mat_desc(mat, nh2d_extension=ext)
    phi, re, K = H.element_u(D, N, desc, params, el.gradN, el.wgt, u.T, Nv=el.Nval, de=d)
    assert abs(phi - ref.v) <= TOL * max(1.0, abs(ref.v))
    assert _rel(re, adiff.grad(ref)) < TOL
    assert _rel(K, adiff.hess(ref)) < TOL


@pytest.mark.parametrize("grad_term", ["reference", "intended"])
@pytest.mark.parametrize("with_hist", [False, True])
@pytest.mark.parametrize("ctor,D,N,mat", PF_ELEM_CASES, ids=lambda x: getattr(x, "__name__", repr(x)))
def test_phasefield_element_d_phase(ctor, D, N, mat, with_hist, grad_term):
    rng = np.random.default_rng(9)
    el = _one_elem(ctor, D, N, mat, rng)
    ext = D == 2 and isinstance(mat.mat, M.NeoHooke)
    neo = isinstance(mat.mat, M.NeoHooke) and mat.AT == "ATn"
    for trial in range(4):
        u = 0.05 * rng.uniform(-1, 1, (D, N))
        if trial % 2:
            u = u - 0.2 * np.array([p for p in (el.gradN[:, 0, :] * 0 + 1)])[:D] * 0  # keep shape
            u = u - 0.1 * rng.uniform(0.5, 1, (D, 1)) * np.sign(rng.uniform(-1, 1, (D, N)))
        d = rng.uniform(0, 0.6, N)
        P = el.P
        hist_o = hist_p = None
        if with_hist:
            h0 = rng.uniform(0, 5.0, (P, 2)) * (trial > 1)
            hist_o = [(h0[g, 0] if neo else (h0[g, 0], h0[g, 1])) for g in range(P)]
            hist_p = np.ascontiguousarray(h0.copy())
        ref = E.getphi_elem_pf(el, u, adiff.seed_D2(d), hist_o, grad_term=grad_term, nh2d_extension=ext)
        desc, params = X.mat_desc(mat, grad_term=grad_term, nh2d_extension=ext)
        phi, re, K = H.element_d(D, N, desc, params, el.gradN, el.wgt, el.Nval, u.T, d, hist_p)
        assert abs(phi - ref.v) <= TOL * max(1.0, abs(ref.v))
        assert _rel(re, adiff.grad(ref)) < TOL
        assert _rel(K, adiff.hess(ref)) < TOL
        if with_hist:
            for g in range(P):
                if neo:
                    assert hist_p[g, 0] == pytest.approx(hist_o[g], rel=1e-13, abs=1e-300)
                else:
                    assert hist_p[g, 0] == pytest.approx(hist_o[g][0], rel=1e-13, abs=1e-300)
                    assert hist_p[g, 1] == pytest.approx(hist_o[g][1], rel=1e-13, abs=1e-300)
