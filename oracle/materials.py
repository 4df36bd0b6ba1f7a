"""ORACLE (test infrastructure only).

CPU restatement of the reference's strain-energy densities `getϕ(F, mat)`:
  /root/reference/src/elasticmaterials.jl  (Hooke, Hooke2D, NeoHooke, MooneyRivlin, Ogden)
  /root/reference/src/materials.jl         (invariants, get1stinvariants, gethyddevdecomp)
  /root/reference/src/phasefieldmaterials.jl (PhaseField ATn / DHn, with and without history)

Works on plain floats and on oracle.adiff duals (object arrays).  Matrices are
numpy arrays; the reference's *linear* (column-major, 1-based) indexing `F[k]`
is reproduced by `L(F, k)`.

parity status: PARITY UNPINNED BY REFERENCE TESTS (see oracle/adiff.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import adiff
from .adiff import dlog, dsqrt


def L(A, k: int):
    """Julia linear index `A[k]` (1-based, column-major)."""
    A = np.asarray(A)
    r = A.shape[0]
    k -= 1
    return A[k % r, k // r]


def _eye_like(F):
    n = F.shape[0]
    return np.eye(n)


def _matmul_TN(F):
    """`transpose(F)F` with left-to-right accumulation over k (LinearAlgebra matmul2x2/3x3)."""
    n = F.shape[0]
    C = np.empty((n, n), dtype=object)
    for i in range(n):
        for j in range(n):
            s = F[0, i] * F[0, j]
            for k in range(1, n):
                s = s + F[k, i] * F[k, j]
            C[i, j] = s
    return C


# ----------------------------------------------------------------------------
# material structs (elasticmaterials.jl:25-120, phasefieldmaterials.jl:5-16)
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class Hooke:  # elasticmaterials.jl:25-31
    E: float
    nu: float
    rho: float = 1.0
    small: bool = False


@dataclass(frozen=True)
class Hooke2D:  # elasticmaterials.jl:65-73 (defaults small=true, plane_stress=true)
    E: float
    nu: float
    rho: float = 1.0
    small: bool = True
    plane_stress: bool = True


@dataclass(frozen=True)
class MooneyRivlin:  # elasticmaterials.jl:87-94
    C1: float
    C2: float
    K: float = -1.0
    small: bool = False


@dataclass(frozen=True)
class NeoHooke:  # elasticmaterials.jl:100-106
    C1: float
    K: float
    rho: float = 1.0
    small: bool = False


@dataclass(frozen=True)
class Ogden:  # elasticmaterials.jl:114-120 (constructor broken upstream)
    alpha: float
    mu: float
    K: float = -1.0


@dataclass(frozen=True)
class PhaseField:  # phasefieldmaterials.jl:5-16
    l0: float
    Gc: float
    mat: object
    n: float
    AT: str = "ATn"  # "ATn" | "DHn"


# ----------------------------------------------------------------------------
# invariants (materials.jl:98-114)
# ----------------------------------------------------------------------------
def getInvariants3(C):
    I1 = L(C, 1) + L(C, 5) + L(C, 9)
    I2 = L(C, 1) * L(C, 5) + L(C, 5) * L(C, 9) + L(C, 1) * L(C, 9) - L(C, 2) ** 2 - L(C, 3) ** 2 - L(C, 6) ** 2
    I3 = (
        L(C, 1) * L(C, 5) * L(C, 9)
        + 2 * L(C, 2) * L(C, 3) * L(C, 6)
        - L(C, 1) * L(C, 6) ** 2
        - L(C, 5) * L(C, 3) ** 2
        - L(C, 9) * L(C, 2) ** 2
    )
    return I1, I2, I3


def getInvariants2(C, C33):
    I1 = L(C, 1) + L(C, 4) + C33
    I2 = L(C, 1) * L(C, 4) + L(C, 4) * C33 + L(C, 1) * C33 - L(C, 2) ** 2
    I3 = L(C, 1) * L(C, 4) * C33 - C33 * L(C, 2) ** 2
    return I1, I2, I3


# ----------------------------------------------------------------------------
# strain measures shared by the Hooke family (elasticmaterials.jl:204-208 etc.)
# ----------------------------------------------------------------------------
def _strain(F, small: bool):
    n = F.shape[0]
    I = np.eye(n)
    E = np.empty((n, n), dtype=object)
    if small:
        for i in range(n):
            for j in range(n):
                E[i, j] = (F[i, j] + F[j, i] - 2 * I[i, j]) / 2
    else:
        C = _matmul_TN(F)
        for i in range(n):
            for j in range(n):
                E[i, j] = (C[i, j] - I[i, j]) / 2
    return E


def _lame(mat):
    nu, Es = mat.nu, mat.E
    lam = Es * nu / (1 + nu) / (1 - 2 * nu)
    mu = Es / 2 / (1 + nu)
    return lam, mu


# ----------------------------------------------------------------------------
# elastic energies
# ----------------------------------------------------------------------------
def getphi(F, mat):
    """`getϕ(F, mat)` dispatcher."""
    F = np.asarray(F, dtype=object)
    if isinstance(mat, Hooke):
        return _phi_hooke(F, mat)
    if isinstance(mat, Hooke2D):
        return _phi_hooke2d(F, mat)
    if isinstance(mat, (NeoHooke, MooneyRivlin)):
        return _phi_hyper(F, mat)
    if isinstance(mat, Ogden):
        return _phi_ogden(F, mat)
    raise TypeError(f"no getϕ for {type(mat)}")


def _phi_hooke(F, mat):  # elasticmaterials.jl:202-219
    E = _strain(F, mat.small)
    lam, mu = _lame(mat)
    phi = (mu + lam / 2) * (L(E, 1) ** 2 + L(E, 5) ** 2 + L(E, 9) ** 2)
    phi = phi + lam * (L(E, 1) * L(E, 5) + L(E, 5) * L(E, 9) + L(E, 9) * L(E, 1))
    phi = phi + 2 * mu * (L(E, 2) ** 2 + L(E, 3) ** 2 + L(E, 6) ** 2)
    return phi


def _phi_hooke2d(F, mat):  # elasticmaterials.jl:220-245
    E = _strain(F, mat.small)
    nu, Es = mat.nu, mat.E
    if not mat.plane_stress:
        lam, mu = _lame(mat)
        return (mu + lam / 2) * (L(E, 1) ** 2 + L(E, 4) ** 2) + lam * (L(E, 1) * L(E, 4)) + 2 * mu * L(E, 2) ** 2
    return (Es / (1 - nu ** 2)) * (L(E, 1) ** 2 + L(E, 4) ** 2 + nu * (L(E, 1) * L(E, 4))) + (Es / (1 + nu)) * L(E, 2) ** 2


def _phi_hyper(F, mat):  # elasticmaterials.jl:138-147
    C = _matmul_TN(F)
    if C.size == 9:
        I1, I2, I3 = getInvariants3(C)
    else:
        L3 = (L(F, 1) * L(F, 4) - L(F, 2) * L(F, 3)) ** -2 if mat.K < 0 else 1
        I1, I2, I3 = getInvariants2(C, L3)
    if isinstance(mat, MooneyRivlin):  # elasticmaterials.jl:173-185
        C1, C2, K = mat.C1, mat.C2, mat.K
        if K < 0:
            return C1 * (I1 - 3) + C2 * (I2 - 3)
        J = dsqrt(I3)
        I1 = I1 * J ** (-2 / 3)
        I2 = I2 * J ** (-4 / 3)
        return C1 * (I1 - 3) + C2 * (I2 - 3) + K * (J - 1) ** 2
    # NeoHooke, elasticmaterials.jl:186-197
    C1, K = mat.C1, mat.K
    if K < 0:
        return C1 * (I1 - 3)
    J = dsqrt(I3)
    I1 = I1 * J ** (-2 / 3)
    return C1 * (I1 - 3) + K * (J - 1) ** 2


def _phi_ogden(F, mat):  # elasticmaterials.jl:148-172 -- scalar energy only
    Fv = adiff.val(F).astype(float)
    alpha, mu, K = mat.alpha, mat.mu, mat.K
    if Fv.size == 9:
        C = Fv.T @ Fv
    else:
        F33 = 1 / (Fv[0, 0] * Fv[1, 1] - Fv[1, 0] * Fv[0, 1]) if K < 0 else 1.0
        F3 = np.zeros((3, 3))
        F3[:2, :2] = Fv
        F3[2, 2] = F33
        C = F3.T @ F3
    lam = np.sqrt(np.linalg.svd(C, compute_uv=False))
    if K < 0:
        return mu / alpha * (np.sum(lam ** alpha) - 3)
    J = np.prod(lam)
    return mu / alpha * (np.sum(lam ** alpha) / J ** (alpha / 3) - 3) + K * (J - 1) ** 2


# ----------------------------------------------------------------------------
# helpers for the phase-field energies
# ----------------------------------------------------------------------------
def get1stinvariants(F, mat):
    """materials.jl:143-155 (3-D), elasticmaterials.jl:306-333 (Hooke2D)."""
    E = _strain(F, mat.small)
    if isinstance(mat, Hooke2D):
        if not mat.plane_stress:
            I1 = L(E, 1) + L(E, 4)
            I1sq = L(E, 1) ** 2 + L(E, 4) ** 2 + 2 * L(E, 3) ** 2
            return I1, I1sq
        E33 = mat.nu / (mat.nu - 1) * (L(E, 1) + L(E, 4))
        I1 = L(E, 1) + L(E, 4) + E33
        I1sq = L(E, 1) ** 2 + L(E, 4) ** 2 + E33 ** 2 + 2 * L(E, 3) ** 2
        return I1, I1sq
    I1 = L(E, 1) + L(E, 5) + L(E, 9)
    I1sq = L(E, 1) ** 2 + L(E, 5) ** 2 + L(E, 9) ** 2 + 2 * (L(E, 2) ** 2 + L(E, 3) ** 2 + L(E, 6) ** 2)
    return I1, I1sq


def _ddot(A, B):
    """Elements.dot(::SMatrix, ::SMatrix): linear-order left fold (elements.jl:30-36)."""
    n = A.shape[0]
    s = 0.0
    for k in range(1, A.size + 1):
        s = s + L(A, k) * L(B, k)
    return s


def gethyddevdecomp(F, mat):
    """materials.jl:117-129 (3-D), elasticmaterials.jl:264-291 (Hooke2D)."""
    E = _strain(F, mat.small)
    n = F.shape[0]
    I = np.eye(n)
    if isinstance(mat, Hooke2D):
        if not mat.plane_stress:
            I1 = L(E, 1) + L(E, 4)
            ed = np.empty((n, n), dtype=object)
            for i in range(n):
                for j in range(n):
                    ed[i, j] = E[i, j] - I[i, j] * I1 / 3
            return I1, _ddot(ed, ed) + (I1 / 3) ** 2
        E33 = mat.nu / (mat.nu - 1) * (L(E, 1) + L(E, 4))
        I1 = L(E, 1) + L(E, 4) + E33
        ed = np.empty((n, n), dtype=object)
        for i in range(n):
            for j in range(n):
                ed[i, j] = E[i, j] - I[i, j] * I1 / 3
        return I1, _ddot(ed, ed) + (E33 - I1 / 3) ** 2
    I1 = L(E, 1) + L(E, 5) + L(E, 9)
    ed = np.empty((n, n), dtype=object)
    for i in range(n):
        for j in range(n):
            ed[i, j] = E[i, j] - I[i, j] * I1 / 3
    return I1, _ddot(ed, ed)


def _vdot(a, b):
    s = a[0] * b[0]
    for k in range(1, len(a)):
        s = s + a[k] * b[k]
    return s


# ----------------------------------------------------------------------------
# phase-field energies (phasefieldmaterials.jl:20-135)
# ----------------------------------------------------------------------------
def getphi_pf(F, d, gd, mat: PhaseField, hist=None, *, nh2d_extension=False):
    """`getϕ(F, d, ∇d, mat::PhaseField [, ϕmax])`.

    Returns ϕ, or (ϕ, new_hist) when `hist` is given.

    `hist` is a 2-tuple for the Hooke-type bases (phasefieldmaterials.jl:75,117) and a scalar
    for the NeoHooke base (:94 -- unreachable upstream; kept as the documented intent).
    `nh2d_extension`: plane-strain embedding of the 3-D NeoHooke formula for 2x2 F.
    """
    F = np.asarray(F, dtype=object)
    l0, Gc, n = mat.l0, mat.Gc, mat.n
    base = mat.mat
    gamma = d ** n + l0 ** 2 * _vdot(gd, gd)
    if isinstance(base, NeoHooke) and mat.AT == "ATn":  # :37-54, :94-115
        C1, K = base.C1, base.K
        if F.size == 9:
            f = lambda k: L(F, k)
            J = (
                f(1) * f(5) * f(9) - f(1) * f(6) * f(8) - f(2) * f(4) * f(9)
                + f(2) * f(6) * f(7) + f(3) * f(4) * f(8) - f(3) * f(5) * f(7)
            )
            I1 = (f(1) ** 2 + f(2) ** 2 + f(3) ** 2 + f(4) ** 2 + f(5) ** 2 + f(6) ** 2
                  + f(7) ** 2 + f(8) ** 2 + f(9) ** 2)
        else:
            if not nh2d_extension:
                raise IndexError("PhaseField{NeoHooke} indexes F[5..9]: BoundsError for 2x2 F upstream (SURVEY §9-3)")
            f = lambda k: L(F, k)
            J = f(1) * f(4) - f(2) * f(3)
            I1 = f(1) ** 2 + f(2) ** 2 + f(3) ** 2 + f(4) ** 2 + 1
        if hist is None:
            phiel = C1 * (I1 - 3 - 2 * dlog(J)) + K * (J - 1) ** 2
            psi = (1 - d) ** 2 * phiel if J >= 1 else phiel
            return psi + Gc / (2 * l0) * gamma
        if J >= 1:
            hist = max(hist, adiff.val(C1 * (I1 - 3 - 2 * dlog(J)) + K * (J - 1) ** 2))
            psi = (1 - d) ** 2 * hist
        else:
            psi = C1 * (I1 - 3 - 2 * dlog(J)) + K * (J - 1) ** 2
        return psi + Gc / (2 * l0) * gamma, hist

    lam, mu = _lame(base)
    if mat.AT == "ATn":  # :20-36, :75-93
        I1, I1sq = get1stinvariants(F, base)
        if hist is None:
            if I1 >= 0:
                psi = (1 - d) ** 2 * (lam / 2 * I1 ** 2 + mu * I1sq)
            else:
                psi = lam / 2 * I1 ** 2 + (1 - d) ** 2 * mu * I1sq
            return psi + Gc / (2 * l0) * gamma
        if I1 >= 0:
            hist = (_hmax(lam / 2 * I1 ** 2 + mu * I1sq, hist[0]), hist[1])
            psi = (1 - d) ** 2 * hist[0]
        else:
            hist = (hist[0], _hmax(mu * I1sq, hist[1]))
            psi = lam / 2 * I1 ** 2 + (1 - d) ** 2 * hist[1]
        return psi + Gc / (2 * l0) * gamma, hist
    if mat.AT == "DHn":  # :56-72, :117-135
        I1, ed = gethyddevdecomp(F, base)
        K = lam + mu / 3
        if hist is None:
            if I1 >= 0:
                psi = (1 - d) ** 2 * (K / 2 * I1 ** 2 + mu * ed)
            else:
                psi = K / 2 * I1 ** 2 + (1 - d) ** 2 * mu * ed
            return psi + Gc / (2 * l0) * gamma
        if I1 >= 0:
            hist = (_hmax(K / 2 * I1 ** 2 + mu * ed, hist[0]), hist[1])
            psi = (1 - d) ** 2 * hist[0]
        else:
            hist = (hist[0], _hmax(mu * ed, hist[1]))
            psi = lam / 2 * I1 ** 2 + (1 - d) ** 2 * hist[1]  # λ/2 (not K/2) as written, :131
        return psi + Gc / (2 * l0) * gamma, hist
    raise ValueError(mat.AT)


def _hmax(a, b):
    """`max(a, b)` on the history path: u is plain Float64 there (phasefieldelements.jl:246,259),
    so both arguments are floats."""
    a = adiff.val(a)
    b = adiff.val(b)
    return a if a > b else b
