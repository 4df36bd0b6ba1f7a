"""ORACLE (test infrastructure only).

CPU restatement of the reference's element layer on the `makeϕrKt` path:
  /root/reference/src/elements.jl              (CEElem / CPElem data model, :71-111)
  /root/reference/src/elements.toolkit.jl      (getF :8-20, get_d_and_∇d :159-167,
                                                assembly :169-215, chain operator × :231-245)
  /root/reference/src/elasticelements.jl       (constructors :35-96,129-162,197; _integrate_ϕ :199-206;
                                                _assemble_ϕrKt :208-221; C3DE chain path :339-350)
  /root/reference/src/phasefieldelements.jl    (constructors :9-145; element energies :159-223;
                                                makeϕrKt(elems,u,d) :228-238; makeϕrKt_d :239-262)

All indices exposed by this module are 1-based like the reference (node ids, CSC
colptr/rowval); internal numpy arrays are 0-based.

parity status: PARITY UNPINNED BY REFERENCE TESTS (see oracle/adiff.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import adiff
from . import materials as M
from .adiff import D1, D2

GP2 = ((-0.577350269189626, 1.0), (0.577350269189626, 1.0))  # elasticelements.jl:52-53,131
GP1 = ((0.0, 1.0),)  # QuadR / Hex08R, elasticelements.jl:81,197


@dataclass
class Elem:
    """CEElem (kind='CE', elements.jl:71-77) or CPElem (kind='CP', elements.jl:104-111)."""
    kind: str
    D: int
    nodes: np.ndarray  # 1-based node ids, length N
    gradN: np.ndarray  # [D][P][N]   == elem.∇N[d][gp][a]
    wgt: np.ndarray  # [P]
    V: float
    mat: object
    Nval: np.ndarray | None = None  # [P][N]  (CPElem; CAS: the shape-function values N)
    X0: np.ndarray | None = None  # [P]  radius at the Gauss points (kind='CAS' only, elasticelements.jl:247, 273)

    @property
    def P(self):
        return len(self.wgt)

    @property
    def N(self):
        return len(self.nodes)


# ----------------------------------------------------------------------------
# constructors
# ----------------------------------------------------------------------------
def _det2(J):  # toolkit:33
    return J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]


def _det3(F):  # toolkit:34-36
    return (F[0, 0] * (F[1, 1] * F[2, 2] - F[1, 2] * F[2, 1])
            - F[0, 1] * (F[1, 0] * F[2, 2] - F[1, 2] * F[2, 0])
            + F[0, 2] * (F[1, 0] * F[2, 1] - F[1, 1] * F[2, 0]))


def _quad_shape(xi, eta):  # elasticelements.jl:57
    return [(1 - xi) * (1 - eta) / 4, (1 + xi) * (1 - eta) / 4, (1 + xi) * (1 + eta) / 4, (1 - xi) * (1 + eta) / 4]


def _hex_shape(x, e, z):  # elasticelements.jl:134-137
    return [(1 - x) * (1 - e) * (1 - z) / 8, (1 + x) * (1 - e) * (1 - z) / 8,
            (1 + x) * (1 + e) * (1 - z) / 8, (1 - x) * (1 + e) * (1 - z) / 8,
            (1 - x) * (1 - e) * (1 + z) / 8, (1 + x) * (1 - e) * (1 + z) / 8,
            (1 + x) * (1 + e) * (1 + z) / 8, (1 - x) * (1 + e) * (1 + z) / 8]


def _wdg_shape(x, e, z):  # phasefieldelements.jl:98-99
    return [(1 - z) * (1 - x - e) / 2, (1 - z) * x / 2, (1 - z) * e / 2,
            (1 + z) * (1 - x - e) / 2, (1 + z) * x / 2, (1 + z) * e / 2]


def _iso_point(shape, pt, p0):
    """One Gauss point of an isoparametric constructor (elasticelements.jl:66-74 / :148-157):
    N0 on D1 duals -> J[j,i] = ∂x_i/∂ξ_j -> ∇N = J \\ ∂N/∂ξ."""
    D = len(pt)
    N0 = shape(*adiff.seed_D1(np.asarray(pt, dtype=float)))
    p0 = np.asarray(p0, dtype=float)
    nn = len(N0)
    # p = Σ N0[a] p0[a]  (vector of D1)
    p = [None] * D
    for i in range(D):
        s = N0[0] * p0[0][i]
        for a in range(1, nn):
            s = s + N0[a] * p0[a][i]
        p[i] = s
    J = np.array([[p[i].g[j] for i in range(D)] for j in range(D)])
    dN = np.stack([adiff.grad(n) for n in N0], axis=1)  # D x N
    Nxyz = np.linalg.solve(J, dN)
    return np.array([n.v for n in N0]), Nxyz, J


def _tensor_gauss(D, GP):
    """Loop order `for ii (ξ), jj (η), kk (ζ)` stored [ii,jj,kk] and flattened column-major
    (tuple(Nx...)) -> ξ index fastest (elasticelements.jl:64-79, 144-161)."""
    n = len(GP)
    pts = []
    if D == 2:
        for jj in range(n):
            for ii in range(n):
                pts.append(((GP[ii][0], GP[jj][0]), GP[ii][1] * GP[jj][1]))
    else:
        for kk in range(n):
            for jj in range(n):
                for ii in range(n):
                    pts.append(((GP[ii][0], GP[jj][0], GP[kk][0]), (GP[ii][1] * GP[jj][1]) * GP[kk][1]))
    return pts


def _iso_elem(kind, D, shape, pts, nodes, p0, mat, det):
    P = len(pts)
    nn = len(nodes)
    gradN = np.empty((D, P, nn))
    Nval = np.empty((P, nn))
    wgt = np.empty(P)
    for g, (pt, w) in enumerate(pts):
        N0, Nxyz, J = _iso_point(shape, pt, p0)
        gradN[:, g, :] = Nxyz
        Nval[g] = N0
        wgt[g] = det(J) * w
    V = 0.0
    # V accumulates in loop order ii,jj(,kk) (ii OUTER) -- a different order from the storage order;
    # V is not on the makeϕrKt path, keep the simple sum.
    for g in range(P):
        V += wgt[g]
    return Elem(kind, D, np.asarray(nodes, dtype=np.int64), gradN, wgt, V, mat, Nval if kind == "CP" else None)


def Quad(nodes, p0, mat, GP=GP2):  # elasticelements.jl:50-80
    return _iso_elem("CE", 2, _quad_shape, _tensor_gauss(2, GP), nodes, p0, mat, _det2)


def QuadR(nodes, p0, mat):  # :81
    return Quad(nodes, p0, mat, GP=GP1)


def QuadP(nodes, p0, mat, GP=GP2):  # phasefieldelements.jl:25-54
    return _iso_elem("CP", 2, _quad_shape, _tensor_gauss(2, GP), nodes, p0, mat, _det2)


def QuadPR(nodes, p0, mat):  # :55
    return QuadP(nodes, p0, mat, GP=GP1)


def Hex08(nodes, p0, mat, GP=GP2):  # elasticelements.jl:129-162
    return _iso_elem("CE", 3, _hex_shape, _tensor_gauss(3, GP), nodes, p0, mat, _det3)


def Hex08R(nodes, p0, mat):  # :197
    return Hex08(nodes, p0, mat, GP=GP1)


def Hex08P(nodes, p0, mat, GP=GP2):  # phasefieldelements.jl:56-93
    return _iso_elem("CP", 3, _hex_shape, _tensor_gauss(3, GP), nodes, p0, mat, _det3)


def Hex08PR(nodes, p0, mat):  # :94
    return Hex08P(nodes, p0, mat, GP=GP1)


def Wdg06P(nodes, p0, mat):  # phasefieldelements.jl:95-129 (weights 1/3 as written)
    s = 3 ** 0.5 / 3
    pts = [((2 / 3, 1 / 6, s), 1 / 3), ((2 / 3, 1 / 6, -s), 1 / 3),
           ((1 / 6, 2 / 3, s), 1 / 3), ((1 / 6, 2 / 3, -s), 1 / 3),
           ((1 / 6, 1 / 6, s), 1 / 3), ((1 / 6, 1 / 6, -s), 1 / 3)]
    return _iso_elem("CP", 3, _wdg_shape, pts, nodes, p0, mat, np.linalg.det)


def _tria_data(p0):  # elasticelements.jl:38-46
    (x1, x2, x3) = (p0[0][0], p0[1][0], p0[2][0])
    (y1, y2, y3) = (p0[0][1], p0[1][1], p0[2][1])
    Delta = x1 * y2 - x2 * y1 - x1 * y3 + x3 * y1 + x2 * y3 - x3 * y2
    Nx = np.array([y2 - y3, y3 - y1, y1 - y2]) / Delta
    Ny = np.array([x3 - x2, x1 - x3, x2 - x1]) / Delta
    A = abs(Delta) / 2
    return Nx, Ny, A


def Tria(nodes, p0, mat):  # elasticelements.jl:35-49
    Nx, Ny, A = _tria_data(p0)
    return Elem("CE", 2, np.asarray(nodes, dtype=np.int64), np.stack([Nx[None], Ny[None]]), np.array([A]), A, mat)


def TriaP(nodes, p0, mat):  # phasefieldelements.jl:9-24
    Nx, Ny, A = _tria_data(p0)
    return Elem("CP", 2, np.asarray(nodes, dtype=np.int64), np.stack([Nx[None], Ny[None]]), np.array([A]), A, mat,
                np.array([[1 / 3, 1 / 3, 1 / 3]]))


# ---- axisymmetric constructors (elasticelements.jl:235-284).  They end in `CAS(nodes, N, Nx, Ny, X0, wgt, V, mat)`, a
# struct that does not exist upstream: only this data recipe is the reference's; kind='CAS' holds it.
def ASTria(nodes, p0, mat):  # elasticelements.jl:235-252
    (x1, x2, x3) = (p0[0][0], p0[1][0], p0[2][0])
    (y1, y2, y3) = (p0[0][1], p0[1][1], p0[2][1])
    Delta = x1 * y2 - x2 * y1 - x1 * y3 + x3 * y1 + x2 * y3 - x3 * y2
    N = np.array([1, 1, 1]) / 3
    Nx = np.array([y2 - y3, y3 - y1, y1 - y2]) / Delta
    Ny = np.array([x3 - x2, x1 - x3, x2 - x1]) / Delta
    A = abs(Delta) / 2
    X0 = (x1 + x2 + x3) / 3
    wgt = A * 2 * np.pi * X0
    return Elem("CAS", 2, np.asarray(nodes, dtype=np.int64), np.stack([Nx[None], Ny[None]]), np.array([wgt]), A, mat,
                N[None], np.array([X0]))


def ASQuad(nodes, p0, mat):  # elasticelements.jl:253-285
    r = np.array([-1, 1]) * 0.577350269189626
    p0 = np.asarray(p0, dtype=float)
    N0 = np.empty((2, 2), dtype=object)
    Nx = np.empty((2, 2), dtype=object)
    Ny = np.empty((2, 2), dtype=object)
    X0 = np.empty((2, 2))
    wgt = np.empty((2, 2))
    V = 0
    for ii, xi in enumerate(r):  # `for (ii, ξ) in enumerate(r), (jj, η) in enumerate(r)`: ii outer
        for jj, eta in enumerate(r):
            Nij = _quad_shape(*adiff.seed_D1(np.array([xi, eta])))
            p = [None, None]
            for i in range(2):
                s_ = Nij[0] * p0[0][i]
                for a in range(1, 4):
                    s_ = s_ + Nij[a] * p0[a][i]
                p[i] = s_
            J = np.array([[p[i].g[j] for i in range(2)] for j in range(2)])  # J[jj,ii] = p[ii].g[jj]
            Nxy = np.linalg.solve(J, np.stack([adiff.grad(n) for n in Nij], axis=1))
            N0[ii, jj] = np.array([n.v for n in Nij])
            Nx[ii, jj] = Nxy[0, :]
            Ny[ii, jj] = Nxy[1, :]
            X0[ii, jj] = p[0].v
            wgt[ii, jj] = _det2(J) * 2 * np.pi * p[0].v
            V += wgt[ii, jj]
    flat = lambda A: [A[ii, jj] for jj in range(2) for ii in range(2)]  # tuple(A...): column-major, ii fastest
    gradN = np.stack([np.stack(flat(Nx)), np.stack(flat(Ny))])
    return Elem("CAS", 2, np.asarray(nodes, dtype=np.int64), gradN, np.array(flat(wgt)), V, mat, np.stack(flat(N0)),
                np.array(flat(X0)))


def getF_cas(elem: Elem, u, gp: int):
    """EXTENSION (no reference counterpart): 3x3 F of an axisymmetric element, rows/columns (r, z, θ):
    in-plane block I + ∇u from (Nx, Ny), hoop stretch F33 = 1 + (N·u_r)/X0."""
    F = np.empty((3, 3), dtype=object)
    F[:, :] = 0.0
    for j in range(2):
        for k in range(2):
            F[j, k] = _dot(elem.gradN[k, gp], u[j, :])
    for j in range(2):
        F[j, j] = F[j, j] + 1
    F[2, 2] = _dot(elem.Nval[gp], u[0, :]) / elem.X0[gp] + 1
    return F


def _tet_data(p0):  # elasticelements.jl:85-94
    A = np.ones((4, 4))
    for i in range(4):
        A[i, 1:4] = p0[i]
    C = np.linalg.inv(A)
    V = np.linalg.det(A) / 6
    return V, C[1, :], C[2, :], C[3, :]


def Tet04(nodes, p0, mat):  # elasticelements.jl:82-96  (wgt = V signed)
    V, Nx, Ny, Nz = _tet_data(p0)
    return Elem("CE", 3, np.asarray(nodes, dtype=np.int64), np.stack([Nx[None], Ny[None], Nz[None]]),
                np.array([V]), V, mat)


def Tet04P(nodes, p0, mat):  # phasefieldelements.jl:130-145
    V, Nx, Ny, Nz = _tet_data(p0)
    return Elem("CP", 3, np.asarray(nodes, dtype=np.int64), np.stack([Nx[None], Ny[None], Nz[None]]),
                np.array([V]), V, mat, np.array([[0.25, 0.25, 0.25, 0.25]]))


# ----------------------------------------------------------------------------
# kinematics (toolkit:8-20, 159-167)
# ----------------------------------------------------------------------------
def _dot(a, b):
    """Elements.dot (elements.jl:22-28): s = zero(T); s += a[i]*b[i]."""
    s = 0.0
    for i in range(len(a)):
        s = s + a[i] * b[i]
    return s


def getF(elem: Elem, u, gp: int):
    """F[j,k] = δ_jk + ∇N[k][gp]·u[j,:]  (only rows 1:D of u are used, toolkit:13)."""
    D = elem.D
    F = np.empty((D, D), dtype=object)
    for j in range(D):
        for k in range(D):
            F[j, k] = _dot(elem.gradN[k, gp], u[j, :])
    for j in range(D):
        F[j, j] = F[j, j] + 1
    return F


def get_d_and_gradd(elem: Elem, d, gp: int, grad_term: str = "reference"):
    """toolkit:159-167.  `∇d` lives in an `@MVector zeros(T_elem, D)`; a dual assigned to it is
    truncated to its value (adiff.jl:73-74) -- SURVEY §9-5.  grad_term='intended' keeps the duals."""
    dv = _dot(elem.Nval[gp], d)
    gd = []
    for j in range(elem.D):
        x = _dot(elem.gradN[j, gp], d)
        gd.append(adiff.val(x) if grad_term == "reference" else x)
    return dv, gd


# ----------------------------------------------------------------------------
# chain operator ×  (toolkit:231-245)
# ----------------------------------------------------------------------------
def chain(phi: D2, F) -> D2:
    """`ϕ × F` with ϕ::D2{9,45}, F::3x3 array of D1{P}.  F is addressed linearly (column-major)."""
    Fl = [M.L(F, k) for k in range(1, F.size + 1)]
    n = phi.N
    P = Fl[0].N
    g = np.zeros(P)
    h = np.zeros((P + 1) * P // 2)
    for ii in range(n):
        g = g + phi.g[ii] * Fl[ii].g
    for ii in range(1, n):
        for jj in range(ii):
            hij = phi.h[adiff.packed_index(ii, jj)]
            h = h + hij * (adiff.outer(Fl[ii].g, Fl[jj].g) + adiff.outer(Fl[jj].g, Fl[ii].g))
    for ii in range(n):
        hii = phi.h[adiff.packed_index(ii, ii)]
        h = h + adiff.outer(hii * Fl[ii].g, Fl[ii].g)
    return D2(phi.v, g, h)


# ----------------------------------------------------------------------------
# element energies
# ----------------------------------------------------------------------------
def _integrate(elem: Elem, u):  # elasticelements.jl:199-206
    phi = 0.0
    for gp in range(elem.P):
        F = getF_cas(elem, u, gp) if elem.kind == "CAS" else getF(elem, u, gp)
        phi = phi + elem.wgt[gp] * M.getphi(F, elem.mat)
    return phi


def getphi_elem(elem: Elem, u):
    """`getϕ(elem, u)`; u is D x N (floats, or D2 duals seeded on u[:, nodes])."""
    u = np.asarray(u, dtype=object)
    is_d2 = isinstance(u.flat[0], D2)
    if elem.kind == "CE" and elem.D == 3 and is_d2:  # elasticelements.jl:339-350
        u1 = adiff.to_D1(u)
        n = u.size
        phi = D2(0.0, np.zeros(n), np.zeros((n + 1) * n // 2))
        for gp in range(elem.P):
            F = getF(elem, u1, gp)
            valF = adiff.val(F)
            dphi = M.getphi(adiff.seed_D2(valF), elem.mat)
            phi = phi + chain(elem.wgt[gp] * dphi, F)
        return phi
    return _integrate(elem, u)  # :327-329


def getphi_elem_pf(elem: Elem, u, d, hist=None, *, grad_term="reference", nh2d_extension=False):
    """`getϕ(elem::CPElem, u0, d0[, ϕmax])`  (phasefieldelements.jl:159-223).
    With `hist` (list over Gauss points, mutated in place like `ϕmax[ii]`, :180)."""
    u = np.asarray(u, dtype=object)
    d = np.asarray(d, dtype=object).reshape(-1)
    u_is_d2 = isinstance(u.flat[0], D2)
    if elem.D == 3 and u_is_d2:  # :194-223 chain path
        u1 = adiff.to_D1(u)
        n = u.size
        phi = D2(0.0, np.zeros(n), np.zeros((n + 1) * n // 2))
        for gp in range(elem.P):
            F = getF(elem, u1, gp)
            dv, gd = get_d_and_gradd(elem, d, gp, grad_term)
            valF = adiff.val(F)
            if hist is None:
                dphi = M.getphi_pf(adiff.seed_D2(valF), dv, gd, elem.mat)
            else:
                dphi, hist[gp] = M.getphi_pf(adiff.seed_D2(valF), dv, gd, elem.mat, hist[gp])
            phi = phi + chain(elem.wgt[gp] * dphi, F)
        return phi
    phi = 0.0
    for gp in range(elem.P):  # :159-184
        F = getF(elem, u, gp)
        dv, gd = get_d_and_gradd(elem, d, gp, grad_term)
        if hist is None:
            phi = phi + elem.wgt[gp] * M.getphi_pf(F, dv, gd, elem.mat, nh2d_extension=nh2d_extension)
        else:
            phii, hist[gp] = M.getphi_pf(F, dv, gd, elem.mat, hist[gp], nh2d_extension=nh2d_extension)
            phi = phi + elem.wgt[gp] * phii
    return phi


# ----------------------------------------------------------------------------
# assembly  (toolkit:169-203)
# ----------------------------------------------------------------------------
def assemble(Phi, elems, ushape):
    """`makeϕrKt(Φ, elems, u)`: returns (ϕ, r, (colptr, rowval, nzval)) with 1-based Int64 CSC
    arrays, duplicates summed as a left fold in `elems` order and exact zeros dropped
    (`dropzeros(sparse(II,JJ,Kt,N,N))`, :202)."""
    dpn = ushape[0]
    Ntot = int(np.prod(ushape))
    r = np.zeros(Ntot)
    phi = 0
    II, JJ, VV = [], [], []
    for e, elem in enumerate(elems):
        nodes = np.asarray(elem.nodes, dtype=np.int64)
        idx = ((nodes[None, :] - 1) * dpn + np.arange(1, dpn + 1)[:, None]).reshape(-1, order="F")  # :187
        phi = phi + Phi[e].v
        r[idx - 1] = r[idx - 1] + adiff.grad(Phi[e])
        n = len(idx)
        H = adiff.hess(Phi[e])
        II.append(np.repeat(idx[:, None], n, axis=1).reshape(-1, order="F"))  # idx * ones'
        JJ.append(np.repeat(idx[None, :], n, axis=0).reshape(-1, order="F"))  # ones * idx'
        VV.append(H.reshape(-1, order="F"))
    II = np.concatenate(II)
    JJ = np.concatenate(JJ)
    VV = np.concatenate(VV)
    return float(phi), r, coo_to_csc_leftfold(II, JJ, VV, Ntot)


def coo_to_csc_leftfold(II, JJ, VV, N, drop=True):
    """Julia `sparse(I,J,V,N,N)` (+ `dropzeros`): CSC, rows ascending within a column, duplicates
    combined with `+` in input order; stored ±0.0 removed."""
    key = (JJ - 1) * N + (II - 1)
    uniq, inv = np.unique(key, return_inverse=True)
    nz = np.zeros(len(uniq))
    np.add.at(nz, inv, VV)  # sequential, in input order
    rows = uniq % N + 1
    cols = uniq // N + 1
    if drop:
        keep = nz != 0
        rows, cols, nz = rows[keep], cols[keep], nz[keep]
    colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(cols - 1, minlength=N))]).astype(np.int64)
    return colptr, rows.astype(np.int64), nz


def assemble_r(Phi, elems, ushape):
    """`makeϕr` (toolkit:204-215)."""
    dpn = ushape[0]
    r = np.zeros(int(np.prod(ushape)))
    phi = 0
    for e, elem in enumerate(elems):
        nodes = np.asarray(elem.nodes, dtype=np.int64)
        idx = ((nodes[None, :] - 1) * dpn + np.arange(1, dpn + 1)[:, None]).reshape(-1, order="F")
        phi = phi + Phi[e].v
        r[idx - 1] = r[idx - 1] + adiff.grad(Phi[e])
    return float(phi), r


# ----------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------
def _cols(u, nodes):
    return np.asarray(u)[:, np.asarray(nodes) - 1]


def element_duals(elems, u):
    """Φ[ii] = getϕ(elems[ii], D2(u[:, nodes]))  (elasticelements.jl:216-218)."""
    assert len(elems) > 0, "makeϕrKt: `elems` is empty"  # :210
    return [getphi_elem(e, adiff.seed_D2(_cols(u, e.nodes))) for e in elems]


def make_phi_r_Kt(elems, u):
    """`makeϕrKt(elems, u)`  (elasticelements.jl:356-358)."""
    u = np.asarray(u, dtype=float)
    return assemble(element_duals(elems, u), elems, u.shape)


def make_phi_r_Kt_pf_u(elems, u, d, *, nh2d_extension=False):
    """`makeϕrKt(elems::Vector{<:CPElem}, u, d)` with the intended dual size D*N
    (phasefieldelements.jl:228-238)."""
    u = np.asarray(u, dtype=float)
    d = np.asarray(d, dtype=float).reshape(-1)
    Phi = [getphi_elem_pf(e, adiff.seed_D2(_cols(u, e.nodes)), d[e.nodes - 1], nh2d_extension=nh2d_extension)
           for e in elems]
    return assemble(Phi, elems, u.shape)


def make_phi_r_Kt_pf_d(elems, u, d, hist=None, *, grad_term="reference", nh2d_extension=False):
    """`makeϕrKt_d(elems, u, d[, ϕmax])` (phasefieldelements.jl:239-262).  `hist[e][gp]` is mutated."""
    u = np.asarray(u, dtype=float)
    d = np.asarray(d, dtype=float).reshape(-1)
    Phi = []
    for i, e in enumerate(elems):
        dd = adiff.seed_D2(d[e.nodes - 1])
        Phi.append(getphi_elem_pf(e, _cols(u, e.nodes), dd, None if hist is None else hist[i],
                                  grad_term=grad_term, nh2d_extension=nh2d_extension))
    return assemble(Phi, elems, (1, d.size))
