"""Host-side mirror of `AD4SM.Elements` for the `makeϕrKt` path.

Same names, argument meaning and error behaviour as the reference (file:line into /root/reference/src):
  element data model   CEElem / CPElem                elements.jl:71-77, 104-111
  constructors         Tria, Quad, QuadR, Tet04, Hex08, Hex08R            elasticelements.jl:35-96, 129-162, 197
                       TriaP, QuadP, QuadPR, Hex08P, Hex08PR, Wdg06P, Tet04P   phasefieldelements.jl:9-145
  hot path             makeϕrKt(elems, u)             elasticelements.jl:356-358
                       makeϕrKt(elems, u, d)          phasefieldelements.jl:228-238
                       makeϕrKt_d(elems, u, d[, ϕmax])  phasefieldelements.jl:239-262
                       makeϕr(elems, u)               elements.toolkit.jl:204-215 over the loop of :216-218

The constructors (model-build time, host, numpy) produce exactly the per-element data the reference
stores -- physical shape-function gradients ∇N[d][gp][a], weights wgt[gp] = det(J)·w, and N[gp][a]
for CPElem -- either one element at a time (reference signature) or vectorised over a whole mesh
(`*_batch`).  All energy/gradient/Hessian evaluation and the assembly run on the GPU through the
C ABI (`_capi.py` -> libad4sm_b200.so); nothing here computes them on the CPU.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field

import numpy as np

from . import Materials
from . import _capi as capi

__all__ = [
    "CEElem", "CPElem", "ElemBatch", "MeshBatch", "SparseMatrixCSC", "Plan",
    "Tria", "Quad", "QuadR", "Tet04", "Hex08", "Hex08R",
    "TriaP", "QuadP", "QuadPR", "Hex08P", "Hex08PR", "Wdg06P", "Tet04P",
    "CAS", "ASTria", "ASQuad",
    "makeϕrKt", "makeϕrKt_d", "makeϕr", "make_phi_r_Kt", "make_phi_r_Kt_d", "make_phi_r",
]

GP2 = ((-0.577350269189626, 1.0), (0.577350269189626, 1.0))  # elasticelements.jl:52-53,131 (truncated √3/3)
GP1 = ((0.0, 1.0),)


# ------------------------------------------------------------------------------------------------
# data model
# ------------------------------------------------------------------------------------------------
@dataclass
class ElemBatch:
    """Structure-of-arrays packing of a homogeneous `Vector{CEElem}` / `Vector{CPElem}`: one element
    type (kind, D, P, N) and one material.  Layouts follow the reference's nesting order."""
    kind: str  # "CE" | "CP" | "CAS" (axisymmetric: D = 2, Nval = N, X0 = radius at the Gauss points)
    D: int
    nodes: np.ndarray  # [ne][N] int64, 1-based
    gradN: np.ndarray  # [ne][D][P][N]
    wgt: np.ndarray  # [ne][P]
    mat: object
    Nval: np.ndarray | None = None  # [ne][P][N]  (CP)
    V: np.ndarray | None = None  # [ne]
    orig_index: np.ndarray | None = None  # [ne] position in `elems` (summation order); None = in order
    halo: bool = False  # mesh-partitioned runs: elements evaluated by another rank (dist.py); gradN/wgt may be None
    halo_P: int = 0  # number of Gauss points of a halo batch without wgt
    X0: np.ndarray | None = None  # [ne][P]  (CAS)

    def __post_init__(self):
        self.nodes = np.ascontiguousarray(self.nodes, dtype=np.int64)
        if self.halo and self.gradN is None:
            return
        self.gradN = np.ascontiguousarray(self.gradN, dtype=np.float64)
        self.wgt = np.ascontiguousarray(self.wgt, dtype=np.float64)
        if self.Nval is not None:
            self.Nval = np.ascontiguousarray(self.Nval, dtype=np.float64)
        ne, N = self.nodes.shape
        assert self.gradN.shape[0] == ne and self.gradN.shape[1] == self.D and self.gradN.shape[3] == N
        assert self.wgt.shape == (ne, self.gradN.shape[2])

    @property
    def ne(self):
        return self.nodes.shape[0]

    @property
    def N(self):
        return self.nodes.shape[1]

    @property
    def P(self):
        return self.wgt.shape[1] if self.wgt is not None else self.halo_P

    def __len__(self):
        return self.ne

    def __getitem__(self, i):
        cls = {"CE": CEElem, "CP": CPElem, "CAS": CAS}[self.kind]
        kw = dict(nodes=self.nodes[i].copy(), gradN=self.gradN[i], wgt=self.wgt[i],
                  V=float(self.V[i]) if self.V is not None else float(self.wgt[i].sum()), mat=self.mat, D=self.D)
        if self.kind != "CE":
            kw["N"] = self.Nval[i]
        if self.kind == "CAS":
            kw["X0"] = self.X0[i]
        return cls(**kw)

    def hoop_row(self):
        """CAS: N[gp][a] / X0[gp] -- ∂F33/∂u_r, what the C ABI takes as `Nval` for AD4SM_ELEM_CAS."""
        return np.ascontiguousarray(self.Nval / self.X0[:, :, None])


_SHAPES = {  # shape -> (capi code, D, N, tensor rule?)
    "Tria": (capi.SHAPE_TRIA3, 2, 3, False), "Quad": (capi.SHAPE_QUAD4, 2, 4, True), "Tet04": (capi.SHAPE_TET4, 3, 4, False),
    "Hex08": (capi.SHAPE_HEX8, 3, 8, True), "Wdg06": (capi.SHAPE_WDG6, 3, 6, False),
}


@dataclass
class MeshBatch:
    """A homogeneous group of elements given by connectivity + node coordinates: what the reference constructors
    (`Hex08(nodes, p0; GP, mat)` ...) take, for all elements at once.  ∇N, wgt (and N) are constructed ON THE DEVICE
    (csrc/construct.cu) -- when a Plan is built from it (no host ∇N, no upload), or explicitly into
    host arrays by `materialize()`, which returns the `ElemBatch` the host constructors (`*_batch`) would."""
    kind: str  # "CE" | "CP" | "CAS" (axisymmetric: D = 2, Nval = N, X0 = radius at the Gauss points)
    shape: str  # key of _SHAPES
    nodes: np.ndarray  # [ne][N] int64, 1-based
    coords: np.ndarray  # [n_nodes][>= D]
    mat: object
    GP: tuple = ()  # tensor shapes: ((xi, w), ...) -- the reference's GP keyword
    orig_index: np.ndarray | None = None
    halo = False
    gradN = wgt = Nval = None

    def __post_init__(self):
        self.nodes = np.ascontiguousarray(self.nodes, dtype=np.int64)
        self.coords = np.ascontiguousarray(self.coords, dtype=np.float64)
        code, D, N, tensor = _SHAPES[self.shape]
        if self.nodes.ndim != 2 or self.nodes.shape[1] != N:
            raise ValueError(f"{self.shape}: connectivity must be [ne][{N}]")
        if self.coords.ndim != 2 or self.coords.shape[1] < D:
            raise ValueError(f"{self.shape}: coordinates must be [n_nodes][>= {D}]")
        if tensor and not 1 <= len(self.GP) <= 3:
            raise ValueError("GP must hold 1..3 (abscissa, weight) pairs")
        self.D = D

    @property
    def ne(self):
        return self.nodes.shape[0]

    @property
    def N(self):
        return self.nodes.shape[1]

    @property
    def P(self):
        code, D, N, tensor = _SHAPES[self.shape]
        return len(self.GP) ** D if tensor else (6 if self.shape == "Wdg06" else 1)

    def __len__(self):
        return self.ne

    def c_mesh(self):
        m = capi.Mesh()
        m.shape = _SHAPES[self.shape][0]
        m.n_gp_1d = len(self.GP)
        for i, (xi, w) in enumerate(self.GP):
            m.gp_xi[i], m.gp_w[i] = xi, w
        m.coord_dim = self.coords.shape[1]
        m.n_nodes = self.coords.shape[0]
        m.coords = self.coords.ctypes.data
        return m

    def materialize(self, device=-1):
        """Run the batched constructor on the device and return the result as a host `ElemBatch`."""
        ne, N, P, D = self.ne, self.N, self.P, self.D
        gradN, wgt, V = np.empty((ne, D, P, N)), np.empty((ne, P)), np.empty(ne)
        Nval = np.empty((ne, P, N)) if self.kind == "CP" else None
        m = self.c_mesh()
        capi.check(capi.lib().ad4sm_construct(C.byref(m), ne, self.nodes.ctypes.data, device, gradN.ctypes.data, wgt.ctypes.data,
                                             Nval.ctypes.data if Nval is not None else None, V.ctypes.data))
        return ElemBatch(self.kind, D, self.nodes, gradN, wgt, self.mat, Nval, V=V, orig_index=self.orig_index)


@dataclass
class CEElem:
    """`CEElem{D,P,M,T,N}` (elements.jl:71-77).  `gradN[d][gp][a]` is the reference's `∇N`."""
    nodes: np.ndarray
    gradN: np.ndarray  # [D][P][N]
    wgt: np.ndarray  # [P]
    V: float
    mat: object
    D: int = 0
    kind = "CE"

    def __post_init__(self):
        self.nodes = np.asarray(self.nodes, dtype=np.int64)
        self.gradN = np.asarray(self.gradN, dtype=np.float64)
        self.wgt = np.asarray(self.wgt, dtype=np.float64)
        self.D = self.gradN.shape[0]

    @property
    def ᐁN(self):  # noqa: N802  (∇ itself is not an identifier character in Python)
        return self.gradN

    @property
    def P(self):
        return self.wgt.shape[0]

    @property
    def Nn(self):
        return self.nodes.shape[0]


@dataclass
class CPElem(CEElem):
    """`CPElem{D,P,M,T,N}` (elements.jl:104-111): adds the shape-function values `N[gp][a]`."""
    N: np.ndarray = None  # [P][N]
    kind = "CP"

    def __post_init__(self):
        super().__post_init__()
        self.N = np.asarray(self.N, dtype=np.float64)


@dataclass
class CAS(CPElem):
    """The axisymmetric continuum element `CAS(nodes, N, Nx, Ny, X0, wgt, V, mat)` that `ASTria` / `ASQuad` construct
    (elasticelements.jl:235-284).  Upstream the struct itself and its `getϕ` are missing; here it is a 2-D
    element with dofs (u_r, u_z), `gradN = (Nx, Ny)`, the shape-function values `N` and the radius `X0` per Gauss point,
    evaluated with a 3x3 F (in-plane block + hoop stretch 1 + (N·u_r)/X0) and the 3-D materials -- an extension."""
    X0: np.ndarray = None  # [P]
    kind = "CAS"

    def __post_init__(self):
        super().__post_init__()
        self.X0 = np.asarray(self.X0, dtype=np.float64)


@dataclass
class SparseMatrixCSC:
    """Julia's `SparseMatrixCSC{Float64,Int64}`: 1-based `colptr`/`rowval`, rows ascending per column."""
    m: int
    n: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray

    @property
    def nnz(self):
        return int(self.nzval.shape[0])

    @property
    def shape(self):
        return (self.m, self.n)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))


# ------------------------------------------------------------------------------------------------
# constructors (vectorised over elements; the per-element reference signatures wrap them)
# ------------------------------------------------------------------------------------------------
def _shape_quad(xi, eta):  # elasticelements.jl:57; returns N [4], dN/dξ [2][4]
    N = np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)]) / 4
    dN = np.array([[-(1 - eta), (1 - eta), (1 + eta), -(1 + eta)],
                   [-(1 - xi), -(1 + xi), (1 + xi), (1 - xi)]]) / 4
    return N, dN


def _shape_hex(x, e, z):  # elasticelements.jl:134-137
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1.0])
    se = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1.0])
    fx, fe, fz = 1 + sx * x, 1 + se * e, 1 + sz * z
    N = fx * fe * fz / 8
    dN = np.array([sx * fe * fz, fx * se * fz, fx * fe * sz]) / 8
    return N, dN


def _shape_wdg(x, e, z):  # phasefieldelements.jl:98-99
    t = 1 - x - e
    N = np.array([(1 - z) * t, (1 - z) * x, (1 - z) * e, (1 + z) * t, (1 + z) * x, (1 + z) * e]) / 2
    dN = np.array([[-(1 - z), (1 - z), 0, -(1 + z), (1 + z), 0],
                   [-(1 - z), 0, (1 - z), -(1 + z), 0, (1 + z)],
                   [-t, -x, -e, t, x, e]]) / 2
    return N, dN


def _tensor_points(D, GP):
    """Gauss points in storage order: `Nx[ii,jj(,kk)]` flattened column-major by `tuple(Nx...)`, i.e. the
    ξ index runs fastest (elasticelements.jl:64-79, 144-161)."""
    pts = []
    if D == 2:
        for (eta, we) in GP:
            for (xi, wx) in GP:
                pts.append(((xi, eta), wx * we))
    else:
        for (ze, wz) in GP:
            for (eta, we) in GP:
                for (xi, wx) in GP:
                    pts.append(((xi, eta, ze), wx * we * wz))
    return pts


def _det(J):
    """detJ (toolkit:33-36) over a stack of D x D matrices."""
    if J.shape[-1] == 2:
        return J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    return (J[..., 0, 0] * (J[..., 1, 1] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 1])
            - J[..., 0, 1] * (J[..., 1, 0] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 0])
            + J[..., 0, 2] * (J[..., 1, 0] * J[..., 2, 1] - J[..., 1, 1] * J[..., 2, 0]))


def _iso_batch(kind, D, shape, pts, conn, coords, mat, chunk=65536):
    """Isoparametric constructor over all elements: J[j,i] = ∂x_i/∂ξ_j, ∇N = J \\ ∂N/∂ξ, wgt = det(J)·w
    (elasticelements.jl:66-74, 148-157)."""
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    coords = np.asarray(coords, dtype=np.float64)
    ne, N = conn.shape
    P = len(pts)
    gradN = np.empty((ne, D, P, N))
    wgt = np.empty((ne, P))
    Nval = np.empty((ne, P, N)) if kind == "CP" else None
    shp = [shape(*pt) for pt, _ in pts]
    for s in range(0, ne, chunk):
        sl = slice(s, min(ne, s + chunk))
        X = coords[conn[sl] - 1][:, :, :D]  # [e][a][i]
        for g, (pt, w) in enumerate(pts):
            N0, dN = shp[g]
            J = np.einsum("ja,eai->eji", dN, X)
            gradN[sl, :, g, :] = np.linalg.solve(J, np.broadcast_to(dN, (J.shape[0],) + dN.shape))
            wgt[sl, g] = _det(J) * w
            if Nval is not None:
                Nval[sl, g, :] = N0
    return ElemBatch(kind, D, conn, gradN, wgt, mat, Nval, V=wgt.sum(axis=1))


def _tria_batch(kind, conn, coords, mat):  # elasticelements.jl:38-46
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    X = np.asarray(coords, dtype=np.float64)[conn - 1]
    x1, x2, x3 = X[:, 0, 0], X[:, 1, 0], X[:, 2, 0]
    y1, y2, y3 = X[:, 0, 1], X[:, 1, 1], X[:, 2, 1]
    Delta = x1 * y2 - x2 * y1 - x1 * y3 + x3 * y1 + x2 * y3 - x3 * y2
    Nx = np.stack([y2 - y3, y3 - y1, y1 - y2], axis=1) / Delta[:, None]
    Ny = np.stack([x3 - x2, x1 - x3, x2 - x1], axis=1) / Delta[:, None]
    A = np.abs(Delta) / 2
    gradN = np.stack([Nx, Ny], axis=1)[:, :, None, :]
    Nval = np.full((len(conn), 1, 3), 1 / 3) if kind == "CP" else None
    return ElemBatch(kind, 2, conn, gradN, A[:, None], mat, Nval, V=A)


def _tet_batch(kind, conn, coords, mat):  # elasticelements.jl:85-94 (wgt = V = det(A)/6, signed)
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    X = np.asarray(coords, dtype=np.float64)[conn - 1]
    A = np.ones((len(conn), 4, 4))
    A[:, :, 1:4] = X[:, :, :3]
    Cm = np.linalg.inv(A)
    V = np.linalg.det(A) / 6
    gradN = Cm[:, 1:4, :][:, :, None, :]
    Nval = np.full((len(conn), 1, 4), 0.25) if kind == "CP" else None
    return ElemBatch(kind, 3, conn, gradN, V[:, None], mat, Nval, V=V)


_WDG_PTS = [((2 / 3, 1 / 6, 3 ** 0.5 / 3), 1 / 3), ((2 / 3, 1 / 6, -(3 ** 0.5) / 3), 1 / 3),
            ((1 / 6, 2 / 3, 3 ** 0.5 / 3), 1 / 3), ((1 / 6, 2 / 3, -(3 ** 0.5) / 3), 1 / 3),
            ((1 / 6, 1 / 6, 3 ** 0.5 / 3), 1 / 3), ((1 / 6, 1 / 6, -(3 ** 0.5) / 3), 1 / 3)]


# `device=True`: return a MeshBatch instead -- the same constructor runs on the GPU when the Plan is built (or on
# `.materialize()`), and the host never holds ∇N.
def Tria_batch(conn, coords, mat, device=False):
    return MeshBatch("CE", "Tria", conn, coords, mat) if device else _tria_batch("CE", conn, coords, mat)


def TriaP_batch(conn, coords, mat, device=False):
    return MeshBatch("CP", "Tria", conn, coords, mat) if device else _tria_batch("CP", conn, coords, mat)


def Quad_batch(conn, coords, mat, GP=GP2, device=False):
    if device:
        return MeshBatch("CE", "Quad", conn, coords, mat, GP)
    return _iso_batch("CE", 2, _shape_quad, _tensor_points(2, GP), conn, coords, mat)


def QuadP_batch(conn, coords, mat, GP=GP2, device=False):
    if device:
        return MeshBatch("CP", "Quad", conn, coords, mat, GP)
    return _iso_batch("CP", 2, _shape_quad, _tensor_points(2, GP), conn, coords, mat)


def Tet04_batch(conn, coords, mat, device=False):
    return MeshBatch("CE", "Tet04", conn, coords, mat) if device else _tet_batch("CE", conn, coords, mat)


def Tet04P_batch(conn, coords, mat, device=False):
    return MeshBatch("CP", "Tet04", conn, coords, mat) if device else _tet_batch("CP", conn, coords, mat)


def Hex08_batch(conn, coords, mat, GP=GP2, device=False):
    if device:
        return MeshBatch("CE", "Hex08", conn, coords, mat, GP)
    return _iso_batch("CE", 3, _shape_hex, _tensor_points(3, GP), conn, coords, mat)


def Hex08P_batch(conn, coords, mat, GP=GP2, device=False):
    if device:
        return MeshBatch("CP", "Hex08", conn, coords, mat, GP)
    return _iso_batch("CP", 3, _shape_hex, _tensor_points(3, GP), conn, coords, mat)


def Wdg06_batch(conn, coords, mat, device=False):  # elasticelements.jl:163-197
    return MeshBatch("CE", "Wdg06", conn, coords, mat) if device else _iso_batch("CE", 3, _shape_wdg, _WDG_PTS, conn, coords, mat)


def Wdg06P_batch(conn, coords, mat, device=False):
    return MeshBatch("CP", "Wdg06", conn, coords, mat) if device else _iso_batch("CP", 3, _shape_wdg, _WDG_PTS, conn, coords, mat)


def ASTria_batch(conn, coords, mat):  # elasticelements.jl:235-252
    b = _tria_batch("CP", conn, coords, mat)   # N = [1,1,1]/3, Nx, Ny, A as in Tria
    X = np.asarray(coords, dtype=np.float64)[b.nodes - 1]
    X0 = ((X[:, 0, 0] + X[:, 1, 0] + X[:, 2, 0]) / 3)[:, None]
    A = b.wgt
    return ElemBatch("CAS", 2, b.nodes, b.gradN, A * 2 * np.pi * X0, mat, b.Nval, V=A[:, 0], X0=X0)


def ASQuad_batch(conn, coords, mat):  # elasticelements.jl:253-285 (2x2 rule, wgt = detJ·2π·r, V = Σ wgt)
    b = _iso_batch("CP", 2, _shape_quad, _tensor_points(2, GP2), conn, coords, mat)
    X = np.asarray(coords, dtype=np.float64)[b.nodes - 1]
    X0 = np.einsum("epa,ea->ep", b.Nval, X[:, :, 0])   # p[1].v = Σ_a N_a x_a, accumulated in node order
    wgt = b.wgt * 2 * np.pi * X0
    return ElemBatch("CAS", 2, b.nodes, b.gradN, wgt, mat, b.Nval, V=wgt.sum(axis=1), X0=X0)


def _single(batch_ctor, nodes, p0, **kw):
    nodes = np.asarray(nodes, dtype=np.int64)
    p0 = np.asarray(p0, dtype=np.float64)
    b = batch_ctor(np.arange(1, len(nodes) + 1)[None, :], p0, **kw)
    el = b[0]
    el.nodes = nodes
    return el


def _default_mat(mat):
    return Materials.Hooke(1.0, 0.3) if mat is None else mat


# reference signatures: Ctor(nodes, p0; mat=..., GP=...)
def Tria(nodes, p0, mat=None):
    return _single(Tria_batch, nodes, p0, mat=_default_mat(mat))


def ASTria(nodes, p0, mat=None):
    return _single(ASTria_batch, nodes, p0, mat=_default_mat(mat))


def ASQuad(nodes, p0, mat=None):
    return _single(ASQuad_batch, nodes, p0, mat=_default_mat(mat))


def Quad(nodes, p0, GP=GP2, mat=None):
    return _single(Quad_batch, nodes, p0, mat=_default_mat(mat), GP=GP)


def QuadR(nodes, p0, mat=None):
    return Quad(nodes, p0, GP=GP1, mat=mat)


def Tet04(nodes, p0, mat=None):
    return _single(Tet04_batch, nodes, p0, mat=_default_mat(mat))


def Hex08(nodes, p0, GP=GP2, mat=None):
    return _single(Hex08_batch, nodes, p0, mat=_default_mat(mat), GP=GP)


def Hex08R(nodes, p0, mat=None):
    return Hex08(nodes, p0, GP=GP1, mat=mat)


def TriaP(nodes, p0, mat=None):
    return _single(TriaP_batch, nodes, p0, mat=_default_mat(mat))


def QuadP(nodes, p0, GP=GP2, mat=None):
    return _single(QuadP_batch, nodes, p0, mat=_default_mat(mat), GP=GP)


def QuadPR(nodes, p0, mat=None):
    return QuadP(nodes, p0, GP=GP1, mat=mat)


def Hex08P(nodes, p0, GP=GP2, mat=None):
    return _single(Hex08P_batch, nodes, p0, mat=_default_mat(mat), GP=GP)


def Hex08PR(nodes, p0, mat=None):
    return Hex08P(nodes, p0, GP=GP1, mat=mat)


def Wdg06P(nodes, p0, mat=None):
    return _single(Wdg06P_batch, nodes, p0, mat=_default_mat(mat))


def Tet04P(nodes, p0, mat=None):
    return _single(Tet04P_batch, nodes, p0, mat=_default_mat(mat))


# ------------------------------------------------------------------------------------------------
# packing `elems` -> batches
# ------------------------------------------------------------------------------------------------
def material_descriptor(mat, pf_grad_term="reference", nh2d_extension=False):
    """(mat_kind, mat_base, mat_flags, params[8]) of include/ad4sm_b200.h from a Materials object."""
    p = np.zeros(8)

    def elastic(m):
        q = np.zeros(4)
        f = 0
        if isinstance(m, Materials.Hooke):
            k, q[:2] = capi.MAT_HOOKE, (m.E, m.nu)
            f |= capi.F_SMALL if m.small else 0
        elif isinstance(m, Materials.Hooke2D):
            k, q[:2] = capi.MAT_HOOKE2D, (m.E, m.nu)
            f |= (capi.F_SMALL if m.small else 0) | (capi.F_PLANE_STRESS if m.plane_stress else 0)
        elif isinstance(m, Materials.NeoHooke):
            k, q[:2] = capi.MAT_NEOHOOKE, (m.C1, m.K)
        elif isinstance(m, Materials.MooneyRivlin):
            k, q[:3] = capi.MAT_MOONEYRIVLIN, (m.C1, m.C2, m.K)
        elif isinstance(m, Materials.Ogden):
            k, q[:3] = capi.MAT_OGDEN, (m.alpha, m.mu, m.K)
        else:
            raise TypeError(f"MethodError: no method matching getϕ(F, ::{type(m).__name__})")
        return k, q, f

    if isinstance(mat, Materials.PhaseField):
        base, q, f = elastic(mat.mat)
        p[:4] = q
        p[4:7] = (mat.l0, mat.Gc, mat.n)
        f |= capi.F_DHN if mat.AT == "DHn" else 0
        if pf_grad_term == "intended":
            f |= capi.F_PF_GRAD_INTENDED
        elif pf_grad_term != "reference":
            raise ValueError("pf_grad_term must be 'reference' or 'intended'")
        if nh2d_extension:
            f |= capi.F_NH2D_EXT
        return capi.MAT_PHASEFIELD, base, f, p
    kind, q, f = elastic(mat)
    p[:4] = q
    return kind, 0, f, p


def as_batches(elems):
    """`elems` (a list of CEElem/CPElem, an ElemBatch, or a list of ElemBatch) -> list[ElemBatch].
    Element objects are grouped by (kind, D, P, N, material); `orig_index` keeps the position in `elems`
    so the assembly sums duplicates in `elems` order like the reference."""
    if isinstance(elems, (ElemBatch, MeshBatch)):
        return [elems]
    if len(elems) == 0:
        raise AssertionError("makeϕrKt: `elems` is empty")  # elasticelements.jl:210
    if all(isinstance(e, (ElemBatch, MeshBatch)) for e in elems):
        return list(elems)
    groups = {}
    for i, e in enumerate(elems):
        if not isinstance(e, CEElem):
            raise TypeError(f"MethodError: makeϕrKt does not accept {type(e).__name__}")
        key = (e.kind, e.D, e.P, e.Nn, id(e.mat) if not getattr(e.mat, "__hash__", None) else e.mat)
        groups.setdefault(key, []).append(i)
    out = []
    single = len(groups) == 1
    for (kind, D, P, N, _), idx in groups.items():
        es = [elems[i] for i in idx]
        out.append(ElemBatch(kind, D, np.stack([e.nodes for e in es]), np.stack([e.gradN for e in es]),
                             np.stack([e.wgt for e in es]), es[0].mat,
                             np.stack([e.N for e in es]) if kind != "CE" else None,
                             V=np.array([e.V for e in es]),
                             orig_index=None if single else np.asarray(idx, dtype=np.int64),
                             X0=np.stack([e.X0 for e in es]) if kind == "CAS" else None))
    return out


# ------------------------------------------------------------------------------------------------
# plan = device-resident copy of `elems` + pattern + scatter maps
# ------------------------------------------------------------------------------------------------
class Plan:
    """Built once per (`elems`, size(u)); evaluates makeϕrKt / makeϕrKt_d many times (Newton loop)."""

    def __init__(self, elems, n_nodes, dofs_per_node, device=-1, pf_grad_term="reference", nh2d_extension=False):
        self.batches = as_batches(elems)
        self.n_nodes = int(n_nodes)
        self.dpn = int(dofs_per_node)
        L = capi.lib()
        arr = (capi.Batch * len(self.batches))()
        self._keep = []
        for i, b in enumerate(self.batches):
            kind, base, flags, params = material_descriptor(b.mat, pf_grad_term, nh2d_extension)
            s = arr[i]
            s.elem_kind = {"CE": capi.ELEM_CE, "CP": capi.ELEM_CP, "CAS": capi.ELEM_CAS}[b.kind]
            s.dim, s.n_gauss, s.n_elem_nodes = b.D, b.P, b.N
            s.mat_kind, s.mat_base, s.mat_flags = kind, base, flags
            for k in range(8):
                s.mat_params[k] = params[k]
            s.n_elems = b.ne
            s.batch_flags = capi.BATCH_HALO if b.halo else 0
            s.nodes = b.nodes.ctypes.data
            if isinstance(b, MeshBatch):  # constructed on the device inside ad4sm_plan_create
                m = b.c_mesh()
                self._keep.append(m)
                s.mesh = C.pointer(m)
            s.gradN = b.gradN.ctypes.data if b.gradN is not None else None
            s.wgt = b.wgt.ctypes.data if b.wgt is not None else None
            if b.kind == "CAS" and not b.halo:   # the C ABI takes the hoop row N / X0
                hr = b.hoop_row()
                self._keep.append(hr)
                s.Nval = hr.ctypes.data
            else:
                s.Nval = b.Nval.ctypes.data if b.Nval is not None else None
            if b.orig_index is not None:
                oi = np.ascontiguousarray(b.orig_index, dtype=np.int64)
                self._keep.append(oi)
                s.orig_index = oi.ctypes.data
            else:
                s.orig_index = None
        h = C.c_void_p()
        capi.check(L.ad4sm_plan_create(arr, len(self.batches), self.n_nodes, self.dpn, device, C.byref(h)))
        self._h = h
        self._L = L
        self._fin = weakref.finalize(self, L.ad4sm_plan_destroy, h)
        self.has_cp = any(b.kind == "CP" for b in self.batches)
        self.ne = sum(b.ne for b in self.batches)

    def close(self):
        self._fin()

    # -- pattern ---------------------------------------------------------------------------------
    def pattern_size(self, field=capi.FIELD_U):
        n, z = C.c_int64(), C.c_int64()
        capi.check(self._L.ad4sm_plan_pattern_size(self._h, field, C.byref(n), C.byref(z)))
        return n.value, z.value

    def pattern(self, field=capi.FIELD_U):
        n, z = self.pattern_size(field)
        colptr = np.empty(n + 1, dtype=np.int64)
        rowval = np.empty(z, dtype=np.int64)
        capi.check(self._L.ad4sm_plan_pattern(self._h, field, colptr.ctypes.data, rowval.ctypes.data))
        return colptr, rowval

    # -- host API (reference-facing) --------------------------------------------------------------
    def _u(self, u):
        u = np.asarray(u, dtype=np.float64)
        if u.ndim != 2 or u.shape != (self.dpn, self.n_nodes):
            raise ValueError(f"DimensionMismatch: u must be {self.dpn}x{self.n_nodes}, got {u.shape}")
        return np.asfortranarray(u)

    def _d(self, d):
        d = np.ascontiguousarray(np.asarray(d, dtype=np.float64).reshape(-1))
        if d.size != self.n_nodes:
            raise ValueError(f"DimensionMismatch: d must have {self.n_nodes} entries")
        return d

    def _fetch(self, n_dof, nnz):
        colptr = np.empty(n_dof + 1, dtype=np.int64)
        rowval = np.empty(nnz, dtype=np.int64)
        nzval = np.empty(nnz, dtype=np.float64)
        capi.check(self._L.ad4sm_fetch_csc(self._h, colptr.ctypes.data, rowval.ctypes.data, nzval.ctypes.data))
        return SparseMatrixCSC(n_dof, n_dof, colptr, rowval, nzval)

    def makeϕrKt(self, u, d=None, fetch=True):
        u = self._u(u)
        phi, nnz = C.c_double(), C.c_int64()
        r = np.empty(u.size)
        if d is None:
            capi.check(self._L.ad4sm_eval(self._h, u.ctypes.data, C.byref(phi), r.ctypes.data, C.byref(nnz)))
        else:
            d = self._d(d)
            capi.check(self._L.ad4sm_eval_pf_u(self._h, u.ctypes.data, d.ctypes.data, C.byref(phi), r.ctypes.data,
                                               C.byref(nnz)))
        return phi.value, r, (self._fetch(u.size, nnz.value) if fetch else None)

    def assemble_duals(self, Phi, fetch=True):
        """makeϕrKt(Φ, elems, u) (elements.toolkit.jl:169-203): assemble caller-supplied element duals.  `Phi` is a
        C-contiguous float64 array [ne][stride >= 1+n+m] whose rows are (v | g[n] | packed upper-triangle h[m]) -- the
        memory of Julia's Vector{D2{n,m,Float64}} -- in `elems` order."""
        Phi = np.ascontiguousarray(Phi, dtype=np.float64)
        if Phi.ndim != 2 or Phi.shape[0] != sum(b.ne for b in self.batches):
            raise ValueError("Phi must have one row per element")
        phi, nnz = C.c_double(), C.c_int64()
        ndof = self.n_nodes * self.dpn
        r = np.empty(ndof)
        capi.check(self._L.ad4sm_assemble_duals(self._h, C.c_void_p(Phi.ctypes.data), Phi.shape[1], C.byref(phi),
                                                C.c_void_p(r.ctypes.data), C.byref(nnz)))
        return phi.value, r, (self._fetch(ndof, nnz.value) if fetch else None)

    def assemble_duals_phi_r(self, Phi):
        """makeϕr(Φ, elems, u) (elements.toolkit.jl:204-215): rows (v | g | ...), D1 or D2 records."""
        Phi = np.ascontiguousarray(Phi, dtype=np.float64)
        if Phi.ndim != 2 or Phi.shape[0] != sum(b.ne for b in self.batches):
            raise ValueError("Phi must have one row per element")
        phi = C.c_double()
        r = np.empty(self.n_nodes * self.dpn)
        capi.check(self._L.ad4sm_assemble_duals_phi_r(self._h, C.c_void_p(Phi.ctypes.data), Phi.shape[1], C.byref(phi),
                                                      C.c_void_p(r.ctypes.data)))
        return phi.value, r

    def makeϕr(self, u):
        u = self._u(u)
        phi = C.c_double()
        r = np.empty(u.size)
        capi.check(self._L.ad4sm_eval_phi_r(self._h, u.ctypes.data, C.byref(phi), r.ctypes.data))
        return phi.value, r

    def makeϕrKt_d(self, u, d, ϕmax=None, fetch=True):
        u = self._u(u)
        d = self._d(d)
        phi, nnz = C.c_double(), C.c_int64()
        r = np.empty(self.n_nodes)
        hp = None
        hist = None
        if ϕmax is not None:
            hist = self._split_hist(ϕmax)
            hp = (C.c_void_p * len(hist))(*[h.ctypes.data for h in hist])
        capi.check(self._L.ad4sm_eval_pf_d(self._h, u.ctypes.data, d.ctypes.data, hp, C.byref(phi), r.ctypes.data,
                                           C.byref(nnz)))
        if ϕmax is not None:
            self._merge_hist(ϕmax, hist)
        return phi.value, r, (self._fetch(self.n_nodes, nnz.value) if fetch else None)

    def _split_hist(self, hmax):
        """ϕmax[e][gp] = (ϕmax⁺, ϕmax⁻): one contiguous [ne_b][P][2] array per batch (views when possible)."""
        if len(self.batches) == 1 and self.batches[0].orig_index is None:
            b = self.batches[0]
            if not (isinstance(hmax, np.ndarray) and hmax.dtype == np.float64 and hmax.flags.c_contiguous
                    and hmax.shape == (b.ne, b.P, 2)):
                raise ValueError(f"ϕmax must be a C-contiguous float64 array of shape ({b.ne}, {b.P}, 2)")
            return [hmax]
        out = []
        off = 0
        for b in self.batches:
            idx = b.orig_index if b.orig_index is not None else np.arange(off, off + b.ne)
            out.append(np.ascontiguousarray(np.stack([np.asarray(hmax[i], dtype=np.float64).reshape(b.P, 2) for i in idx])))
            off += b.ne
        return out

    def _merge_hist(self, hmax, hist):
        if len(hist) == 1 and hist[0] is hmax:
            return
        off = 0
        for b, h in zip(self.batches, hist):
            idx = b.orig_index if b.orig_index is not None else np.arange(off, off + b.ne)
            for k, i in enumerate(idx):
                hmax[i][...] = h[k]
            off += b.ne

    # -- Newton-step glue on the device (Solvers.solvestep!, solvers.jl:183-184, 211-224, 257-291) ---------------
    def newton_setup(self, bfree, field=capi.FIELD_U):
        """`bfree` = bfreeu (bool, shaped like u or flat in u's column-major order).  Returns (nfree, ncnst, nnz_ff)."""
        b = np.asarray(bfree)
        b = np.ascontiguousarray(b.reshape(-1, order="F"), dtype=np.uint8)
        n, _ = self.pattern_size(field)
        if b.size != n:
            raise ValueError(f"DimensionMismatch: bfree must have {n} entries")
        a, c, z = C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(self._L.ad4sm_newton_setup(self._h, field, b.ctypes.data, C.byref(a), C.byref(c), C.byref(z)))
        self._newton = {field: (a.value, c.value, z.value)}
        self._newton_pat = {}
        return a.value, c.value, z.value

    def newton_pattern(self, field=capi.FIELD_U):
        nfree, _, nnz = self._newton[field]
        if field not in self._newton_pat:
            cp = np.empty(nfree + 1, dtype=np.int64)
            rv = np.empty(nnz, dtype=np.int64)
            capi.check(self._L.ad4sm_newton_pattern(self._h, field, cp.ctypes.data, rv.ctypes.data))
            self._newton_pat[field] = (cp, rv)
        return self._newton_pat[field]

    def newton_reduce(self, kind, fe=None, ducnst=None, want_K=True, field=capi.FIELD_U, out=None):
        """After makeϕrKt on this plan: (res, Kt[ifree,ifree] or None, norm(res), norm(fi[icnst])).
        kind 0 = predictor (res = fi[ifree] - fe[ifree] - Kt[ifree,icnst]*ducnst), 1 = corrector (res = fe[ifree] - fi[ifree]).
        `out`: optional preallocated (res, nzval_ff) arrays (e.g. pinned)."""
        nfree, ncnst, nnz = self._newton[field]
        res = out[0] if out else np.empty(nfree)
        nz = (out[1] if out else np.empty(nnz)) if want_K else None
        norms = np.empty(2)
        fe_p = None
        if fe is not None:
            fe = np.ascontiguousarray(np.asarray(fe, dtype=np.float64).reshape(-1, order="F"))
            fe_p = fe.ctypes.data
        du_p = None
        if ducnst is not None:
            ducnst = np.ascontiguousarray(ducnst, dtype=np.float64)
            if ducnst.size != ncnst:
                raise ValueError(f"DimensionMismatch: ducnst must have {ncnst} entries")
            du_p = ducnst.ctypes.data
        capi.check(self._L.ad4sm_newton_reduce(self._h, field, int(kind), fe_p, du_p, res.ctypes.data,
                                               nz.ctypes.data if want_K else None, norms.ctypes.data))
        Kff = None
        if want_K:
            cp, rv = self.newton_pattern(field)
            Kff = SparseMatrixCSC(nfree, nfree, cp, rv, nz)
        return res, Kff, float(norms[0]), float(norms[1])

    def newton_reduce_device(self, kind, fe_ptr=None, du_ptr=None, field=capi.FIELD_U):
        capi.check(self._L.ad4sm_newton_reduce_device(self._h, field, int(kind), C.c_void_p(fe_ptr) if fe_ptr else None,
                                                      C.c_void_p(du_ptr) if du_ptr else None))

    def newton_view(self, field=capi.FIELD_U):
        v = capi.NewtonView()
        capi.check(self._L.ad4sm_newton_view_get(self._h, field, C.byref(v)))
        return v

    # -- batched post-processing (elements.toolkit.jl:8-132) -------------------------------------------------------
    def post(self, u, F=False, detJ=False, V=False, sigma=False, P=False):
        """Per-element post-processing in `elems` order; returns a dict of the requested arrays:
        F [ne, D, D], detJ [ne], V [ne] (current volumes), sigma [ne, D, D], P [ne, D, D] (the average of P1K')."""
        u = self._u(u)
        D = self.batches[0].D
        out, ptr = {}, []
        for name, want, shape in (("F", F, (self.ne, D * D)), ("detJ", detJ, (self.ne,)), ("V", V, (self.ne,)),
                                  ("sigma", sigma, (self.ne, D * D)), ("P", P, (self.ne, D * D))):
            if want:
                out[name] = np.empty(shape)
                ptr.append(C.c_void_p(out[name].ctypes.data))
            else:
                ptr.append(None)
        capi.check(self._L.ad4sm_post(self._h, C.c_void_p(u.ctypes.data), *ptr))
        for k in ("F", "sigma", "P"):
            if k in out:   # column-major D x D per element, like Julia's SMatrix
                out[k] = out[k].reshape(self.ne, D, D).transpose(0, 2, 1)
        return out

    # -- device-resident API (bench / GPU-side callers) ------------------------------------------
    def set_stream(self, cuda_stream):
        capi.check(self._L.ad4sm_set_stream(self._h, C.c_void_p(cuda_stream)))

    def eval_device(self, mode, u_ptr, d_ptr=None, hist_ptr=None):
        capi.check(self._L.ad4sm_eval_device(self._h, mode, C.c_void_p(u_ptr), C.c_void_p(d_ptr) if d_ptr else None,
                                             C.c_void_p(hist_ptr) if hist_ptr else None))

    def eval_elements_device(self, mode, u_ptr, d_ptr=None, hist_ptr=None):
        """K1 only (non-halo batches): element blocks / residuals / energies -> staging."""
        capi.check(self._L.ad4sm_eval_elements_device(self._h, mode, C.c_void_p(u_ptr), C.c_void_p(d_ptr) if d_ptr else None,
                                                      C.c_void_p(hist_ptr) if hist_ptr else None))

    def eval_elements_host(self, mode, u, d=None):
        """Element stage with host inputs (numpy, column-major D x nNodes like `makeϕrKt`'s u): asynchronous upload --
        chunked and overlapped with K1 on large plans -- then K1.  The arrays must stay untouched until `sync()`."""
        u = np.asarray(u)
        ok = u.dtype == np.float64 and u.ndim == 2 and (
            (u.shape == (self.dpn, self.n_nodes) and u.flags.f_contiguous) or      # the reference's u, column-major
            (u.shape == (self.n_nodes, self.dpn) and u.flags.c_contiguous))        # its C-ordered transpose: same bytes
        if not ok:
            raise ValueError(f"u must be float64 memory laid out [node][component]: a column-major {self.dpn} x {self.n_nodes} "
                             f"matrix or its C-ordered transpose, got shape {u.shape} (no hidden copy on the asynchronous path; a "
                             "C-ordered D x nNodes array would be read transposed)")
        dp = None
        if d is not None:
            d = np.asarray(d)
            if d.dtype != np.float64 or d.size != self.n_nodes or not (d.flags.c_contiguous or d.flags.f_contiguous):
                raise ValueError(f"d must be a contiguous float64 array with {self.n_nodes} entries")
            dp = C.c_void_p(d.ctypes.data)
        capi.check(self._L.ad4sm_eval_elements_host(self._h, mode, C.c_void_p(u.ctypes.data), dp, None))

    def fetch_phi_r_async(self, r_out, phi_out=None, field=capi.FIELD_U):
        """Enqueue the device-to-host copy of r (and ϕ, a 1-element float64 array) of the last assembly; it starts as soon
        as they are final and overlaps the Kt assembly.  Valid after `sync()`."""
        if r_out.dtype != np.float64 or not r_out.flags.c_contiguous and not r_out.flags.f_contiguous:
            raise ValueError("r_out must be a contiguous float64 array")
        capi.check(self._L.ad4sm_fetch_phi_r_async(self._h, field, C.c_void_p(phi_out.ctypes.data) if phi_out is not None else None,
                                                   C.c_void_p(r_out.ctypes.data)))

    def eval_assemble_device(self, mode):
        """K2/K3 + ϕ reduction over the staging of all batches (after any interface exchange)."""
        capi.check(self._L.ad4sm_eval_assemble_device(self._h, mode))

    def eval_assemble_interior_device(self, mode):
        """First half of a partitioned run's assembly (ad4sm_eval_assemble_split_device): the rows and node pairs no HALO
        slot contributes to; may run while the interface rows are still in flight.  `eval_assemble_device` then finishes."""
        capi.check(self._L.ad4sm_eval_assemble_split_device(self._h, mode))

    def eval_halo_prepare_device(self, mode, cuda_stream):
        """Behind the interface exchange, on its stream: HALO rows into the sorted staging + the interface rows of r
        (ad4sm_eval_halo_prepare_device), overlapping the interior assembly on the plan's stream."""
        capi.check(self._L.ad4sm_eval_halo_prepare_device(self._h, mode, C.c_void_p(cuda_stream)))

    def staging_view(self, field=capi.FIELD_U):
        v = capi.StagingView()
        capi.check(self._L.ad4sm_staging_view_get(self._h, field, C.byref(v)))
        return v

    def set_option(self, option, value):
        capi.check(self._L.ad4sm_set_option(self._h, option, int(value)))

    def info(self, what):
        v = C.c_int64()
        capi.check(self._L.ad4sm_get_info(self._h, what, C.byref(v)))
        return v.value

    def sync(self):
        capi.check(self._L.ad4sm_sync(self._h))

    def device_view(self, field=capi.FIELD_U):
        v = capi.DeviceView()
        capi.check(self._L.ad4sm_device_view_get(self._h, field, C.byref(v)))
        return v

    def fetch_csc(self, field=capi.FIELD_U):
        n, z = self.pattern_size(field)
        # nnz after dropzeros is only known to the library: over-allocate to the structural size
        colptr = np.empty(n + 1, dtype=np.int64)
        rowval = np.empty(z, dtype=np.int64)
        nzval = np.empty(z, dtype=np.float64)
        capi.check(self._L.ad4sm_fetch_csc(self._h, colptr.ctypes.data, rowval.ctypes.data, nzval.ctypes.data))
        k = int(colptr[-1] - 1)
        return SparseMatrixCSC(n, n, colptr, rowval[:k], nzval[:k])

    def profile_enable(self, on=True):
        capi.check(self._L.ad4sm_profile_enable(self._h, int(on)))

    def profile_read(self, reset=True):
        ms = (C.c_double * capi.PROF_NCLASS)()
        n = (C.c_int64 * capi.PROF_NCLASS)()
        capi.check(self._L.ad4sm_profile_read(self._h, ms, n, int(reset)))
        return list(ms), list(n)


# ------------------------------------------------------------------------------------------------
# reference-named entry points
# ------------------------------------------------------------------------------------------------
_plans: dict = {}


def _plan_for(elems, u, **opts):
    """Plan cache keyed on the identity of `elems` (like the Julia wrapper's `objectid(elems)`)."""
    u = np.asarray(u)
    if u.ndim != 2:
        raise ValueError("u must be a D x nNodes matrix")
    key = (id(elems), u.shape, tuple(sorted(opts.items())))
    hit = _plans.get(key)
    if hit is not None and hit[0]() is elems:
        return hit[1]
    plan = Plan(elems, u.shape[1], u.shape[0], **opts)
    try:
        ref = weakref.ref(elems, lambda _r, k=key, d=_plans: d.pop(k, None) if d is not None else None)
    except TypeError:  # plain lists cannot be weakly referenced: keep at most a few plans alive
        holder = elems
        ref = lambda h=holder: h  # noqa: E731
        while len(_plans) >= 4:
            _plans.pop(next(iter(_plans)))
    _plans[key] = (ref, plan)
    return plan


def makeϕrKt(elems, u, d=None, **opts):
    """`makeϕrKt(elems, u)` -> (ϕ, r, Kt)  and  `makeϕrKt(elems, u, d)` (phase-field u-phase)."""
    return _plan_for(elems, u, **opts).makeϕrKt(u, d)


def makeϕrKt_d(elems, u, d, ϕmax=None, **opts):
    """`makeϕrKt_d(elems, u, d)` / `makeϕrKt_d(elems, u, d, ϕmax)`: ϕmax is updated in place."""
    return _plan_for(elems, u, **opts).makeϕrKt_d(u, d, ϕmax)


def makeϕr(elems, u, **opts):
    return _plan_for(elems, u, **opts).makeϕr(u)


def assemble_ϕrKt(Φ, elems, u, **opts):
    """The reference's 3-argument `makeϕrKt(Φ, elems, u)` (elements.toolkit.jl:169-203; what `Solvers.solvestep!` calls,
    solvers.jl:213-216): Φ = element duals, rows (v | g | packed h) like a Julia Vector{D2{n,m,Float64}}; only the
    assembly runs on the device."""
    return _plan_for(elems, u, **opts).assemble_duals(Φ)


def pack_duals(val, grad, hess):
    """(val[ne], grad[ne][n], hess[ne][n][n]) -> Φ rows (v | g | h packed upper triangle, entry (i<=j) at j(j+1)/2+i,
    0-based: adiff.jl:26,84)."""
    val, grad, hess = np.asarray(val), np.asarray(grad), np.asarray(hess)
    n = grad.shape[1]
    jj, ii = np.triu_indices(n)          # pairs (ii >= jj) ...
    order = np.lexsort((jj, ii))          # ... sorted by column ii, then row jj: index ii(ii+1)/2 + jj
    rows, cols = jj[order], ii[order]
    return np.ascontiguousarray(np.concatenate([val[:, None], grad, hess[:, rows, cols]], axis=1))


# ---- post-processing over `elems` (elements.toolkit.jl:10, 28, 47, 59, 88-113): the reference's array methods ----------
def getF(elems, u, **opts):
    """`getF(elems, u)`: one D x D volume-averaged deformation gradient per element."""
    return _plan_for(elems, u, **opts).post(u, F=True)["F"]


def get_grad_u(elems, u, **opts):
    """`get∇u(elems, u)` (∇ is not a legal Python identifier character)."""
    F = getF(elems, u, **opts)
    return F - np.eye(F.shape[1])[None]


def detJ(elems, u, **opts):
    return _plan_for(elems, u, **opts).post(u, detJ=True)["detJ"]


def getV(elems, u, **opts):
    """`getV(elems, u)` = sum(getV(elem, u[:,elem.nodes]) for elem in elems): left fold in `elems` order."""
    total = 0.0
    for v in _plan_for(elems, u, **opts).post(u, V=True)["V"]:
        total += v
    return total


def getσ(elems, u, **opts):
    """`getσ(elems::Vector{<:C3D}, u)` -> 6 x nElems (xx yy zz xy zx yz), `::Vector{<:C2D}` -> 3 x nElems (xx yy xy)."""
    s = _plan_for(elems, u, **opts).post(u, sigma=True)["sigma"]
    D = s.shape[1]
    flat = s.transpose(0, 2, 1).reshape(len(s), D * D)        # column-major linear index, like σii[idx]
    idx = [0, 4, 8, 1, 2, 5] if D == 3 else [0, 3, 1]           # [1,5,9,2,3,6] / [1,4,2], toolkit:92, 106
    return np.ascontiguousarray(flat[:, idx].T)


def getP(elems, u, **opts):
    return _plan_for(elems, u, **opts).post(u, P=True)["P"]


get_sigma = getσ
make_phi_r_Kt = makeϕrKt
make_phi_r_Kt_d = makeϕrKt_d
make_phi_r = makeϕr
assemble_phi_r_Kt = assemble_ϕrKt
