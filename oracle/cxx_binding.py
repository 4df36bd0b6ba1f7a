"""ORACLE (test infrastructure only): ctypes binding of oracle/cxx/ad4sm_oracle.cpp.

Builds `oracle/_build/libad4sm_oracle.so` on demand (g++ -fopenmp, see oracle/Makefile) and
exposes the reference-algorithm restatement for medium/large meshes and as the timed CPU
baseline ("port").  Takes raw SoA arrays so it does not depend on the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libad4sm_oracle.so")
_lib = None

HOOKE, HOOKE2D, NEOHOOKE, MOONEY, OGDEN, PHASEFIELD = 1, 2, 3, 4, 5, 6
F_SMALL, F_PLANE_STRESS, F_DHN, F_PF_GRAD_INTENDED, F_NH2D_EXT = 1, 2, 4, 8, 16
M_ELASTIC, M_PF_U, M_PF_D = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "cxx", "ad4sm_oracle.cpp")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_assemble.restype = C.c_void_p
        _lib.oracle_result_nnz.restype = C.c_int64
        _lib.oracle_result_nnz.argtypes = [C.c_void_p]
        _lib.oracle_result_copy.argtypes = [C.c_void_p] * 4
        _lib.oracle_result_free.argtypes = [C.c_void_p]
    return _lib


def max_threads():
    return lib().oracle_max_threads()


def mat_desc(mat, grad_term="reference", nh2d_extension=False):
    """(kind, base, flags) int32[3] and params double[8] from a material object (duck-typed on the
    class name so both oracle.materials and the product's Materials classes work)."""
    name = type(mat).__name__
    p = np.zeros(8)
    flags = 0
    base = 0

    def elastic(m):
        nm = type(m).__name__
        q = np.zeros(4)
        f = 0
        if nm == "Hooke":
            k = HOOKE
            q[:2] = (m.E, m.nu)
            f |= F_SMALL if m.small else 0
        elif nm == "Hooke2D":
            k = HOOKE2D
            q[:2] = (m.E, m.nu)
            f |= (F_SMALL if m.small else 0) | (F_PLANE_STRESS if m.plane_stress else 0)
        elif nm == "NeoHooke":
            k = NEOHOOKE
            q[:2] = (m.C1, m.K)
        elif nm == "MooneyRivlin":
            k = MOONEY
            q[:3] = (m.C1, m.C2, m.K)
        elif nm == "Ogden":
            k = OGDEN
            q[:3] = (m.alpha, m.mu, m.K)
        else:
            raise TypeError(nm)
        return k, q, f

    if name == "PhaseField":
        kind = PHASEFIELD
        base, q, f = elastic(mat.mat)
        p[:4] = q
        p[4:7] = (mat.l0, mat.Gc, mat.n)
        flags = f | (F_DHN if mat.AT == "DHn" else 0)
        if grad_term == "intended":
            flags |= F_PF_GRAD_INTENDED
        if nh2d_extension:
            flags |= F_NH2D_EXT
    else:
        kind, q, flags = elastic(mat)
        p[:4] = q
    return np.array([kind, base, flags], dtype=np.int32), p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def elem_duals(kind, D, P, N, mat, nodes, gradN, wgt, Nval, mode, u, d=None, hist=None, nthreads=0,
               grad_term="reference", nh2d_extension=False):
    """Φ for one batch: returns (val[ne], grad[ne,n], hess[ne,n,n]); `hist` [ne,P,2] is updated in place."""
    L = lib()
    nodes = np.ascontiguousarray(nodes, dtype=np.int64)
    gradN = np.ascontiguousarray(gradN, dtype=np.float64)
    wgt = np.ascontiguousarray(wgt, dtype=np.float64)
    Nval = None if Nval is None else np.ascontiguousarray(Nval, dtype=np.float64)
    u = np.asfortranarray(u, dtype=np.float64)
    dpn = u.shape[0]
    d = None if d is None else np.ascontiguousarray(np.asarray(d, dtype=np.float64).reshape(-1))
    ne = nodes.shape[0]
    n = N if mode == M_PF_D else D * N
    md, mp = mat_desc(mat, grad_term, nh2d_extension)
    if type(mat).__name__ == "Ogden" and ((D != 3 and kind != "CAS") or float(mat.alpha) != int(mat.alpha) or not 1 <= mat.alpha <= 8):
        # the C++ restatement differentiates Σλ^α = tr(U^α) through a matrix square root: 3-D, integer α only
        raise ValueError("oracle (C++): Ogden is restated for 3-D elements and integer alpha in 1..8; use oracle/materials.py otherwise")
    val = np.empty(ne)
    g = np.empty((ne, n))
    H = np.empty((ne, n, n))
    if hist is not None:
        assert hist.dtype == np.float64 and hist.flags.c_contiguous and hist.shape == (ne, P, 2)
    rc = L.oracle_elem_duals(C.c_int({"CE": 0, "CP": 1, "CAS": 2}[kind]), C.c_int(D), C.c_int(P), C.c_int(N), _ptr(md), _ptr(mp),
                             C.c_int64(ne), _ptr(nodes), _ptr(gradN), _ptr(wgt), _ptr(Nval), C.c_int(mode), _ptr(u),
                             C.c_int(dpn), _ptr(d), _ptr(hist), C.c_int(nthreads), _ptr(val), _ptr(g), _ptr(H))
    if rc != 0:
        raise ValueError(f"oracle_elem_duals: unsupported (D={D}, N={N})")
    return val, g, H


def assemble(nodes, dpn, val, g, H, ndof):
    """Serial COO->CSC assembly: returns ϕ, r, (colptr, rowval, nzval) (1-based Int64)."""
    L = lib()
    nodes = np.ascontiguousarray(nodes, dtype=np.int64)
    ne, N = nodes.shape
    n = dpn * N
    assert g.shape == (ne, n) and H.shape == (ne, n, n)
    phi = C.c_double(0.0)
    r = np.empty(ndof)
    h = L.oracle_assemble(C.c_int64(ne), C.c_int(N), C.c_int(dpn), C.c_int(n), _ptr(nodes), _ptr(val), _ptr(g), _ptr(H),
                          C.c_int64(ndof), C.byref(phi), _ptr(r))
    nnz = L.oracle_result_nnz(h)
    colptr = np.empty(ndof + 1, dtype=np.int64)
    rowval = np.empty(nnz, dtype=np.int64)
    nzval = np.empty(nnz)
    L.oracle_result_copy(h, _ptr(colptr), _ptr(rowval), _ptr(nzval))
    L.oracle_result_free(h)
    return phi.value, r, (colptr, rowval, nzval)


def make_phi_r_Kt(batches, u, mode=M_ELASTIC, d=None, hist=None, nthreads=0, timing=None, **kw):
    """`makeϕrKt` over a list of batches (objects with kind, D, P, N, mat, nodes, gradN, wgt, Nval;
    concatenated in list order = `elems` order).  mode selects elastic / PF u-phase / PF d-phase."""
    u = np.asfortranarray(u, dtype=np.float64)
    t0 = time.perf_counter()
    vals, gs, Hs, nds = [], [], [], []
    for i, b in enumerate(batches):
        h = None if hist is None else hist[i]
        Nv = b.Nval
        if b.kind == "CAS":   # axisymmetric extension: the C++ side takes the hoop row N[gp][a] / X0[gp]
            Nv = np.ascontiguousarray(np.asarray(b.Nval) / np.asarray(b.X0)[:, :, None])
        v, g, H = elem_duals(b.kind, b.D, b.P, b.N, b.mat, b.nodes, b.gradN, b.wgt, Nv, mode, u, d, h, nthreads, **kw)
        if mode != M_PF_D and u.shape[0] > b.D:
            # the reference seeds every row of u[:, nodes] (elasticelements.jl:217) but getF reads rows 1:D only
            # (toolkit:13): the extra dofs carry exact-zero derivatives.  Expand (D*N) -> (dpn*N).
            dpn_, D_, N_ = u.shape[0], b.D, b.N
            keep = (np.arange(N_)[:, None] * dpn_ + np.arange(D_)[None, :]).reshape(-1)
            g2 = np.zeros((g.shape[0], dpn_ * N_))
            H2 = np.zeros((H.shape[0], dpn_ * N_, dpn_ * N_))
            g2[:, keep] = g
            H2[:, keep[:, None], keep[None, :]] = H
            g, H = g2, H2
        vals.append(v), gs.append(g), Hs.append(H), nds.append(np.asarray(b.nodes, dtype=np.int64))
    t1 = time.perf_counter()
    if mode == M_PF_D:
        dpn, ndof = 1, int(np.asarray(d).size)
    else:
        dpn, ndof = u.shape[0], u.size
    out = assemble(np.concatenate(nds), dpn, np.concatenate(vals), np.concatenate(gs), np.concatenate(Hs), ndof)
    t2 = time.perf_counter()
    if timing is not None:
        timing["elements_s"] = t1 - t0
        timing["assembly_s"] = t2 - t1
    return out
