"""TEST INFRASTRUCTURE: builds tests/cxx/host_harness.cpp (the product's csrc/*.cuh compiled for the
host) and exposes it through ctypes, so that the kernels' arithmetic can be compared with the oracle
on a box without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "cxx", "host_harness.cpp")
_INC = os.path.join(_ROOT, "ad4sm.jl_b200", "csrc")
_SO = os.path.join(_HERE, "cxx", "_build", "libhost_harness.so")
_lib = None

MF_HOOKE, MF_NEO, MF_MOONEY, MF_PFNEO, MF_OGDEN = 0, 1, 2, 3, 4


def lib():
    global _lib
    if _lib is None:
        deps = [_SRC] + [os.path.join(_INC, f) for f in os.listdir(_INC) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", _INC, "-o", _SO, _SRC])
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def family(desc):
    kind, base = int(desc[0]), int(desc[1])
    k = base if kind == 6 else kind
    if kind == 6 and k == 3 and not (int(desc[2]) & 4):
        return MF_PFNEO
    return {1: MF_HOOKE, 2: MF_HOOKE, 3: MF_NEO, 4: MF_MOONEY, 5: MF_OGDEN}[k]


def tangent(D, desc, params, F, pf=None):
    """ψ, P[DD], A[packed] at F (column-major flattening F[k*D+j] = F_jk)."""
    F = np.asarray(F, dtype=float).reshape(D, D)
    Fl = np.ascontiguousarray(F.reshape(-1, order="F"))
    DD = D * D
    psi = C.c_double()
    P = np.empty(DD)
    A = np.empty(DD * (DD + 1) // 2)
    desc = np.ascontiguousarray(desc, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    d, gd2 = (pf if pf is not None else (0.0, 0.0))
    rc = lib().h_tangent(D, family(desc), _p(desc), _p(params), _p(Fl), int(pf is not None), C.c_double(d), C.c_double(gd2),
                         C.byref(psi), _p(P), _p(A))
    assert rc == 0
    return psi.value, P, A


def pair_slot(N, a, b):
    i, t = C.c_int(), C.c_int()
    lib().h_pair_slot(N, a, b, C.byref(i), C.byref(t))
    return i.value, t.value


def expand_blocks(Ke, N, DB):
    """block storage [NPB][DB*DB] -> dense (N*DB) x (N*DB), dof = a*DB + j"""
    K = np.empty((N * DB, N * DB))
    B = Ke.reshape(-1, DB, DB)
    for a in range(N):
        for b in range(N):
            idx, tr = pair_slot(N, a, b)
            K[a * DB:(a + 1) * DB, b * DB:(b + 1) * DB] = B[idx].T if tr else B[idx]
    return K


def element_u(D, N, desc, params, gN, wgt, ue, Nv=None, de=None):
    """one element: gN [D][P][N], wgt [P], ue [N][D] (row a = node) -> ϕ, r[N*D], K dense"""
    P = len(wgt)
    gN = np.ascontiguousarray(gN, dtype=float)
    wgt = np.ascontiguousarray(wgt, dtype=float)
    ue = np.ascontiguousarray(ue, dtype=float)
    Nv = None if Nv is None else np.ascontiguousarray(Nv, dtype=float)
    de = None if de is None else np.ascontiguousarray(de, dtype=float)
    desc = np.ascontiguousarray(desc, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    phi = C.c_double()
    re = np.empty(N * D)
    Ke = np.empty(N * (N + 1) // 2 * D * D)
    rc = lib().h_element_u(D, N, family(desc), int(de is not None), _p(desc), _p(params), P, _p(gN), _p(wgt), _p(Nv),
                           _p(ue), _p(de), C.byref(phi), _p(re), _p(Ke))
    assert rc == 0, rc
    return phi.value, re, expand_blocks(Ke, N, D)


def element_d(D, N, desc, params, gN, wgt, Nv, ue, de, hist=None):
    P = len(wgt)
    gN = np.ascontiguousarray(gN, dtype=float)
    wgt = np.ascontiguousarray(wgt, dtype=float)
    ue = np.ascontiguousarray(ue, dtype=float)
    Nv = np.ascontiguousarray(Nv, dtype=float)
    de = np.ascontiguousarray(de, dtype=float)
    desc = np.ascontiguousarray(desc, dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float64)
    phi = C.c_double()
    re = np.empty(N)
    Ke = np.empty(N * (N + 1) // 2)
    rc = lib().h_element_d(D, N, _p(desc), _p(params), P, _p(gN), _p(wgt), _p(Nv), _p(ue), _p(de), _p(hist),
                           C.byref(phi), _p(re), _p(Ke))
    assert rc == 0, rc
    return phi.value, re, expand_blocks(Ke, N, 1)
