"""ctypes binding of `libad4sm_b200.so` (include/ad4sm_b200.h).  The library is the product: if it is
missing or no CUDA device is present the compute calls raise -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AD4SM_B200_LIB") or os.path.join(_HERE, "lib", "libad4sm_b200.so")

OK, EINVAL, ENODEV, ENOMEM, ESTATE, ENCCL = 0, -1, -2, -3, -4, -5
ELEM_CE, ELEM_CP, ELEM_CAS = 0, 1, 2
MAT_HOOKE, MAT_HOOKE2D, MAT_NEOHOOKE, MAT_MOONEYRIVLIN, MAT_OGDEN, MAT_PHASEFIELD = 1, 2, 3, 4, 5, 6
F_SMALL, F_PLANE_STRESS, F_DHN, F_PF_GRAD_INTENDED, F_NH2D_EXT = 1, 2, 4, 8, 16
FIELD_U, FIELD_D = 0, 1
MODE_ELASTIC, MODE_PF_U, MODE_PF_D, MODE_PHI_R = 0, 1, 2, 3
PROF_ELEMENT, PROF_ASSEMBLE, PROF_REDUCE, PROF_GATHER, PROF_NCLASS = 0, 1, 2, 3, 4


SHAPE_TRIA3, SHAPE_QUAD4, SHAPE_TET4, SHAPE_HEX8, SHAPE_WDG6 = 1, 2, 3, 4, 5


class Mesh(C.Structure):
    _fields_ = [("shape", C.c_int32), ("n_gp_1d", C.c_int32), ("gp_xi", C.c_double * 3), ("gp_w", C.c_double * 3),
                ("coord_dim", C.c_int32), ("reserved", C.c_int32), ("n_nodes", C.c_int64), ("coords", C.c_void_p)]


class ConstEqSet(C.Structure):
    _fields_ = [("n_eqs", C.c_int64), ("dof_ptr", C.c_void_p), ("dofs", C.c_void_p), ("c", C.c_void_p), ("c0", C.c_void_p),
                ("q_ptr", C.c_void_p), ("Q", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [
        ("elem_kind", C.c_int32), ("dim", C.c_int32), ("n_gauss", C.c_int32), ("n_elem_nodes", C.c_int32),
        ("mat_kind", C.c_int32), ("mat_base", C.c_int32), ("mat_flags", C.c_uint32), ("batch_flags", C.c_int32),
        ("mat_params", C.c_double * 8),
        ("n_elems", C.c_int64),
        ("nodes", C.c_void_p), ("gradN", C.c_void_p), ("wgt", C.c_void_p), ("Nval", C.c_void_p),
        ("orig_index", C.c_void_p), ("mesh", C.POINTER(Mesh)),
    ]


BATCH_HALO = 1
OPT_SORTED, OPT_PEER_PUSH = 3, 4
INFO_SORTED = 3
INFO_SPLIT_PAIRS = 4
INFO_DIST_TRANSPORT = 5


class StagingView(C.Structure):
    _fields_ = [("n_slots", C.c_int64), ("ke_width", C.c_int64), ("re_width", C.c_int64),
                ("Ke_dev", C.c_void_p), ("re_dev", C.c_void_p), ("phie_dev", C.c_void_p)]


PEER_HANDLE_BYTES = 128


class PeerLink(C.Structure):
    _fields_ = [("first_slot", C.c_int64), ("count", C.c_int64), ("remote_first_slot", C.c_int64), ("peer_handles", C.c_void_p),
                ("remote_first_slot_b", C.c_int64)]


class DeviceView(C.Structure):
    _fields_ = [
        ("n_dof", C.c_int64), ("nnz", C.c_int64),
        ("phi_dev", C.c_void_p), ("r_dev", C.c_void_p), ("nzval_dev", C.c_void_p),
        ("colptr_dev", C.c_void_p), ("rowval_dev", C.c_void_p), ("n_zero_dev", C.c_void_p),
    ]


NCCL_ID_BYTES = 128


class DistPeer(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_send", C.c_int64), ("send_slots", C.c_void_p), ("recv_first_slot", C.c_int64),
                ("n_recv", C.c_int64)]


class NewtonView(C.Structure):
    _fields_ = [("nfree", C.c_int64), ("ncnst", C.c_int64), ("nnz_ff", C.c_int64), ("res_dev", C.c_void_p),
                ("nzval_ff_dev", C.c_void_p), ("norms_dev", C.c_void_p), ("colptr_ff_dev", C.c_void_p), ("rowval_ff_dev", C.c_void_p)]


class AD4SMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libad4sm_b200 error {code}: {msg}")
        self.code = code


_lib = None

# name -> (restype, argtypes); also the list of symbols include/ad4sm_b200.h declares
SIGNATURES = {
    "ad4sm_abi_version": (C.c_int, []),
    "ad4sm_last_error": (C.c_char_p, []),
    "ad4sm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ad4sm_plan_create": (C.c_int, [C.POINTER(Batch), C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ad4sm_plan_destroy": (None, [C.c_void_p]),
    "ad4sm_consteq_eval": (C.c_int, [C.POINTER(ConstEqSet), C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "ad4sm_construct": (C.c_int, [C.POINTER(Mesh), C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_construct_device": (C.c_int, [C.POINTER(Mesh), C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "ad4sm_plan_pattern_size": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ad4sm_plan_pattern": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ad4sm_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int64)]),
    "ad4sm_eval_phi_r": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "ad4sm_eval_pf_u": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int64)]),
    "ad4sm_eval_pf_d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.c_void_p,
                                  C.POINTER(C.c_int64)]),
    "ad4sm_fetch_csc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ad4sm_eval_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_device_view_get": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(DeviceView)]),
    "ad4sm_eval_elements_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_assemble_duals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_assemble_duals_phi_r": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ad4sm_eval_elements_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_fetch_phi_r_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ad4sm_eval_assemble_device": (C.c_int, [C.c_void_p, C.c_int32]),
    "ad4sm_eval_assemble_split_device": (C.c_int, [C.c_void_p, C.c_int32]),
    "ad4sm_eval_halo_prepare_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "ad4sm_staging_view_get": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(StagingView)]),
    "ad4sm_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ad4sm_peer_attach": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(PeerLink)]),
    "ad4sm_export_ranges": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ad4sm_sync": (C.c_int, [C.c_void_p]),
    "ad4sm_set_option": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    "ad4sm_get_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "ad4sm_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "ad4sm_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "ad4sm_newton_setup": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ad4sm_newton_pattern": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ad4sm_newton_reduce": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_newton_reduce_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ad4sm_newton_view_get": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(NewtonView)]),
    "ad4sm_dist_unique_id": (C.c_int, [C.c_void_p]),
    "ad4sm_dist_init": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "ad4sm_dist_destroy": (None, [C.c_void_p]),
    "ad4sm_dist_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(DistPeer)]),
    "ad4sm_dist_eval_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_dist_eval": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ad4sm_post": (C.c_int, [C.c_void_p] * 7),
    "ad4sm_post_device": (C.c_int, [C.c_void_p] * 7),
    "ad4sm_measure_fp64_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def lib():
    """Load the shared library (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AD4SMError(ENODEV, f"{LIB_PATH} not built: run `make -C ad4sm.jl_b200/csrc` "
                                     "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise AD4SMError(rc, lib().ad4sm_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().ad4sm_device_count(C.byref(n))
    return n.value if rc == OK else 0
