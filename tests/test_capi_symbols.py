"""CPU-only checks of the drop-in boundary: libad4sm_b200.so loads, exports every entry point that
include/ad4sm_b200.h declares (and the ctypes mirror binds exactly those), and -- there being no CPU fallback --
refuses to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ad4sm_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ad4sm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ad4sm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 20
    L = C.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ad4sm_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == names, "ctypes mirror and header disagree"


def test_header_constants_match_python_mirror():
    txt = open(os.path.join(ROOT, "include", "ad4sm_b200.h")).read()
    for name, val in re.findall(r"#define\s+AD4SM_([A-Z0-9_]+)\s+\(?(-?\d+)u?\)?", txt):
        py = {"ABI_VERSION": None, "OK": "OK"}.get(name, name)
        if py and hasattr(capi, py):
            assert getattr(capi, py) == int(val), name
    assert capi.lib().ad4sm_abi_version() == 2
    assert C.sizeof(capi.Batch) == 8 * 4 + 8 * 8 + 8 + 6 * 8
    assert C.sizeof(capi.Mesh) == 2 * 4 + 6 * 8 + 2 * 4 + 8 + 8


@pytest.mark.skipif(capi.device_count() > 0, reason="needs a box WITHOUT a GPU")
def test_no_cpu_fallback():
    from ad4sm_b200 import Elements as El, Materials as Ma, mesh
    coords, conn = mesh.hex_block(2, 2, 2)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    with pytest.raises(capi.AD4SMError) as ei:
        El.makeϕrKt(b, np.zeros((3, len(coords))))
    assert ei.value.code == capi.ENODEV


def test_eval_elements_host_rejects_wrong_layout():
    """A C-ordered D x nNodes array is contiguous but would be read as [node][component]: refuse it (no device needed,
    the check runs before the library call)."""
    from ad4sm_b200 import Elements as El

    class _P(El.Plan):
        def __init__(self):
            self.dpn, self.n_nodes = 3, 10
    p = _P()
    for bad in (np.zeros((3, 10)), np.zeros((3, 9), order="F"), np.zeros((3, 10), dtype=np.float32, order="F")):
        with pytest.raises(ValueError):
            p.eval_elements_host(0, bad)
    with pytest.raises(ValueError):
        p.eval_elements_host(1, np.zeros((3, 10), order="F"), d=np.zeros(9))


def test_nccl_is_bound_at_run_time():
    """ad4sm_dist_unique_id dlopens libnccl.so.2 (no link-time dependency) and returns a 128-byte id; without a CUDA
    device ad4sm_dist_init must fail with a status code, not crash."""
    L = capi.lib()
    buf = (C.c_ubyte * capi.NCCL_ID_BYTES)()
    assert L.ad4sm_dist_unique_id(buf) == capi.OK, L.ad4sm_last_error()
    assert any(bytes(buf))
    if capi.device_count() == 0:
        h = C.c_void_p()
        rc = L.ad4sm_dist_init(1, 0, buf, -1, C.byref(h))
        assert rc in (capi.ENODEV, capi.ENCCL) and not h.value
    import subprocess
    out = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl" not in out, "libnccl must not be a NEEDED dependency of libad4sm_b200.so"
