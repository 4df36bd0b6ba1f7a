"""BASELINE.json's configurations at FULL size on the GPU.  Where the oracle finishes in seconds (C1, C4) the result is
compared with it directly; C2, C3 and C5 are compared with the oracle on the LARGEST block of the same mesh /
material / u generator the oracle finishes in seconds (56^3 Hex8 NeoHooke, 209 088 Tet4 MooneyRivlin, a 200x200x2 slab
of Hex8 Ogden + its NeoHooke control), per entry (tests/_cases.assert_parity), and -- opt-in, AD4SM_FULL_ORACLE=1, it
needs ~60 GB of host memory and minutes of serial assembly -- C2 at its full 100^3; at the full multi-million-element
sizes parity is checked through size-independent
properties of the reference's output contract: run-to-run bitwise reproducibility, sorted 1-based CSC, the translation
null space of Kt (ϕ depends on ∇u only), and consistency of ϕ -> r -> Kt by central differences along a random
direction.  Sparse products run on the device (torch CSR view of the plan's arrays; Kt is symmetric, CSC ≡ CSR)."""
import os

import numpy as np
import pytest

from ad4sm_b200 import Elements as El, Materials as Ma, _capi as capi, mesh
from ad4sm_b200.dist import _wrap
from oracle import cxx_binding as X

import _cases as K

pytestmark = pytest.mark.gpu


def _device_csr(plan, torch):
    v = plan.device_view()
    n, nnz = v.n_dof, v.nnz
    crow = torch.as_tensor(_I64(v.colptr_dev, (n + 1,)), device="cuda") - 1
    col = torch.as_tensor(_I64(v.rowval_dev, (nnz,)), device="cuda") - 1
    val = _wrap(torch, v.nzval_dev, (nnz,))
    return torch.sparse_csr_tensor(crow, col, val, size=(n, n)), crow, col, val


class _I64:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _properties(batch, n_nodes, D, amp_h, seed):
    import torch
    plan = El.Plan(batch, n_nodes, D)
    rng = np.random.default_rng(seed)
    u = 0.02 * amp_h * rng.uniform(-1, 1, (n_nodes, D))
    ut = torch.from_numpy(u).cuda()
    v = plan.device_view()

    def ev(x):
        plan.eval_device(capi.MODE_ELASTIC, x.data_ptr())
        plan.sync()
        return float(_wrap(torch, v.phi_dev, (1,))[0]), _wrap(torch, v.r_dev, (v.n_dof,)).clone()

    phi, r = ev(ut)
    A, crow, col, val = _device_csr(plan, torch)
    nz1 = val.clone()
    phi2, r2 = ev(ut)
    assert phi2 == phi and torch.equal(r, r2) and torch.equal(nz1, val), "two evaluations must be bitwise identical"
    # structure: 1-based ascending rows inside each column, no exact zeros on a (jitter-free but deformed) mesh is not
    # guaranteed, so only ordering is asserted
    d = col[1:] - col[:-1]
    starts = torch.zeros_like(d, dtype=torch.bool)
    starts[(crow[1:-1] - 1).clamp(min=0)] = True
    assert bool(((d > 0) | starts).all()), "row indices must ascend inside every column"
    scale = float(val.abs().max())
    # translations are in the null space of Kt
    for j in range(D):
        t = torch.zeros(n_nodes, D, dtype=torch.float64, device="cuda")
        t[:, j] = 1.0
        assert float((A @ t.reshape(-1)).abs().max()) <= 1e-9 * scale
    # ϕ -> r -> Kt along a random direction (central differences)
    w = torch.from_numpy(rng.uniform(-1, 1, (n_nodes, D))).cuda()
    h = 1e-6 * amp_h
    pp, rp = ev(ut + h * w)
    pm, rm = ev(ut - h * w)
    wv = w.reshape(-1)
    dphi = (pp - pm) / (2 * h)
    assert abs(dphi - float(r @ wv)) <= 1e-5 * max(abs(dphi), float(r.abs().max()))
    Kw = A @ wv
    fd = (rp - rm) / (2 * h)
    assert float((fd - Kw).abs().max()) <= 1e-5 * float(Kw.abs().max())
    return plan


def test_c2_hex8_neohooke_1M_properties():
    n = 100
    coords, conn = mesh.hex_block(n, n, n)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    plan = _properties(b, len(coords), 3, 1.0 / n, 1)
    assert plan.pattern_size() == (3090903, 245438109)


def test_c3_tet4_mooneyrivlin_2M_properties():
    coords, conn = mesh.tet_kuhn(70, 70, 68)
    b = El.Tet04_batch(conn, coords, Ma.MooneyRivlin(800.0, 200.0, 5000.0))
    assert b.ne == 1999200
    plan = _properties(b, len(coords), 3, 1.0 / 70, 2)
    assert plan.pattern_size() == (1043487, 45896085)


def test_c1_quad4_hooke2d_200x200_vs_oracle():
    coords, conn = mesh.quad_grid(200, 200, jitter=0.1, seed=5)
    b = El.Quad_batch(conn, coords, Ma.Hooke2D(210000.0, 0.3, plane_stress=False))
    u = mesh.random_u(len(coords), 2, 1 / 200, seed=6)
    got = El.Plan(b, u.shape[1], 2).makeϕrKt(u)
    assert got[2].nnz == 1444804
    K.assert_parity(got, K.oracle_eval(b, u))
    # linear material: r = Kt u and ϕ = u·r/2
    Ks = got[2].to_scipy()
    uf = u.reshape(-1, order="F")
    assert np.abs(Ks @ uf - got[1]).max() <= 1e-10 * np.abs(got[1]).max()
    assert abs(0.5 * uf @ got[1] - got[0]) <= 1e-10 * abs(got[0])


def test_c4_quadp_phasefield_500x500_vs_oracle():
    """staggered u / d assembly with history on the 500x500 RVE (PhaseField over Hooke2D: the variant the reference can
    execute)."""
    coords, conn = mesh.quad_grid(500, 500, jitter=0.1, seed=7)
    pf = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
    b = El.QuadP_batch(conn, coords, pf)
    u = mesh.random_u(len(coords), 2, 1 / 500, a=0.3, seed=8)
    d = K.random_d(u.shape[1], seed=9)
    plan = El.Plan(b, u.shape[1], 2)
    gu = plan.makeϕrKt(u, d)
    assert gu[2].nnz == 9012004
    K.assert_parity(gu, K.oracle_eval(b, u, mode=X.M_PF_U, d=d))
    h_gpu = np.zeros((b.ne, b.P, 2))
    h_ref = h_gpu.copy()
    gd = plan.makeϕrKt_d(u, d, h_gpu)
    assert gd[2].nnz == 2253001
    K.assert_parity(gd, K.oracle_eval(b, u, mode=X.M_PF_D, d=d, hist=[h_ref]))
    assert np.abs(h_gpu - h_ref).max() <= 1e-12 * max(1.0, np.abs(h_ref).max())
    assert h_gpu.max() > 0.0


# ---- the named configurations against the oracle, at the largest size the oracle finishes in seconds -----------------
def _vs_oracle(b, u, label, tol=K.TOL):
    got = El.Plan(b, u.shape[1], u.shape[0]).makeϕrKt(u)
    st = {}
    K.assert_parity(got, K.oracle_eval(b, u), tol=tol, stats=st)
    print(f"{label}: ne={b.ne} nnz={got[2].nnz} max per-entry rel diff {st['max_rel']:.2e}, "
          f"{st['n_floor']} entries on the cancellation floor")
    return got


def test_c2_hex8_neohooke_56cube_vs_oracle():
    """C2's mesh generator, material and u at 56^3 (175 616 elements, 43.4 M stored values), jittered, per entry."""
    b, u = K.hex_case((56, 56, 56), seed=21)
    got = _vs_oracle(b, u, "C2@56^3")
    assert got[2].nnz == (3 * 56 + 1) ** 3 * 9


def test_c3_tet4_mooneyrivlin_209k_vs_oracle():
    b, u = K.tet_case((33, 33, 32), seed=22)
    assert b.ne == 33 * 33 * 32 * 6
    _vs_oracle(b, u, "C3@209k")


@pytest.mark.parametrize("mat", ["ogden", "neohooke"])
def test_c5_hex8_slab_vs_oracle(mat):
    """C5's 200x200 planes (one element layer of the 200^3 mesh): Ogden(3, 1000, 5000) against the oracle's independent
    derivation of the reference's energy formula (dual arithmetic through tr(U^α) with a matrix square root, no
    eigen-decomposition: oracle/cxx phi_ogden3; the reference itself cannot differentiate Ogden) -- and the
    NeoHooke control the reference can execute."""
    m = Ma.Ogden(3.0, 1000.0, 5000.0) if mat == "ogden" else Ma.NeoHooke(1000.0, 5000.0)
    b, u = K.hex_case((200, 200, 1), mat=m, seed=23)
    # NeoHooke: the north-star 1e-12.  Ogden has NO reference values (the reference cannot differentiate it); product and
    # oracle are two different formulations of the derivative (spectral divided differences vs dual arithmetic through a
    # matrix square root), not two orderings of one arithmetic, so they are held to 1e-11 -- measured: every entry within
    # 1e-13 of its natural scale, 0.4 % of the entries beyond a pure relative 1e-12.
    _vs_oracle(b, u, f"C5 slab ({mat})", tol=1e-11 if mat == "ogden" else K.TOL)


@pytest.mark.skipif(os.environ.get("AD4SM_FULL_ORACLE") != "1", reason="opt-in: the serial oracle assembly of 576 M "
                    "triplets needs ~60 GB of host memory and minutes (set AD4SM_FULL_ORACLE=1)")
def test_c2_hex8_neohooke_100cube_vs_oracle():
    b, u = K.hex_case((100, 100, 100), seed=24)
    got = _vs_oracle(b, u, "C2@100^3")
    assert got[2].nnz == 245438109
