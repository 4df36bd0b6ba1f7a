"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bar (north_star): CSC pattern bit-exact, ϕ / r / nzval to relative 1e-12."""
import numpy as np
import pytest

from ad4sm_b200 import Elements as El, Materials as Ma, _capi as capi
from oracle import cxx_binding as X

import _cases as K

pytestmark = pytest.mark.gpu


def _run(b, u, **opts):
    plan = El.Plan(b, u.shape[1], u.shape[0], **opts)
    return plan, plan.makeϕrKt(u)


ELASTIC = [
    ("hex8-neohooke", lambda: K.hex_case()),
    ("hex8-neohooke-incompressible", lambda: K.hex_case(mat=Ma.NeoHooke(1000.0, -1.0))),
    ("hex8-mooney", lambda: K.hex_case(mat=Ma.MooneyRivlin(800.0, 200.0, 5000.0))),
    ("hex8-hooke-gl", lambda: K.hex_case(mat=Ma.Hooke(210000.0, 0.3))),
    ("hex8-hooke-small", lambda: K.hex_case(mat=Ma.Hooke(210000.0, 0.3, small=True))),
    ("hex8r-neohooke", lambda: K.hex_case(ctor=lambda c, x, m: El.Hex08_batch(c, x, m, GP=El.GP1))),
    ("tet4-mooney", lambda: K.tet_case()),
    ("tet4-neohooke", lambda: K.tet_case(mat=K.nh())),
    ("tet4-hooke", lambda: K.tet_case(mat=Ma.Hooke(210000.0, 0.3))),
    ("quad4-hooke2d-pe", lambda: K.quad_case()),
    ("quad4-hooke2d-ps", lambda: K.quad_case(mat=Ma.Hooke2D(210000.0, 0.3))),
    ("quad4-hooke2d-gl", lambda: K.quad_case(mat=Ma.Hooke2D(210000.0, 0.3, small=False))),
    ("quad4-neohooke", lambda: K.quad_case(mat=K.nh())),
    ("quad4-neohooke-ps", lambda: K.quad_case(mat=Ma.NeoHooke(1000.0, -1.0))),
    ("quad4-mooney", lambda: K.quad_case(mat=Ma.MooneyRivlin(800.0, 200.0, 5000.0))),
    ("quadr-hooke2d", lambda: K.quad_case(ctor=lambda c, x, m: El.Quad_batch(c, x, m, GP=El.GP1))),
    ("tria3-hooke2d", lambda: K.tri_case()),
    ("tria3-neohooke", lambda: K.tri_case(mat=K.nh())),
]


@pytest.mark.parametrize("name,make", ELASTIC, ids=[c[0] for c in ELASTIC])
def test_elastic_parity(name, make):
    b, u = make()
    plan, got = _run(b, u)
    K.assert_parity(got, K.oracle_eval(b, u))
    Kd = K.csc_dense(got[2])
    assert np.array_equal(Kd, Kd.T), "Kt must be bitwise symmetric"
    again = plan.makeϕrKt(u)
    assert again[0] == got[0] and np.array_equal(again[1], got[1]) and np.array_equal(again[2].nzval, got[2].nzval), \
        "two evaluations must be bitwise identical (deterministic assembly)"
    phi2, r2 = plan.makeϕr(u)
    assert phi2 == got[0] and np.array_equal(r2, got[1])


def test_structural_pattern_matches_csc():
    b, u = K.hex_case((3, 3, 2))
    plan, got = _run(b, u)
    cp, rv = plan.pattern()
    assert np.array_equal(cp, got[2].colptr) and np.array_equal(rv, got[2].rowval)
    n, nnz = plan.pattern_size()
    assert n == u.size and nnz == len(rv)


def test_dropzeros_extra_u_rows():
    """A 3-row u with 2-D elements: the reference seeds all 3 rows but getF reads rows 1:2 (toolkit:13), so
    every entry touching component 3 is an exact zero that dropzeros removes (toolkit:202)."""
    b, u = K.quad_case((5, 4), dpn=3)
    plan, got = _run(b, u)
    ref = K.oracle_eval(b, u)
    K.assert_parity(got, ref)
    n, nnz_struct = plan.pattern_size()
    assert got[2].nnz < nnz_struct
    assert np.all(got[1].reshape(-1, 3)[:, 2] == 0.0)


def test_mixed_batches_follow_elems_order():
    """A Vector of element objects with two materials interleaved: duplicates must be summed in `elems` order."""
    from ad4sm_b200 import mesh
    coords, conn = mesh.hex_block(3, 3, 2, jitter=0.1, seed=5)
    m1, m2 = K.nh(), Ma.MooneyRivlin(800.0, 200.0, 5000.0)
    elems = [El.Hex08(conn[e], coords[conn[e] - 1], mat=(m1 if e % 3 else m2)) for e in range(len(conn))]
    u = mesh.random_u(len(coords), 3, 1 / 3, seed=9)
    got = El.makeϕrKt(elems, u)
    # oracle: one batch per element keeps the exact `elems` order
    ob = [El.ElemBatch("CE", 3, e.nodes[None], e.gradN[None], e.wgt[None], e.mat) for e in elems]
    K.assert_parity(got, K.oracle_eval(ob, u))


PF2 = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
PF2S = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(210000.0, 0.3), 2, AT="DHn")
PF3 = Ma.PhaseField(0.01, 100.0, Ma.Hooke(210000.0, 0.3, small=True), 2)
PF3N = Ma.PhaseField(0.01, 100.0, K.nh(), 2)
PFCASES = [
    ("quadp-hooke2d-atn", lambda: K.quad_case(mat=PF2, ctor=El.QuadP_batch, amp=0.3), {}),
    ("quadp-hooke2d-ps-dhn", lambda: K.quad_case(mat=PF2S, ctor=El.QuadP_batch, amp=0.3), {}),
    ("quadp-neohooke-ext", lambda: K.quad_case(mat=PF3N, ctor=El.QuadP_batch, amp=0.3), {"nh2d_extension": True}),
    ("triap-hooke2d", lambda: K.tri_case(mat=PF2, ctor=El.TriaP_batch, amp=0.3), {}),
    ("hex8p-hooke", lambda: K.hex_case((4, 3, 3), mat=PF3, ctor=El.Hex08P_batch, amp=0.3), {}),
    ("hex8p-neohooke", lambda: K.hex_case((4, 3, 3), mat=PF3N, ctor=El.Hex08P_batch, amp=0.3), {}),
    ("tet4p-neohooke", lambda: K.tet_case(mat=PF3N, ctor=El.Tet04P_batch, amp=0.1), {}),  # amp 0.3 gives det F < 0: log(J) = NaN on both sides
]


@pytest.mark.parametrize("name,make,opts", PFCASES, ids=[c[0] for c in PFCASES])
def test_phasefield_u_phase(name, make, opts):
    b, u = make()
    d = K.random_d(u.shape[1])
    plan = El.Plan(b, u.shape[1], u.shape[0], **opts)
    got = plan.makeϕrKt(u, d)
    K.assert_parity(got, K.oracle_eval(b, u, mode=X.M_PF_U, d=d, **opts))


@pytest.mark.parametrize("grad_term", ["reference", "intended"])
@pytest.mark.parametrize("with_hist", [False, True])
@pytest.mark.parametrize("name,make,opts", PFCASES, ids=[c[0] for c in PFCASES])
def test_phasefield_d_phase(name, make, opts, with_hist, grad_term):
    b, u = make()
    d = K.random_d(u.shape[1])
    plan = El.Plan(b, u.shape[1], u.shape[0], pf_grad_term=grad_term, **opts)
    rng = np.random.default_rng(3)
    h_gpu = h_ref = None
    if with_hist:
        h0 = rng.uniform(0.0, 2.0, (b.ne, b.P, 2)) * (rng.uniform(size=(b.ne, 1, 1)) > 0.5)
        h_gpu, h_ref = h0.copy(), h0.copy()
    got = plan.makeϕrKt_d(u, d, h_gpu)
    ref = K.oracle_eval(b, u, mode=X.M_PF_D, d=d, hist=None if h_ref is None else [h_ref], grad_term=grad_term, **opts)
    K.assert_parity(got, ref)
    if with_hist:
        assert np.abs(h_gpu - h_ref).max() <= 1e-12 * max(1.0, np.abs(h_ref).max())
        assert not np.array_equal(h_gpu, h0), "history must be updated in place"


def test_error_behaviour():
    b, u = K.hex_case((2, 2, 2))
    with pytest.raises(AssertionError):
        El.makeϕrKt([], u)  # @assert nElems > 0, elasticelements.jl:210
    bad = El.ElemBatch("CE", 3, b.nodes + 1000, b.gradN, b.wgt, b.mat)
    with pytest.raises(capi.AD4SMError):
        El.Plan(bad, u.shape[1], 3)
    with pytest.raises(capi.AD4SMError):  # PhaseField material without d
        p = El.Plan(El.Hex08P_batch(b.nodes, np.zeros((u.shape[1], 3)) + np.random.default_rng(0).uniform(size=(u.shape[1], 3)),
                                    PF3), u.shape[1], 3)
        p.makeϕrKt(u)
    with pytest.raises(capi.AD4SMError):  # 2-D PhaseField{NeoHooke}: BoundsError upstream unless the extension is requested
        q, uq = K.quad_case(mat=PF3N, ctor=El.QuadP_batch)
        El.Plan(q, uq.shape[1], 2)


def test_device_resident_path_matches_host_path():
    import torch
    b, u = K.hex_case()
    plan, got = _run(b, u)
    ut = torch.from_numpy(np.ascontiguousarray(u.T)).cuda()  # [node][comp] == column-major (dpn, n_nodes)
    plan.eval_device(capi.MODE_ELASTIC, ut.data_ptr())
    plan.sync()
    Kt = plan.fetch_csc()
    assert np.array_equal(Kt.nzval, got[2].nzval) and np.array_equal(Kt.rowval, got[2].rowval)


def test_fp64_peak_microbenchmark_runs():
    import ctypes as C
    tf, mhz = C.c_double(), C.c_double()
    capi.check(capi.lib().ad4sm_measure_fp64_peak(0, C.byref(tf), C.byref(mhz)))
    assert 5.0 < tf.value < 80.0, tf.value


# ---- golden vectors (tests/golden/*.npz, literal Python restatement of the reference) through the C ABI ----------
import glob as _glob
import os as _os

_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLDEN, ids=[_os.path.basename(p)[:-4] for p in _GOLDEN])
def test_cuda_path_reproduces_golden_vectors(path):
    import _golden
    case = _golden.load(path)
    _golden.compare(_golden.run_product(case), case, tol=K.TOL)


@pytest.mark.parametrize("shape,split", [("hex8", False), ("tet4", False), ("hex8", True), ("tet4", True), ("hex8", "prepare")])
def test_two_rank_halo_exchange_single_gpu(shape, split):
    """Mesh-partitioned evaluation (dist.py) emulated on ONE GPU: two in-process plans with HALO batches, the
    interface staging copied between them where NCCL send/recv would carry it.  Owned rows must be BITWISE equal to
    the single-plan result (same contributions, same `elems` order).  Both shapes run on the destination-sorted path with
    export ranges + HALO scatter (hex8: persistent K1; tet4, one Gauss point: the warp-cooperative K1).
    split: the interior rows / pairs are assembled BEFORE the interface rows arrive (ad4sm_eval_assemble_split_device),
    the rest after -- what DistPlan.step does to hide the exchange; the HALO rows hold garbage until then."""
    import torch
    from ad4sm_b200 import dist as AD, mesh
    dims = (7, 6, 10) if split else (4, 3, 6)
    if shape == "hex8":
        coords, conn = mesh.hex_block(*dims, jitter=0.1, seed=21)
        b = El.Hex08_batch(conn, coords, K.nh())
    else:
        coords, conn = mesh.tet_kuhn(*((6, 5, 8) if split else (3, 3, 4)), jitter=0.1, seed=21)
        b = El.Tet04_batch(conn, coords, Ma.MooneyRivlin(800.0, 200.0, 5000.0))
    u = mesh.random_u(len(coords), 3, 1 / 6, seed=22)
    phi0, r0, Kt0 = El.Plan(b, u.shape[1], 3).makeϕrKt(u)
    world = 2
    er = AD.contiguous_elem_ranks(len(conn), world)
    parts = [AD.build_partition(conn, er, g, world) for g in range(world)]
    backs = []
    for p in parts:
        own = El.ElemBatch("CE", 3, p.local_conn(conn[p.own_elems]), b.gradN[p.own_elems], b.wgt[p.own_elems], b.mat)
        backs.append(AD.GpuBackend(own, p.local_conn(conn[p.halo_elems]), p, 3))
    assert all(be.plan.info(capi.INFO_SORTED) == 2 for be in backs)
    if split:   # rank 0 owns the interface nodes: it has an interior to assemble ahead; rank 1 receives nothing
        assert backs[0].plan.info(capi.INFO_SPLIT_PAIRS) > 0
    for p, be in zip(parts, backs):
        ul = torch.from_numpy(np.ascontiguousarray(u[:, p.local_nodes - 1].T)).cuda()
        if split:
            for a in be.staging():
                a[be.n_own:] = float("nan")   # nothing may read the HALO rows before the exchange
        be.eval_elements(ul)
        if split:
            be.assemble_interior()
        be.plan.sync()
    for g, (p, be) in enumerate(zip(parts, backs)):
        for peer, idx in p.send.items():
            s, c = parts[peer].recv[g]
            n0 = backs[peer].n_own
            for src, dst in zip(be.staging(), backs[peer].staging()):
                dst[n0 + s:n0 + s + c] = src[torch.as_tensor(idx, device="cuda")]
    torch.cuda.synchronize()
    phis, seen = [], np.zeros(u.shape[1], dtype=int)
    for p, be in zip(parts, backs):
        if split == "prepare":   # HALO scatter + interface rows of r on the plan's copy stream (ad4sm_eval_halo_prepare_device)
            be.halo_prepare(None)
            be.plan.sync()
        be.assemble()
        be.plan.sync()
        phis.append(float(be.phi_tensor().cpu()[0]))
        Kt = be.plan.fetch_csc()
        r = be.r_tensor().cpu().numpy()
        ln = p.local_nodes
        for loc in np.where(p.owned)[0]:
            gn = ln[loc]
            seen[gn - 1] += 1
            for c in range(3):
                lc, gc = loc * 3 + c, (gn - 1) * 3 + c
                assert r[lc] == r0[gc]
                ls, gs = slice(Kt.colptr[lc] - 1, Kt.colptr[lc + 1] - 1), slice(Kt0.colptr[gc] - 1, Kt0.colptr[gc + 1] - 1)
                lrows = Kt.rowval[ls] - 1
                assert np.array_equal((ln[lrows // 3] - 1) * 3 + lrows % 3 + 1, Kt0.rowval[gs])
                assert np.array_equal(Kt.nzval[ls], Kt0.nzval[gs])
    assert np.all(seen == 1)
    assert abs(sum(phis) - phi0) <= 1e-13 * abs(phi0)


@pytest.mark.parametrize("make", [lambda: K.tet_case((9, 8, 7)), lambda: K.tet_case((4, 3, 3), mat=K.nh()),
                                  lambda: K.tet_case((5, 4, 4), mat=Ma.Ogden(3.0, 1000.0, 5000.0))], ids=["tet4-mr", "tet4-nh", "tet4-ogden"])
def test_one_gauss_point_kernels_agree(make, monkeypatch):
    """P = 1 shapes: the warp-cooperative K1 (k1_u_kernel_w1) on the destination-sorted staging, the same kernel on the
    element-major staging, and the round-1 thread-per-element kernel (AD4SM_K1_T1=1) sum the same contributions in the same
    `elems` order."""
    b, u = make()
    res = []
    for t1, sorted_on in ((0, 1), (0, 0), (1, 0)):
        monkeypatch.setenv("AD4SM_K1_T1", str(t1))
        plan = El.Plan(b, u.shape[1], 3)
        assert plan.info(capi.INFO_SORTED) == 2
        plan.set_option(capi.OPT_SORTED, sorted_on)
        res.append(plan.makeϕrKt(u))
        phi2, r2 = plan.makeϕr(u)
        assert phi2 == res[-1][0] and np.array_equal(r2, res[-1][1])
    f, g, t = res
    # the same kernel on the two staging layouts: every bit
    assert f[0] == g[0] and np.array_equal(f[1], g[1]) and np.array_equal(f[2].nzval, g[2].nzval)
    assert np.array_equal(f[2].colptr, g[2].colptr) and np.array_equal(f[2].rowval, g[2].rowval)
    # the round-1 kernel is a different instantiation of the same source (the compiler contracts other FMAs): rounding
    assert abs(f[0] - t[0]) <= 1e-14 * abs(f[0]) and np.abs(f[1] - t[1]).max() <= 1e-13 * np.abs(f[1]).max()
    assert np.array_equal(f[2].rowval, t[2].rowval) and np.abs(f[2].nzval - t[2].nzval).max() <= 1e-13 * np.abs(f[2].nzval).max()


# ---- Ogden: no reference derivative exists (tests/test_ogden.py pins the tangent against mpmath) ----
def _ogden(alpha=3.0, mu=1000.0, K=5000.0):
    return Ma.Ogden(alpha, mu, K)


@pytest.mark.parametrize("case", ["hex8", "tet4", "quad4", "quad4-K<0"])
def test_ogden_alpha2_equals_neohooke_on_gpu(case):
    """Ogden(α = 2, μ, K) is NeoHooke(μ/2, K) (Σλ² = I1): the spectral-tangent kernels must reproduce the
    invariant-based ones, and those are pinned by the oracle."""
    mk = {"hex8": K.hex_case, "tet4": K.tet_case, "quad4": K.quad_case, "quad4-K<0": K.quad_case}[case]
    Kb = -1.0 if case.endswith("K<0") else 5000.0
    b1, u = mk(mat=Ma.Ogden(2.0, 1000.0, Kb))
    b2, _ = mk(mat=Ma.NeoHooke(500.0, Kb))
    g1 = El.Plan(b1, u.shape[1], u.shape[0]).makeϕrKt(u)
    g2 = El.Plan(b2, u.shape[1], u.shape[0]).makeϕrKt(u)
    assert abs(g1[0] - g2[0]) <= 1e-11 * abs(g2[0])
    assert np.abs(g1[1] - g2[1]).max() <= 1e-10 * np.abs(g2[1]).max()
    assert np.array_equal(g1[2].rowval, g2[2].rowval)
    assert np.abs(g1[2].nzval - g2[2].nzval).max() <= 1e-10 * np.abs(g2[2].nzval).max()


def test_ogden_hex8_contract_properties():
    """α = 3 (no closed-form twin): energy against the reference's scalar formula (oracle restatement), and
    ϕ -> r -> Kt consistency, symmetry, translation null space, determinism."""
    from oracle import elements as OE, materials as OM
    b, u = K.hex_case((4, 3, 3), mat=_ogden())
    plan = El.Plan(b, u.shape[1], 3)
    phi, r, Kt = plan.makeϕrKt(u)
    om = OM.Ogden(3.0, 1000.0, 5000.0)
    phi_ref = sum(float(OE.getphi_elem(OE.Elem("CE", 3, b.nodes[e], b.gradN[e], b.wgt[e], 0.0, om), u[:, b.nodes[e] - 1]))
                  for e in range(b.ne))
    assert abs(phi - phi_ref) <= 1e-12 * abs(phi_ref)
    Kd = K.csc_dense(Kt)
    assert np.array_equal(Kd, Kd.T)
    again = plan.makeϕrKt(u)
    assert again[0] == phi and np.array_equal(again[2].nzval, Kt.nzval)
    for j in range(3):
        t = np.zeros_like(u)
        t[j] = 1.0
        assert np.abs(Kd @ t.reshape(-1, order="F")).max() <= 1e-10 * np.abs(Kd).max()
    rng = np.random.default_rng(0)
    w = rng.uniform(-1, 1, u.shape)
    h = 1e-6 * np.abs(u).max()
    pp, rp, _ = plan.makeϕrKt(u + h * w)
    pm, rm, _ = plan.makeϕrKt(u - h * w)
    wv = w.reshape(-1, order="F")
    assert abs((pp - pm) / (2 * h) - r @ wv) <= 1e-6 * np.abs(r).max() * np.abs(wv).sum() / len(wv) * 10
    assert np.abs((rp - rm) / (2 * h) - Kd @ wv).max() <= 1e-6 * np.abs(Kd @ wv).max()
    # undeformed state: all stretches equal (the spectral formulas' singular point)
    p0, r0, K0 = plan.makeϕrKt(np.zeros_like(u))
    assert abs(p0) <= 1e-12 and np.abs(r0).max() <= 1e-9 * np.abs(K0.nzval).max() and np.all(np.isfinite(K0.nzval))


@pytest.mark.parametrize("make", [lambda: K.hex_case((7, 6, 5)), lambda: K.hex_case((5, 4, 3), mat=Ma.MooneyRivlin(800.0, 200.0, 5000.0)),
                                  lambda: K.hex_case((20, 20, 20))], ids=["hex8-nh", "hex8-mr", "hex8-8000"])
def test_sorted_staging_matches_element_major_staging(make):
    """Destination-sorted staging + streaming assembly (K2s) vs element-major staging + gather assembly (K2): the same
    contributions are summed in the same `elems` order, so every bit must agree."""
    b, u = make()
    plan = El.Plan(b, u.shape[1], 3)
    assert plan.info(capi.INFO_SORTED) == 2
    s = plan.makeϕrKt(u)
    plan.set_option(capi.OPT_SORTED, 0)
    e = plan.makeϕrKt(u)
    assert s[0] == e[0] and np.array_equal(s[1], e[1])
    assert np.array_equal(s[2].colptr, e[2].colptr) and np.array_equal(s[2].rowval, e[2].rowval)
    assert np.array_equal(s[2].nzval, e[2].nzval)
    Kd = s[2].to_scipy()
    assert abs(Kd - Kd.T).max() == 0.0


def test_sorted_staging_with_several_batches():
    """Two materials and two integration rules interleaved in `elems` (three batches keyed by orig_index): the plan still
    gets the destination-sorted staging, its result is bitwise that of the element-major path and matches the oracle
    evaluated element by element in `elems` order."""
    from ad4sm_b200 import mesh
    coords, conn = mesh.hex_block(6, 5, 4, jitter=0.1, seed=41)
    m1, m2 = K.nh(), Ma.MooneyRivlin(800.0, 200.0, 5000.0)

    def make(e):
        if e % 5 == 0:
            return El.Hex08R(conn[e], coords[conn[e] - 1], mat=m1)   # one Gauss point: other P, other K1 kernel... same N
        return El.Hex08(conn[e], coords[conn[e] - 1], mat=(m1 if e % 3 else m2))
    elems = [make(e) for e in range(len(conn))]
    u = mesh.random_u(len(coords), 3, 1 / 6, seed=42)
    plan = El.Plan(elems, len(coords), 3)
    assert len(plan.batches) == 3 and plan.info(capi.INFO_SORTED) == 2
    s = plan.makeϕrKt(u)
    plan.set_option(capi.OPT_SORTED, 0)
    e = plan.makeϕrKt(u)
    assert s[0] == e[0] and np.array_equal(s[1], e[1]) and np.array_equal(s[2].nzval, e[2].nzval)
    assert np.array_equal(s[2].colptr, e[2].colptr) and np.array_equal(s[2].rowval, e[2].rowval)
    ob = [El.ElemBatch("CE", 3, x.nodes[None], x.gradN[None], x.wgt[None], x.mat) for x in elems]
    K.assert_parity(s, K.oracle_eval(ob, u))


def test_pipelined_upload_matches_single_copy(monkeypatch):
    """Host API on a mesh large enough for the pipelined upload (u in node chunks, K1 in element slices): same bits as
    one copy + one launch, for the elastic path and for the phase-field d-phase with history."""
    from ad4sm_b200 import mesh
    b, u = K.hex_case((41, 41, 41))
    piped = El.Plan(b, u.shape[1], 3).makeϕrKt(u)
    monkeypatch.setenv("AD4SM_H2D_SLICES", "0")
    plain_plan = El.Plan(b, u.shape[1], 3)
    monkeypatch.delenv("AD4SM_H2D_SLICES")
    plain = plain_plan.makeϕrKt(u)
    assert piped[0] == plain[0] and np.array_equal(piped[1], plain[1]) and np.array_equal(piped[2].nzval, plain[2].nzval)
    coords, conn = mesh.quad_grid(260, 260, jitter=0.1, seed=7)
    pf = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
    bq = El.QuadP_batch(conn, coords, pf)
    uq = mesh.random_u(len(coords), 2, 1 / 260, a=0.3, seed=8)
    d = K.random_d(uq.shape[1], seed=9)
    h1, h2 = np.zeros((bq.ne, bq.P, 2)), np.zeros((bq.ne, bq.P, 2))
    g1 = El.Plan(bq, uq.shape[1], 2).makeϕrKt_d(uq, d, h1)
    monkeypatch.setenv("AD4SM_H2D_SLICES", "0")
    plain_plan = El.Plan(bq, uq.shape[1], 2)
    monkeypatch.delenv("AD4SM_H2D_SLICES")
    g2 = plain_plan.makeϕrKt_d(uq, d, h2)
    assert g1[0] == g2[0] and np.array_equal(g1[1], g2[1]) and np.array_equal(g1[2].nzval, g2[2].nzval)
    assert np.array_equal(h1, h2) and h1.max() > 0.0


@pytest.mark.parametrize("n", [(41, 41, 41), (6, 5, 4)], ids=["pipelined-upload", "single-copy"])
@pytest.mark.parametrize("declare", [False, True], ids=["element-major", "sorted"])
def test_two_stage_host_api_matches_makephirkt(declare, n):
    """ad4sm_eval_elements_host -> ad4sm_eval_assemble_device -> ad4sm_fetch_phi_r_async (what a partitioned run with
    host-resident u / r calls per rank) against the one-call host API, on a mesh large enough for the pipelined upload."""
    b, u = K.hex_case(n)
    ref = El.Plan(b, u.shape[1], 3).makeϕrKt(u)
    plan = El.Plan(b, u.shape[1], 3)
    if declare:   # a caller that reads no staging rows between the stages gets the destination-sorted path
        capi.check(capi.lib().ad4sm_export_ranges(plan._h, 0, None, None))
    r, phi = np.full(u.size, np.nan), np.full(1, np.nan)
    for _ in range(2):   # twice: the second upload must wait for nothing stale
        plan.eval_elements_host(capi.MODE_ELASTIC, u)
        plan.eval_assemble_device(capi.MODE_ELASTIC)
        plan.fetch_phi_r_async(r, phi)
        plan.sync()
        assert phi[0] == ref[0] and np.array_equal(r, ref[1])
        r[:] = np.nan
    Kt = plan.fetch_csc()
    assert np.array_equal(Kt.nzval, ref[2].nzval) and np.array_equal(Kt.rowval, ref[2].rowval)


@pytest.mark.parametrize("case", ["hex8", "quad4", "tet4", "hex8-keyed"])
def test_assembly_only_entry_is_bit_exact(case):
    """makeϕrKt(Φ, elems, u) (elements.toolkit.jl:169-203): element duals computed by the oracle, assembled by the GPU
    (ad4sm_assemble_duals) and by the oracle's serial COO->CSC -- same contributions, same left fold in `elems` order,
    so r and Kt must agree bit for bit (ϕ: the device sums a fixed tree, the reference a left fold)."""
    b, u = {"hex8": lambda: K.hex_case((6, 5, 4)), "quad4": lambda: K.quad_case((9, 7)), "tet4": lambda: K.tet_case((4, 3, 3)),
            "hex8-keyed": lambda: K.hex_case((5, 4, 3))}[case]()
    val, g, H = X.elem_duals(b.kind, b.D, b.P, b.N, b.mat, b.nodes, b.gradN, b.wgt, b.Nval, X.M_ELASTIC, u)
    nodes = b.nodes
    if case == "hex8-keyed":   # the batch lists the elements in another order than `elems`: orig_index maps slots to elems
        perm = np.random.default_rng(5).permutation(b.ne)
        b = El.ElemBatch("CE", 3, b.nodes, b.gradN, b.wgt, b.mat, orig_index=perm.astype(np.int64))
        inv = np.argsort(perm)
        val, g, H, nodes = val[inv], g[inv], H[inv], nodes[inv]   # row k = element at position k of `elems`
    phi0, r0, (cp0, rv0, nz0) = X.assemble(nodes, b.D, val, g, H, u.size)
    plan = El.Plan(b, u.shape[1], b.D)
    phi, r, Kt = plan.assemble_duals(El.pack_duals(val, g, H))
    assert abs(phi - phi0) <= 1e-13 * max(1.0, abs(phi0))
    assert np.array_equal(r, r0)
    assert np.array_equal(Kt.colptr, cp0) and np.array_equal(Kt.rowval, rv0) and np.array_equal(Kt.nzval, nz0)
    # makeϕr(Φ, elems, u): D1-shaped records (v | g)
    phi1, r1 = plan.assemble_duals_phi_r(np.ascontiguousarray(np.concatenate([val[:, None], g], axis=1)))
    assert abs(phi1 - phi0) <= 1e-13 * max(1.0, abs(phi0)) and np.array_equal(r1, r0)
    # and the assembled result is the one the full path gives (to 1e-12: other arithmetic inside the elements)
    K.assert_parity(plan.makeϕrKt(u), (phi0, r0, (cp0, rv0, nz0)))


def test_sorted_assembly_with_many_contributions_per_pair():
    """Every element listed 20 times (a legal `elems`): each node pair then has up to 160 contributions and a warp's range
    of the sorted staging exceeds its shared-memory slice, so K2s takes its direct-load path.  Same bits as the
    element-major assembly, and parity with the oracle."""
    b, u = K.hex_case((3, 3, 2))
    rep = np.repeat(np.arange(b.ne), 20)
    bb = El.ElemBatch("CE", 3, b.nodes[rep], b.gradN[rep], b.wgt[rep], b.mat)
    plan = El.Plan(bb, u.shape[1], 3)
    assert plan.info(capi.INFO_SORTED) == 2
    s = plan.makeϕrKt(u)
    K.assert_parity(s, K.oracle_eval(bb, u))
    plan.set_option(capi.OPT_SORTED, 0)
    e = plan.makeϕrKt(u)
    assert s[0] == e[0] and np.array_equal(s[1], e[1]) and np.array_equal(s[2].nzval, e[2].nzval)


def test_isolated_nodes_and_shuffled_elements():
    """Nodes that no element touches give empty CSC columns and zero residual entries; a shuffled `elems` order changes
    only the summation order (results still match the oracle evaluated in that same order)."""
    b, u = K.hex_case((4, 3, 3))
    extra = 5
    u2 = np.asfortranarray(np.concatenate([u, np.full((3, extra), 0.123)], axis=1))   # 5 unused trailing nodes
    perm = np.random.default_rng(3).permutation(b.ne)
    bs = El.ElemBatch("CE", 3, b.nodes[perm], b.gradN[perm], b.wgt[perm], b.mat)
    got = El.Plan(bs, u2.shape[1], 3).makeϕrKt(u2)
    ref = K.oracle_eval(bs, u2)
    K.assert_parity(got, ref)
    n0 = u.shape[1] * 3
    assert np.all(got[1][n0:] == 0.0)
    assert np.all(np.diff(got[2].colptr[n0:]) == 0), "untouched dofs must have empty columns"


def test_ragged_element_list_is_rejected():
    """All elements must have the same number of dofs (elasticelements.jl:212-214)."""
    from ad4sm_b200 import mesh
    bh, u = K.hex_case((2, 2, 2))
    coords, conn = mesh.tet_kuhn(2, 2, 2)
    bt = El.Tet04_batch(conn, coords, K.nh())
    with pytest.raises(capi.AD4SMError):
        El.Plan([bh, bt], max(u.shape[1], len(coords)), 3)


# ---- CAS: axisymmetric continuum elements (extension; oracle pins: tests/test_oracle_pins.py §6) ----------
def _cas_case(ctor, n, mat, seed):
    from ad4sm_b200 import mesh
    grid = mesh.quad_grid if ctor is El.ASQuad_batch else mesh.tri_grid
    coords, conn = grid(*n, jitter=0.1, seed=seed)
    coords = coords + np.array([0.6, 0.0])   # off the axis
    b = ctor(conn, coords, mat)
    u = mesh.random_u(len(coords), 2, 1.0 / max(n), a=0.03, seed=seed + 100)
    return b, u, coords


@pytest.mark.parametrize("ctor,mat", [
    (El.ASQuad_batch, K.nh()), (El.ASQuad_batch, Ma.Hooke(210000.0, 0.3)), (El.ASQuad_batch, Ma.Hooke(210000.0, 0.3, small=True)),
    (El.ASQuad_batch, Ma.MooneyRivlin(800.0, 200.0, 5000.0)), (El.ASTria_batch, K.nh()),
    (El.ASTria_batch, Ma.MooneyRivlin(800.0, 200.0, 5000.0)), (El.ASQuad_batch, Ma.Ogden(3.0, 1000.0, 5000.0))],
    ids=["asquad-nh", "asquad-hooke-gl", "asquad-hooke-small", "asquad-mr", "astria-nh", "astria-mr", "asquad-ogden"])
def test_cas_parity(ctor, mat):
    b, u, _ = _cas_case(ctor, (11, 9), mat, seed=51)
    plan = El.Plan(b, u.shape[1], 2)
    got = plan.makeϕrKt(u)
    tol = 1e-11 if type(mat).__name__ == "Ogden" else K.TOL   # Ogden: two formulations
    K.assert_parity(got, K.oracle_eval(b, u), tol=tol)
    Kd = got[2].to_scipy()
    assert abs(Kd - Kd.T).max() == 0.0
    phi2, r2 = plan.makeϕr(u)
    assert phi2 == got[0] and np.array_equal(r2, got[1])
    again = plan.makeϕrKt(u)
    assert again[0] == got[0] and np.array_equal(again[1], got[1]) and np.array_equal(again[2].nzval, got[2].nzval)


def test_cas_known_answers_and_errors():
    """homogeneous expansion u_r = ε r, u_z = γ z: ϕ = ϕ(diag(1+ε, 1+γ, 1+ε)) · 2π∫r dA on the GPU; element objects built
    with the reference's constructor names group into the same plan; unsupported combinations are rejected."""
    mat = K.nh()
    b, _, coords = _cas_case(El.ASQuad_batch, (6, 5), mat, seed=52)
    eps, gam = 0.03, -0.02
    u = np.asfortranarray(np.stack([eps * coords[:, 0], gam * coords[:, 1]]))
    phi, r, Kt = El.Plan(b, len(coords), 2).makeϕrKt(u)
    J = (1 + eps) ** 2 * (1 + gam)
    I1 = 2 * (1 + eps) ** 2 + (1 + gam) ** 2
    psi = mat.C1 * (I1 * J ** (-2.0 / 3.0) - 3) + mat.K * (J - 1) ** 2   # elasticmaterials.jl:186-197
    assert abs(phi - psi * b.wgt.sum()) <= 1e-12 * abs(phi)
    elems = [El.ASQuad(b.nodes[i], coords[b.nodes[i] - 1], mat=mat) for i in range(b.ne)]
    assert all(type(e).__name__ == "CAS" for e in elems)
    phi2, r2, Kt2 = El.Plan(elems, len(coords), 2).makeϕrKt(u)
    assert phi2 == phi and np.array_equal(r2, r) and np.array_equal(Kt2.nzval, Kt.nzval)
    with pytest.raises(capi.AD4SMError):   # a 2-D material on the 3x3 F of an axisymmetric element
        El.Plan(El.ASQuad_batch(b.nodes, coords, Ma.Hooke2D(210000.0, 0.3)), len(coords), 2)
    with pytest.raises(capi.AD4SMError):   # no post-processing for CAS
        El.Plan(b, len(coords), 2).post(u, F=True)
