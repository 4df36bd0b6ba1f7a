"""Batched element constructors on the device (csrc/construct.cu, ad4sm_construct / ad4sm_batch.mesh)
against the oracle's literal per-element restatement of the reference constructors (oracle/elements.py: D1 duals of the
shape functions, J \\ ∂N/∂ξ, detJ·w -- elasticelements.jl:35-197, phasefieldelements.jl:9-145) and against the host
(numpy) constructors the other tests build their inputs with.

Tolerance: ∇N, wgt, N, V to 1e-13 relative to the array scale -- the device solves the D x D system by Gaussian
elimination with partial pivoting like `\\`, but the products are contracted to FMAs, so the bits may differ in the last
place; what the hot path does with the result is covered by the 1e-12 parity of makeϕrKt below."""
import numpy as np
import pytest

from ad4sm_b200 import Elements as El, Materials as Ma, mesh
from oracle import elements as OE

from _cases import assert_parity, oracle_eval, nh

pytestmark = pytest.mark.gpu

GP3 = ((-0.774596669241483, 5.0 / 9.0), (0.0, 8.0 / 9.0), (0.774596669241483, 5.0 / 9.0))
TOL = 1e-13


def _close(a, b, what):
    scale = np.abs(b).max()
    assert np.abs(a - b).max() <= TOL * scale, f"{what}: max |Δ| {np.abs(a - b).max():.3e} at scale {scale:.3e}"


def _oracle_batch(octor, conn, coords, D, **kw):
    els = [octor(conn[e], [coords[n - 1, :D] for n in conn[e]], None, **kw) for e in range(len(conn))]
    return els


HOOKE2D = Ma.Hooke2D(210000.0, 0.3)
PF2 = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(1000.0, 0.3), 2)
PF3 = Ma.PhaseField(0.01, 100.0, Ma.Hooke(1000.0, 0.3), 2)

CASES = [
    # (device ctor, host ctor, oracle ctor, mesh, D, kwargs)
    ("Tria", El.Tria_batch, OE.Tria, lambda: mesh.tri_grid(7, 5, jitter=0.2, seed=11), 2, {}, HOOKE2D),
    ("TriaP", El.TriaP_batch, OE.TriaP, lambda: mesh.tri_grid(4, 6, jitter=0.2, seed=12), 2, {}, PF2),
    ("Quad", El.Quad_batch, OE.Quad, lambda: mesh.quad_grid(8, 5, jitter=0.2, seed=13), 2, {}, HOOKE2D),
    ("QuadR", El.Quad_batch, OE.Quad, lambda: mesh.quad_grid(5, 5, jitter=0.2, seed=14), 2, {"GP": El.GP1}, HOOKE2D),
    ("Quad3", El.Quad_batch, OE.Quad, lambda: mesh.quad_grid(4, 3, jitter=0.2, seed=15), 2, {"GP": GP3}, HOOKE2D),
    ("QuadP", El.QuadP_batch, OE.QuadP, lambda: mesh.quad_grid(6, 4, jitter=0.2, seed=16), 2, {}, PF2),
    ("Tet04", El.Tet04_batch, OE.Tet04, lambda: mesh.tet_kuhn(3, 3, 2, jitter=0.2, seed=17), 3, {}, nh()),
    ("Tet04P", El.Tet04P_batch, OE.Tet04P, lambda: mesh.tet_kuhn(2, 3, 2, jitter=0.2, seed=18), 3, {}, PF3),
    ("Hex08", El.Hex08_batch, OE.Hex08, lambda: mesh.hex_block(4, 3, 3, jitter=0.2, seed=19), 3, {}, nh()),
    ("Hex08R", El.Hex08_batch, OE.Hex08, lambda: mesh.hex_block(3, 3, 3, jitter=0.2, seed=20), 3, {"GP": El.GP1}, nh()),
    ("Hex08_3", El.Hex08_batch, OE.Hex08, lambda: mesh.hex_block(3, 2, 2, jitter=0.2, seed=21), 3, {"GP": GP3}, nh()),
    ("Hex08P", El.Hex08P_batch, OE.Hex08P, lambda: mesh.hex_block(3, 3, 2, jitter=0.2, seed=22), 3, {}, PF3),
]


@pytest.mark.parametrize("name,ctor,octor,mk,D,kw,mat", CASES, ids=[c[0] for c in CASES])
def test_device_constructor_matches_reference_constructor(name, ctor, octor, mk, D, kw, mat):
    coords, conn = mk()
    dev = ctor(conn, coords, mat, device=True, **kw)
    assert isinstance(dev, El.MeshBatch)
    got = dev.materialize()
    host = ctor(conn, coords, mat, **kw)
    assert got.gradN.shape == host.gradN.shape and got.wgt.shape == host.wgt.shape
    _close(got.gradN, host.gradN, "∇N vs host constructor")
    _close(got.wgt, host.wgt, "wgt vs host constructor")
    ref = _oracle_batch(octor, conn, coords, D, **kw)
    _close(got.gradN, np.stack([e.gradN for e in ref]), "∇N vs oracle constructor")
    _close(got.wgt, np.stack([e.wgt for e in ref]), "wgt vs oracle constructor")
    _close(got.V, np.array([e.V for e in ref]), "V vs oracle constructor")
    if got.Nval is not None:
        _close(got.Nval, np.stack([e.Nval for e in ref]), "N vs oracle constructor")
        assert np.array_equal(got.Nval, host.Nval)
    else:
        assert ref[0].Nval is None


def test_wedge_constructor():
    """Wdg06 / Wdg06P: prisms from splitting every hex of a jittered block along the (0,2)/(4,6) diagonal."""
    coords, hexc = mesh.hex_block(3, 2, 2, jitter=0.15, seed=23)
    conn = np.concatenate([hexc[:, [0, 1, 2, 4, 5, 6]], hexc[:, [0, 2, 3, 4, 6, 7]]])
    got = El.Wdg06P_batch(conn, coords, PF3, device=True).materialize()
    host = El.Wdg06P_batch(conn, coords, PF3)
    ref = _oracle_batch(OE.Wdg06P, conn, coords, 3)
    _close(got.gradN, host.gradN, "∇N vs host constructor")
    _close(got.gradN, np.stack([e.gradN for e in ref]), "∇N vs oracle constructor")
    _close(got.wgt, np.stack([e.wgt for e in ref]), "wgt")
    _close(got.Nval, np.stack([e.Nval for e in ref]), "N")
    _close(got.V, np.array([e.V for e in ref]), "V")
    assert (got.wgt > 0).all()
    ce = El.Wdg06_batch(conn, coords, nh(), device=True).materialize()
    assert ce.Nval is None and np.array_equal(ce.gradN, got.gradN)


def test_signed_volume_and_orientation():
    """Tet04 keeps the sign of det(A)/6 (elasticelements.jl:92): a negatively oriented tet gives wgt < 0."""
    coords, conn = mesh.tet_kuhn(2, 2, 2, jitter=0.1, seed=24)
    flipped = conn[:, [0, 2, 1, 3]]
    a = El.Tet04_batch(conn, coords, nh(), device=True).materialize()
    b = El.Tet04_batch(flipped, coords, nh(), device=True).materialize()
    assert (a.wgt > 0).all() and (b.wgt < 0).all()
    _close(b.wgt, -a.wgt, "signed volume")


@pytest.mark.parametrize("shape", ["hex", "quad", "tet", "quadp"])
def test_plan_from_coordinates_parity(shape):
    """makeϕrKt on a plan whose ∇N / wgt never existed on the host (ad4sm_batch.mesh) against the oracle evaluated on the
    host constructors' arrays: the 1e-12 contract of the hot path."""
    from oracle import cxx_binding as X
    if shape == "hex":
        coords, conn = mesh.hex_block(6, 5, 4, jitter=0.1, seed=31)
        mat, ctor, dpn = nh(), El.Hex08_batch, 3
    elif shape == "quad":
        coords, conn = mesh.quad_grid(12, 9, jitter=0.1, seed=32)
        mat, ctor, dpn = Ma.Hooke2D(210000.0, 0.3, plane_stress=False), El.Quad_batch, 2
    elif shape == "tet":
        coords, conn = mesh.tet_kuhn(4, 4, 3, jitter=0.1, seed=33)
        mat, ctor, dpn = Ma.MooneyRivlin(800.0, 200.0, 5000.0), El.Tet04_batch, 3
    else:
        coords, conn = mesh.quad_grid(9, 8, jitter=0.1, seed=34)
        mat, ctor, dpn = PF2, El.QuadP_batch, 2
    u = mesh.random_u(len(coords), dpn, 0.1, a=0.02, seed=35)
    dev = ctor(conn, coords, mat, device=True)
    host = ctor(conn, coords, mat)
    plan = El.Plan(dev, len(coords), dpn)
    if shape == "quadp":
        d = np.random.default_rng(36).uniform(0, 0.5, len(coords))
        got = plan.makeϕrKt(u, d)
        ref = oracle_eval(host, u, mode=X.M_PF_U, d=d)
    else:
        got = plan.makeϕrKt(u)
        ref = oracle_eval(host, u)
    assert_parity(got, ref)
    plan.close()


def test_constructor_errors():
    coords, conn = mesh.hex_block(2, 2, 2)
    with pytest.raises(ValueError):
        El.Hex08_batch(conn[:, :6], coords, nh(), device=True)
    with pytest.raises(ValueError):
        El.Hex08_batch(conn, coords, nh(), GP=((0.0, 1.0),) * 4, device=True)
    bad = conn.copy()
    bad[0, 0] = len(coords) + 1
    with pytest.raises(Exception, match="node id out of range"):
        El.Hex08_batch(bad, coords, nh(), device=True).materialize()
    with pytest.raises(Exception, match="n_nodes"):
        El.Plan(El.Hex08_batch(conn, coords, nh(), device=True), len(coords) + 3, 3)
