"""Batched post-processing on the device against the literal restatement of the reference's per-element
functions (oracle/post.py = elements.toolkit.jl:8-132): getF, detJ, getV, getσ, getP."""
import numpy as np
import pytest

from ad4sm_b200 import Elements as El, Materials as Ma
from oracle import elements as OE, materials as OM, post as OP

import _cases as K

pytestmark = pytest.mark.gpu

CASES = [
    ("hex8-neohooke", lambda: K.hex_case((3, 3, 2)), OM.NeoHooke(1000.0, 5000.0)),
    ("hex8-mooney", lambda: K.hex_case((3, 2, 2), mat=Ma.MooneyRivlin(800.0, 200.0, 5000.0)), OM.MooneyRivlin(800.0, 200.0, 5000.0)),
    ("hex8-hooke", lambda: K.hex_case((3, 2, 2), mat=Ma.Hooke(210000.0, 0.3)), OM.Hooke(210000.0, 0.3)),
    ("tet4-mooney", lambda: K.tet_case((2, 2, 2)), OM.MooneyRivlin(800.0, 200.0, 5000.0)),
    ("quad4-hooke2d", lambda: K.quad_case((5, 4)), OM.Hooke2D(210000.0, 0.3, plane_stress=False)),
    ("quad4-neohooke", lambda: K.quad_case((5, 4), mat=K.nh()), OM.NeoHooke(1000.0, 5000.0)),
    ("tria3-hooke2d", lambda: K.tri_case((4, 3)), OM.Hooke2D(210000.0, 0.3)),
]


@pytest.mark.parametrize("name,make,omat", CASES, ids=[c[0] for c in CASES])
def test_post_processing_matches_reference_functions(name, make, omat):
    b, u = make()
    got = El.Plan(b, u.shape[1], u.shape[0]).post(u, F=True, detJ=True, V=True, sigma=True, P=True)
    ref = {k: [] for k in ("F", "detJ", "V", "sigma", "P")}
    for e in range(b.ne):
        el = OE.Elem("CE", b.D, b.nodes[e], b.gradN[e], b.wgt[e], float(b.wgt[e].sum()), omat)
        ue = u[:, b.nodes[e] - 1]
        ref["F"].append(OP.getF(el, ue))
        ref["detJ"].append(OP.detJ(el, ue))
        ref["V"].append(OP.getV(el, ue))
        ref["sigma"].append(OP.getsigma(el, ue))
        ref["P"].append(OP.getP(el, ue))
    for k in ref:
        r = np.array(ref[k])
        assert got[k].shape == r.shape, k
        assert np.abs(got[k] - r).max() <= 1e-12 * max(np.abs(r).max(), 1e-300), (k, np.abs(got[k] - r).max() / np.abs(r).max())


def test_reference_named_array_methods():
    b, u = K.hex_case((3, 2, 2))
    F = El.getF(b, u)
    assert F.shape == (b.ne, 3, 3)
    assert np.allclose(El.get_grad_u(b, u), F - np.eye(3))
    s6 = El.getσ(b, u)
    s = El.Plan(b, u.shape[1], 3).post(u, sigma=True)["sigma"]
    assert s6.shape == (6, b.ne)
    assert np.array_equal(s6[:, 0], [s[0, 0, 0], s[0, 1, 1], s[0, 2, 2], s[0, 1, 0], s[0, 2, 0], s[0, 2, 1]])
    assert abs(El.getV(b, u) - El.Plan(b, u.shape[1], 3).post(u, V=True)["V"].sum()) <= 1e-13
    # undeformed: F = I, detJ = 1, σ = 0, V = reference volume
    z = np.zeros_like(u)
    p0 = El.Plan(b, u.shape[1], 3).post(z, F=True, detJ=True, V=True, sigma=True)
    assert np.abs(p0["F"] - np.eye(3)).max() == 0.0 and np.abs(p0["detJ"] - 1).max() <= 1e-15
    assert np.abs(p0["sigma"]).max() <= 1e-9 and abs(p0["V"].sum() - b.wgt.sum()) <= 1e-13
    q, u2 = K.quad_case((4, 3))
    assert El.getσ(q, u2).shape == (3, q.ne)


def test_post_rejects_phasefield_stress():
    from ad4sm_b200 import _capi as capi, mesh
    coords, conn = mesh.quad_grid(3, 3, jitter=0.1, seed=1)
    pf = Ma.PhaseField(0.01, 100.0, Ma.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
    b = El.QuadP_batch(conn, coords, pf)
    plan = El.Plan(b, len(coords), 2)
    u = np.zeros((2, len(coords)), order="F")
    assert plan.post(u, F=True, V=True)["V"].shape == (b.ne,)
    with pytest.raises(capi.AD4SMError):
        plan.post(u, sigma=True)
