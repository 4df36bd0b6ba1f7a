#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference (pure Julia) cannot be executed in the build container, so the golden outputs come from the
LITERAL Python restatement of the reference (oracle/adiff.py, materials.py, elements.py: the same dual-number
operations in the same association order, file:line cited there) on small seeded jittered meshes built with
the restated constructors.  The C++ oracle and the CUDA path are both checked against these files.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import elements as E, materials as M  # noqa: E402
import _golden  # noqa: E402
import test_oracle_pins as T  # noqa: E402  (mesh builders)


def save(name, elems, u, mode="elastic", d=None, hist=None, **opts):
    b = T._Batch(elems)
    hist_in = None if hist is None else np.array(hist, dtype=float)
    if mode == "elastic":
        phi, r, (cp, rv, nz) = E.make_phi_r_Kt(elems, u)
    elif mode == "pf_u":
        phi, r, (cp, rv, nz) = E.make_phi_r_Kt_pf_u(elems, u, d, **opts)
    else:
        phi, r, (cp, rv, nz) = E.make_phi_r_Kt_pf_d(elems, u, d, hist, **opts)
    out = dict(meta=json.dumps({"kind": b.kind, "D": b.D, "mode": mode, "mat": _golden.mat_spec(b.mat), "opts": opts}),
               nodes=b.nodes, gradN=b.gradN, wgt=b.wgt, u=np.asfortranarray(u), phi=phi, r=r, colptr=cp, rowval=rv, nzval=nz)
    if b.Nval is not None:
        out["Nval"] = b.Nval
    if getattr(b, "X0", None) is not None:
        out["X0"] = b.X0
    if d is not None:
        out["d"] = d
    if hist is not None:
        out["hist_in"] = hist_in
        out["hist_out"] = np.array(hist, dtype=float)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: ne={len(elems)} nnz={len(nz)} phi={phi:.6e}")


def main():
    NH, MR = M.NeoHooke(1000.0, 5000.0), M.MooneyRivlin(800.0, 200.0, 5000.0)
    e, X0 = T.hex_mesh((2, 2, 2), NH, seed=11)
    save("hex8_neohooke", e, T._u(X0, 3, seed=12))
    e, X0 = T.hex_mesh((2, 2, 1), MR, seed=13)
    save("hex8_mooney", e, T._u(X0, 3, seed=14))
    e, X0 = T.hex_mesh((2, 1, 1), M.Hooke(210000.0, 0.3), seed=15)
    save("hex8_hooke_gl", e, T._u(X0, 3, seed=16))
    e, X0 = T.tet_mesh(MR, seed=17)
    save("tet4_mooney", e, T._u(X0, 3, seed=18))
    e, X0 = T.quad_mesh((4, 3), seed=19)
    save("quad4_hooke2d_pe", e, T._u(X0, 2, seed=20))
    e, X0 = T.quad_mesh((3, 3), M.NeoHooke(1000.0, -1.0), seed=21)
    save("quad4_neohooke_planestress", e, T._u(X0, 2, seed=22))
    e, X0 = T.quad_mesh((3, 2), seed=23)
    save("quad4_hooke2d_u3rows", e, T._u(X0, 3, seed=24))  # 3-row u: exact zeros dropped (toolkit:13,202)
    pf = M.PhaseField(0.01, 100.0, M.Hooke2D(210000.0, 0.3, plane_stress=False), 2)
    e, X0 = T.quad_mesh((3, 3), pf, seed=25, ctor=E.QuadP)
    u, d = T._u(X0, 2, amp=0.05, seed=26), np.random.default_rng(27).uniform(0, 0.5, len(X0))
    save("quadp_pf_hooke2d_u", e, u, "pf_u", d)
    save("quadp_pf_hooke2d_d", e, u, "pf_d", d, grad_term="reference")
    h = [[(0.3 * (i % 2), 0.1) for _ in range(4)] for i in range(len(e))]
    save("quadp_pf_hooke2d_d_hist", e, u, "pf_d", d, h, grad_term="reference")
    pf3 = M.PhaseField(0.01, 100.0, NH, 2)
    e, X0 = T.hex_mesh((2, 1, 1), pf3, seed=28, ctor=E.Hex08P)
    u, d = T._u(X0, 3, amp=0.05, seed=29), np.random.default_rng(30).uniform(0, 0.5, len(X0))
    save("hex8p_pf_neohooke_u", e, u, "pf_u", d)
    save("hex8p_pf_neohooke_d", e, u, "pf_d", d, grad_term="intended")
    # axisymmetric CAS (extension): meshes shifted off the axis (r in [0.7, 1.7])
    e, X0 = T.quad_mesh((3, 3), NH, seed=31, ctor=E.ASQuad, shift=(0.7, 0.0))
    save("asquad_neohooke", e, T._u(X0, 2, seed=32))
    e, X0 = T.quad_mesh((3, 2), M.Hooke(210000.0, 0.3), seed=33, ctor=E.ASQuad, shift=(0.7, 0.0))
    save("asquad_hooke_gl", e, T._u(X0, 2, seed=34))
    e, X0 = T.tri_mesh((3, 3), MR, seed=35, ctor=E.ASTria, shift=(0.7, 0.0))
    save("astria_mooney", e, T._u(X0, 2, seed=36))


if __name__ == "__main__":
    main()
