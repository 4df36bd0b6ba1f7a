"""Golden-vector helpers (test infrastructure).  A golden case is an .npz holding the SoA inputs of one
`makeϕrKt` / `makeϕrKt_d` call and the outputs of the literal Python restatement of the reference
(oracle/*.py), written by tests/golden/make_golden.py."""
import json

import numpy as np

from oracle.check import assert_values_entrywise  # noqa: F401  (re-exported for _cases)

MODES = {"elastic": 0, "pf_u": 1, "pf_d": 2}


class Case:
    pass


def _mat(mod, spec):
    spec = dict(spec)
    cls = getattr(mod, spec.pop("cls"))
    if cls.__name__ == "PhaseField":
        base = _mat(mod, spec.pop("mat"))
        return cls(spec["l0"], spec["Gc"], base, spec["n"], AT=spec["AT"])
    return cls(**spec)


def mat_spec(m):
    n = type(m).__name__
    if n == "PhaseField":
        return {"cls": n, "l0": m.l0, "Gc": m.Gc, "n": m.n, "AT": m.AT, "mat": mat_spec(m.mat)}
    keys = {"Hooke": ("E", "nu", "small"), "Hooke2D": ("E", "nu", "small", "plane_stress"), "NeoHooke": ("C1", "K"),
            "MooneyRivlin": ("C1", "C2", "K")}[n]
    return {"cls": n, **{k: getattr(m, k) for k in keys}}


def load(path):
    z = np.load(path, allow_pickle=False)
    c = Case()
    c.meta = json.loads(str(z["meta"]))
    for k in ("nodes", "gradN", "wgt", "u", "phi", "r", "colptr", "rowval", "nzval"):
        setattr(c, k, z[k])
    c.Nval = z["Nval"] if "Nval" in z.files else None
    c.X0 = z["X0"] if "X0" in z.files else None
    c.d = z["d"] if "d" in z.files else None
    c.hist_in = z["hist_in"] if "hist_in" in z.files else None
    c.hist_out = z["hist_out"] if "hist_out" in z.files else None
    c.kind, c.D, c.mode = c.meta["kind"], int(c.meta["D"]), c.meta["mode"]
    c.opts = c.meta.get("opts", {})
    c.N, c.P = c.nodes.shape[1], c.wgt.shape[1]
    return c


def run_oracle(c):
    """C++ restatement on the golden inputs -> (ϕ, r, colptr, rowval, nzval, hist)."""
    from oracle import cxx_binding as X, materials as M
    c.mat = _mat(M, c.meta["mat"])
    hist = None if c.hist_in is None else c.hist_in.copy()
    phi, r, (cp, rv, nz) = X.make_phi_r_Kt([c], np.asfortranarray(c.u), mode=MODES[c.mode], d=c.d,
                                           hist=None if hist is None else [hist], **c.opts)
    return phi, r, cp, rv, nz, hist


def run_product(c):
    """The CUDA path through the C ABI on the golden inputs."""
    from ad4sm_b200 import Elements as El, Materials as Ma
    b = El.ElemBatch(c.kind, c.D, c.nodes, c.gradN, c.wgt, _mat(Ma, c.meta["mat"]), c.Nval, X0=c.X0)
    popts = {}
    if "grad_term" in c.opts:
        popts["pf_grad_term"] = c.opts["grad_term"]
    if "nh2d_extension" in c.opts:
        popts["nh2d_extension"] = c.opts["nh2d_extension"]
    plan = El.Plan(b, c.u.shape[1], c.u.shape[0], **popts)
    hist = None if c.hist_in is None else np.ascontiguousarray(c.hist_in.copy())
    if c.mode == "elastic":
        phi, r, Kt = plan.makeϕrKt(c.u)
    elif c.mode == "pf_u":
        phi, r, Kt = plan.makeϕrKt(c.u, c.d)
    else:
        phi, r, Kt = plan.makeϕrKt_d(c.u, c.d, hist)
    return phi, r, Kt.colptr, Kt.rowval, Kt.nzval, hist


def compare(got, c, tol):
    phi, r, cp, rv, nz, hist = got
    assert np.array_equal(cp, c.colptr), "colptr differs from the golden vector"
    assert np.array_equal(rv, c.rowval), "rowval differs from the golden vector"
    assert abs(phi - float(c.phi)) <= tol * max(1.0, abs(float(c.phi)))
    assert np.abs(r - c.r).max() <= tol * np.abs(c.r).max()
    assert_values_entrywise(nz, c.nzval, c.colptr, c.rowval, tol)
    if c.hist_out is not None:
        assert np.abs(hist - c.hist_out).max() <= tol * max(1.0, np.abs(c.hist_out).max())
