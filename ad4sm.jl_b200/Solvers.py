"""Host-side mirror of `Solvers.solvestep!` (/root/reference/src/solvers.jl:172-300), nEqs == 0 branch, on top of the
device-side Newton glue (ad4sm_newton_*).

What runs where
  * `makeϕrKt(elems, u)` (solvers.jl:213-216, 257-260)                              -> the plan, on the GPU;
  * `fi[ifreeu]-fe[ifreeu]`, `Kt[ifreeu,icnstu]*Δucnst`, `Kt[ifreeu,ifreeu]`,
    `norm(fi[icnstu])`, `norm(res)` (:220-221, 262-265, 285)                         -> ad4sm_newton_reduce, on the GPU;
  * `lu(Kt[ifreeu,ifreeu]) \\ res` (:223, 285)                                        -> OUT OF SCOPE of the product
    (north_star): it stays on a host sparse direct solver.  In Julia that is UMFPACK's `lu`; here, where the host
    language is Python, it is `scipy.sparse.linalg.splu` -- passed in as `solve`, so any other solver can be used.

Same argument names, defaults and control flow as the reference.

Constraint equations (`ConstEq`, `makeϕrKt(eqns, u, λ)`, solvers.jl:17-73): the reference stores an arbitrary closure per
equation, which cannot cross a C ABI.  Offered instead is a CATALOGUE -- every constraint that is a polynomial of degree
<= 2 in its dofs, ϕ(v) = ½ vᵀQv + cᵀv - c0 with v = u[iDoFs] (`ConstEq.linear`, `.quadratic`, and the usual special cases
`.periodic`, `.tie`, `.average`, `.distance`) -- evaluated on the GPU by `ad4sm_consteq_eval` and assembled into the same
three results the reference returns: veqs, reqs (nDoFs x nEqs sparse), Keqs (nDoFs x nDoFs sparse).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi


@dataclass
class ConstEq:
    """`ConstEq(func, iDoFs)` (solvers.jl:17-22) restricted to the catalogue ϕ(v) = ½ vᵀQv + cᵀv - c0, v = u[iDoFs].
    iDoFs: 1-based linear indices into u[:] (column-major, like the reference)."""
    iDoFs: np.ndarray
    c: np.ndarray
    c0: float = 0.0
    Q: np.ndarray | None = None

    def __post_init__(self):
        self.iDoFs = np.ascontiguousarray(self.iDoFs, dtype=np.int64).reshape(-1)
        self.c = np.ascontiguousarray(self.c, dtype=np.float64).reshape(-1)
        n = len(self.iDoFs)
        if len(self.c) != n:
            raise ValueError("ConstEq: c and iDoFs must have the same length")
        if self.Q is not None:
            self.Q = np.asfortranarray(self.Q, dtype=np.float64)
            if self.Q.shape != (n, n) or not np.array_equal(self.Q, self.Q.T):
                raise ValueError("ConstEq: Q must be a symmetric n x n matrix")

    # ---- the catalogue --------------------------------------------------------------------------------------------
    @classmethod
    def linear(cls, iDoFs, c, c0=0.0):
        """Σ c_k u_k = c0"""
        return cls(iDoFs, c, c0)

    @classmethod
    def quadratic(cls, iDoFs, Q, c=None, c0=0.0):
        """½ vᵀQv + cᵀv = c0"""
        n = len(np.atleast_1d(iDoFs))
        return cls(iDoFs, np.zeros(n) if c is None else c, c0, Q)

    @classmethod
    def periodic(cls, dof_a, dof_b, jump=0.0):
        """u_a - u_b = jump  (periodic boundary condition with a prescribed macroscopic jump)"""
        return cls([dof_a, dof_b], [1.0, -1.0], jump)

    @classmethod
    def tie(cls, dof_slave, dofs_master, weights):
        """u_slave = Σ w_k u_master,k  (multi-point tie / hanging node)"""
        return cls([dof_slave, *dofs_master], [1.0, *(-np.asarray(weights, dtype=float))], 0.0)

    @classmethod
    def average(cls, iDoFs, value=0.0, weights=None):
        """Σ w_k u_k = value  (prescribed weighted average, default w = 1/n)"""
        n = len(iDoFs)
        return cls(iDoFs, np.full(n, 1.0 / n) if weights is None else weights, value)

    @classmethod
    def distance(cls, dofs_a, dofs_b, x_a, x_b, length):
        """|x_a + u_a - x_b - u_b|² = length²  (rigid link between two nodes; dofs_a/dofs_b: the D dofs of each node)"""
        D = len(dofs_a)
        dx = np.asarray(x_a, dtype=float) - np.asarray(x_b, dtype=float)
        S = np.hstack([np.eye(D), -np.eye(D)])     # u_a - u_b = S v
        return cls([*dofs_a, *dofs_b], 2.0 * (S.T @ dx), float(length) ** 2 - float(dx @ dx), 2.0 * (S.T @ S))


def makeϕrKt(eqns, u, λ, device=-1):
    """`makeϕrKt(eqns::Array{ConstEq}, u, λ)` (solvers.jl:24-73) for catalogue constraints, evaluated on the GPU.
    Returns (veqs [nEqs], reqs scipy CSC nDoFs x nEqs, Keqs scipy CSC nDoFs x nDoFs): veqs[ii] = ϕ_ii, reqs[iDoFs,ii] =
    grad, Keqs[iDoFs,iDoFs] += λ[ii] hess, in equation order.  (The reference's `Distributed` chunking, :36-50, only changes
    who evaluates which equation; the sums are the same.)"""
    import scipy.sparse as sp
    nE = len(eqns)
    uf = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(-1, order="F"))
    nD = uf.size
    lam = np.ascontiguousarray(λ, dtype=np.float64).reshape(-1)
    if lam.size != nE:
        raise ValueError("λ must have one entry per equation")
    if nE == 0:
        return np.zeros(0), sp.csc_matrix((nD, 0)), sp.csc_matrix((nD, nD))
    sizes = np.array([len(e.iDoFs) for e in eqns], dtype=np.int64)
    dof_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    dofs = np.concatenate([e.iDoFs for e in eqns]).astype(np.int64)
    c = np.concatenate([e.c for e in eqns])
    c0 = np.array([e.c0 for e in eqns], dtype=np.float64)
    qs = np.array([0 if e.Q is None else e.Q.size for e in eqns], dtype=np.int64)
    q_ptr = np.concatenate([[0], np.cumsum(qs)]).astype(np.int64)
    Q = np.concatenate([np.zeros(0)] + [e.Q.reshape(-1, order="F") for e in eqns if e.Q is not None])
    veqs, grad, hess = np.empty(nE), np.empty(len(dofs)), np.empty(len(Q))
    s = capi.ConstEqSet(nE, dof_ptr.ctypes.data, dofs.ctypes.data, c.ctypes.data, c0.ctypes.data, q_ptr.ctypes.data,
                        Q.ctypes.data if len(Q) else None)
    capi.check(capi.lib().ad4sm_consteq_eval(C.byref(s), nD, uf.ctypes.data, lam.ctypes.data, device, veqs.ctypes.data,
                                             grad.ctypes.data, hess.ctypes.data if len(Q) else None))
    cols = np.repeat(np.arange(nE), sizes)
    # reqs[iDoFs, ii] = grad: an assignment -- a dof listed twice in one equation keeps its LAST value
    last = {}
    for k, (i, j) in enumerate(zip(dofs - 1, cols)):
        last[(i, j)] = k
    keep = np.fromiter(last.values(), dtype=np.int64)
    reqs = sp.csc_matrix((grad[keep], (dofs[keep] - 1, cols[keep])), shape=(nD, nE))
    II, JJ, VV = [], [], []
    for ii, e in enumerate(eqns):
        if e.Q is not None:
            n = len(e.iDoFs)
            II.append(np.tile(e.iDoFs - 1, n))
            JJ.append(np.repeat(e.iDoFs - 1, n))
            VV.append(hess[q_ptr[ii]:q_ptr[ii + 1]])
    Keqs = (sp.csc_matrix((np.concatenate(VV), (np.concatenate(II), np.concatenate(JJ))), shape=(nD, nD)) if II
            else sp.csc_matrix((nD, nD)))
    return veqs, reqs, Keqs


def _splu_solve(K, rhs):
    from scipy.sparse.linalg import splu
    return splu(K.to_scipy()).solve(rhs)


def solvestep(plan, uold, unew, bfreeu, fe=None, dTol=1e-5, dTolu=None, maxiter=11, becho=False, bpredict=True,
              maxupdt=float("nan"), solve=_splu_solve, evaluate=None):
    """`solvestep!(elems, uold, unew, bfreeu; fe, dTol, dTolu, maxiter, becho, bpredict, maxupdt)` (solvers.jl:172-300).

    plan    : an `Elements.Plan` built for `elems` and size(u)  (the reference passes `elems`)
    uold    : converged displacements of the previous step, D x nNodes
    unew    : prescribed values on the constrained dofs (updated IN PLACE on the free dofs, like the reference)
    bfreeu  : bool, shaped like u: True = free dof
    fe      : external forces, length(u) (updated in place at convergence, :281), default zeros
    evaluate: test hook -- a function u -> (ϕ, r, Kt_scipy_csc) replacing the device path AND the device glue with plain
              host slicing (what the reference does); used by the tests to run the identical loop on the oracle.
    Returns (bfailed, normru, iter), like the reference (:329).
    """
    dTolu = dTol if dTolu is None else dTolu
    shape = unew.shape
    bfree = np.asarray(bfreeu, dtype=bool).reshape(-1, order="F")
    ifree = np.flatnonzero(bfree)           # findall(bfreeu[:])      :183
    icnst = np.flatnonzero(~bfree)          # findall(.!bfreeu[:])    :184
    nfree, ncnst = len(ifree), len(icnst)
    if fe is None:
        fe = np.zeros(unew.size)
    uo = uold.reshape(-1, order="F")
    un = unew.reshape(-1, order="F").copy()
    if evaluate is None:
        plan.newton_setup(bfree)

    def assemble(uflat, kind, ducnst=None):
        """-> (res, Kt[ifree,ifree] as scipy CSC or SparseMatrixCSC, norm(res), norm(fi[icnst]), fi or None)"""
        u2 = np.asfortranarray(uflat.reshape(shape, order="F"))
        if evaluate is not None:     # host slicing, as written in the reference
            _, fi, Kt = evaluate(u2)
            Kt = Kt.tocsc()
            if kind == 0:
                res = fi[ifree] - fe[ifree]
                res -= Kt[ifree][:, icnst] @ ducnst
            else:
                res = fe[ifree] - fi[ifree]

            class _W:   # tiny adaptor so that `solve` sees the same interface either way
                def __init__(self, m):
                    self.m = m

                def to_scipy(self):
                    return self.m
            return res, _W(Kt[ifree][:, ifree].tocsc()), float(np.linalg.norm(res)), float(np.linalg.norm(fi[icnst])), fi
        plan.makeϕrKt(u2, fetch=False)
        res, Kff, nres, ncn = plan.newton_reduce(kind, fe=fe, ducnst=ducnst)
        return res, Kff, nres, ncn, None

    bdone = bfailed = False
    it = 0
    normru = float("nan")
    updt = np.zeros(nfree)
    # predictor step (:209-245)
    if bpredict:
        ducnst = un[icnst] - uo[icnst]
        res, Kff, _, _, _ = assemble(uo, 0, ducnst)
        updt[:] = solve(Kff, res)
        un[ifree] = uo[ifree] + updt
    else:
        un[ifree] = uo[ifree]
    # corrector loop (:248-300)
    last_fi = None
    while not bdone and not bfailed:
        oldupdt = updt.copy()  # noqa: F841  (kept like the reference; only used by its line-search comment)
        res, Kff, nres, ncn, last_fi = assemble(un, 1)
        norm0 = ncn / ncnst if ncnst > 0 else 0.0                                   # :263
        normru = nres / nfree / norm0 if norm0 > dTolu else nres / nfree            # :264
        bdone = normru < dTolu                                                      # :265
        if bdone:
            if last_fi is None:   # fe[:] = fi[:]  (:281) -- r of the last evaluation
                last_fi = plan.makeϕr(np.asfortranarray(un.reshape(shape, order="F")))[1]
            fe[:] = last_fi
        elif it < maxiter:
            updt[:] = solve(Kff, res)                                               # :285
            normupdt = np.abs(updt).max()
            if not np.isnan(maxupdt) and normupdt > maxupdt:                        # :287-292
                updt *= maxupdt / normupdt
            un[ifree] += updt                                                       # :293
        else:
            bfailed = True
        if becho:
            print(f"iter {it:2d}: normru = {normru:.3e}")
        it += 1
    unew[...] = un.reshape(shape, order="F")
    return bfailed, normru, it
