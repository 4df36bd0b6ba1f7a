"""Host-side mirror of `Solvers.solvestep!` (/root/reference/src/solvers.jl:172-300), nEqs == 0 branch, on top of the
device-side Newton glue (ad4sm_newton_*).

What runs where
  * `makeϕrKt(elems, u)` (solvers.jl:213-216, 257-260)                              -> the plan, on the GPU;
  * `fi[ifreeu]-fe[ifreeu]`, `Kt[ifreeu,icnstu]*Δucnst`, `Kt[ifreeu,ifreeu]`,
    `norm(fi[icnstu])`, `norm(res)` (:220-221, 262-265, 285)                         -> ad4sm_newton_reduce, on the GPU;
  * `lu(Kt[ifreeu,ifreeu]) \\ res` (:223, 285)                                        -> OUT OF SCOPE of the product
    (north_star): it stays on a host sparse direct solver.  In Julia that is UMFPACK's `lu`; here, where the host
    language is Python, it is `scipy.sparse.linalg.splu` -- passed in as `solve`, so any other solver can be used.

Same argument names, defaults and control flow as the reference; constraint equations (`eqns`, the nEqs != 0 branch)
are closures evaluated by the reference itself and are not covered.
"""
from __future__ import annotations

import numpy as np


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
