"""ORACLE (test infrastructure only): literal restatement of the reference's constraint-equation assembly
`makeϕrKt(eqns::Array{ConstEq}, u, λ, chunk)` (/root/reference/src/solvers.jl:54-73) with the oracle's D2 duals:

    for ii in chunk
        ϕ = eqn.func(eqn.D(u[iDoFs]));  veqs[ii] = val(ϕ);  reqs[iDoFs,ii] = grad(ϕ);  Keqs[iDoFs,iDoFs] += λ[ii] hess(ϕ)

`func` is an arbitrary closure here, exactly as in the reference; the tests write the closures a user would write for
periodic conditions, ties, averages and rigid links and compare the product's catalogue evaluation against them.
Dense arrays stand in for the reference's sparse ones (same indexed-assignment semantics: `reqs[iDoFs,ii] = g` assigns,
`Keqs[iDoFs,iDoFs] += λ H` reads, adds and assigns -- a dof listed twice keeps the last value written)."""
import numpy as np

from . import adiff


class ConstEq:  # solvers.jl:17-22
    def __init__(self, func, iDoFs):
        self.func = func
        self.iDoFs = np.asarray(iDoFs, dtype=np.int64)


def make_phi_r_Kt(eqns, u, lam):
    uf = np.asarray(u, dtype=np.float64).reshape(-1, order="F")
    nE, nD = len(eqns), uf.size
    veqs = np.zeros(nE)
    reqs = np.zeros((nD, nE))
    Keqs = np.zeros((nD, nD))
    for ii, eqn in enumerate(eqns):
        idx = eqn.iDoFs - 1
        phi = eqn.func(adiff.seed_D2(uf[idx]))
        veqs[ii] = adiff.val(phi)
        reqs[idx, ii] = adiff.grad(phi)
        Keqs[np.ix_(idx, idx)] = Keqs[np.ix_(idx, idx)] + lam[ii] * adiff.hess(phi)
    return veqs, reqs, Keqs
