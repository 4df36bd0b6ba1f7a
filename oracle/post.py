"""ORACLE (test infrastructure only): literal restatement of the reference's post-processing functions
(/root/reference/src/elements.toolkit.jl:8-132), first-order duals over F for the stresses.  `u` is the element's
D x N block u[:, elem.nodes], as the reference's array methods pass it.

parity status: PARITY UNPINNED BY REFERENCE TESTS (see oracle/adiff.py)."""
from __future__ import annotations

import numpy as np

from . import adiff
from . import materials as M
from .elements import _det2, _det3


def grad_u_gp(elem, u, ii):  # get∇u(elem, u, ii), toolkit:11-20
    D = elem.D
    g = np.zeros((D, D))
    for jj in range(D):
        for kk in range(D):
            g[jj, kk] = float(np.dot(elem.gradN[kk][ii], u[jj, :]))
    return g


def F_gp(elem, u, ii):  # getF(elem, u, ii), toolkit:8
    return grad_u_gp(elem, u, ii) + np.eye(elem.D)


def grad_u(elem, u):  # get∇u(elem, u), toolkit:21-27
    g = np.zeros((elem.D, elem.D))
    for ii in range(elem.P):
        g = g + elem.wgt[ii] * grad_u_gp(elem, u, ii)
    return g / elem.V


def getF(elem, u):  # toolkit:9
    return grad_u(elem, u) + np.eye(elem.D)


def _det(F):
    return _det2(F) if F.shape[0] == 2 else _det3(F)


def detJ(elem, u):  # toolkit:38-46
    J = 0.0
    for ii in range(elem.P):
        J += elem.wgt[ii] * _det(F_gp(elem, u, ii))
    return J / elem.V


def getV(elem, u):  # toolkit:52-58
    total = 0.0
    for ii in range(elem.P):
        total += elem.wgt[ii] * _det(F_gp(elem, u, ii))
    return total


def _P_gp(elem, u, ii):
    F = F_gp(elem, u, ii)
    dphi = M.getphi(adiff.seed_D1(F), elem.mat)          # getϕ(adiff.D1(F), elem.mat), toolkit:81
    return F, adiff.grad(dphi).reshape(F.shape, order="F")  # reshape(adiff.grad(δϕ), size(F)): column-major


def sigma_gp(elem, u, ii):  # getσ(elem, u, ii), toolkit:77-87
    F, P = _P_gp(elem, u, ii)
    return (1.0 / _det(F) * P) @ F.T


def getsigma(elem, u):  # getσ(elem, u), toolkit:68-75
    s = np.zeros((elem.D, elem.D))
    for ii in range(elem.P):
        s = s + elem.wgt[ii] * sigma_gp(elem, u, ii)
    return s / elem.V


def getP(elem, u):  # getP(elem, u), toolkit:117-132 (returns the average of P1K')
    p = np.zeros((elem.D, elem.D))
    for ii in range(elem.P):
        p = p + elem.wgt[ii] * _P_gp(elem, u, ii)[1].T
    return p / elem.V
