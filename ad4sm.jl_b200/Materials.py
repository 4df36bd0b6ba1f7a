"""Host-side mirror of `AD4SM.Materials` (material *descriptors* only -- the energies are evaluated
on the GPU by csrc/materials.cuh).

Mirrors /root/reference/src/elasticmaterials.jl:25-120 and /root/reference/src/phasefieldmaterials.jl:5-16:
same type names, positional arguments, keyword names and defaults.
"""
from __future__ import annotations

from dataclasses import dataclass

__all__ = ["Material", "Hooke", "Hooke2D", "MooneyRivlin", "NeoHooke", "Ogden", "PhaseField"]


class Material:
    pass


@dataclass(frozen=True)
class Hooke(Material):
    """`Hooke(E, ν, ρ=1; small=false)` -- 3-D St.Venant-Kirchhoff / linear (elasticmaterials.jl:25-31)."""
    E: float
    nu: float
    rho: float = 1.0
    small: bool = False

    @property
    def ν(self):
        return self.nu


@dataclass(frozen=True)
class Hooke2D(Material):
    """`Hooke2D(E, ν, ρ=1; small=true, plane_stress=true)` (elasticmaterials.jl:65-73)."""
    E: float
    nu: float
    rho: float = 1.0
    small: bool = True
    plane_stress: bool = True

    @property
    def ν(self):
        return self.nu


@dataclass(frozen=True)
class MooneyRivlin(Material):
    """`MooneyRivlin(C1, C2[, K])`; K<0 (default -1) selects the incompressible form (:87-94)."""
    C1: float
    C2: float
    K: float = -1.0
    small: bool = False


@dataclass(frozen=True)
class NeoHooke(Material):
    """`NeoHooke(C1, K, ρ=1)` (elasticmaterials.jl:100-106)."""
    C1: float
    K: float
    rho: float = 1.0
    small: bool = False


@dataclass(frozen=True)
class Ogden(Material):
    """`Ogden(α, μ[, K])`.  The reference constructor passes 4 arguments to a 3-field struct and
    throws (elasticmaterials.jl:114-120); this mirror implements the evident intent."""
    alpha: float
    mu: float
    K: float = -1.0

    @property
    def α(self):
        return self.alpha

    @property
    def μ(self):
        return self.mu


@dataclass(frozen=True)
class PhaseField(Material):
    """`PhaseField(l0, Gc, mat, n)` -> `PhaseField{M,:ATn}` (phasefieldmaterials.jl:5-16).
    `AT="DHn"` selects the deviatoric/hydrostatic split (:56-72), which the reference can only
    build through the inner constructor `PhaseField{M,:DHn}(l0,Gc,mat,n)`."""
    l0: float
    Gc: float
    mat: Material
    n: float
    AT: str = "ATn"

    def __post_init__(self):
        if self.AT not in ("ATn", "DHn"):
            raise ValueError("AT must be 'ATn' or 'DHn'")
