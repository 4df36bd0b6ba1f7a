"""Import shim: the product package lives in the directory `ad4sm.jl_b200/` (the name the build
contract asks for), which is not a legal Python identifier.  `import ad4sm_b200` resolves the
sub-modules from there."""
import os as _os

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
__path__.append(_os.path.join(_root, "ad4sm.jl_b200"))

from . import Materials, Elements, mesh  # noqa: E402,F401
from .Elements import makeϕrKt, makeϕrKt_d, makeϕr, make_phi_r_Kt, make_phi_r_Kt_d  # noqa: E402,F401

__version__ = "0.1.0"
