"""Import-path compatibility with the reference: ``from src.fm import FactorizationMachines``.

When this directory precedes the reference checkout on ``sys.path``, the hot-path modules
(``src.base``, ``src.fm``, ``src.mf``) resolve here, and everything else of the
reference's ``src`` package (none today) still resolves to the
reference because its directory is appended to this package's search path.
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
for _p in list(_sys.path):
    _cand = _os.path.join(_os.path.abspath(_p or "."), "src")
    if _os.path.isdir(_cand) and _cand != _here and _cand not in __path__:
        __path__.append(_cand)
