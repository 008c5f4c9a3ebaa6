"""Import-path compatibility with the reference: ``from utils.evaluate import TestEvaluator``.

When this directory precedes the reference checkout on ``sys.path``, the hot-path modules
(``utils.optimizer``, ``utils.metrics``, ``utils.evaluate``) resolve here, and everything else of the
reference's ``utils`` package (``utils.dataloader``, ``utils.plot``, ...) still resolves to the
reference because its directory is appended to this package's search path.
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
for _p in list(_sys.path):
    _cand = _os.path.join(_os.path.abspath(_p or "."), "utils")
    if _os.path.isdir(_cand) and _cand != _here and _cand not in __path__:
        __path__.append(_cand)
