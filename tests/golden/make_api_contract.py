"""Records the reference's API surface at the drop-in boundary as a small JSON fixture.

Run in the build container, where the unmodified reference is importable:
    python tests/golden/make_api_contract.py            # writes tests/golden/api_contract.json
tests/test_api_contract.py then holds rfm_b200's classes to it on any machine (the fixture travels,
/root/reference does not). Nothing here is copied from the reference: names, defaults, parameter
lists and error strings are read off the live objects with ``dataclasses`` / ``inspect``.
"""
import dataclasses
import inspect
import json
import os
import sys

REF = os.environ.get("RFM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np                                             # noqa: E402
import pandas as pd                                            # noqa: E402
from src.fm import FactorizationMachines                       # noqa: E402
from src.mf import LogisticMatrixFactorization                 # noqa: E402
from src.base import PointwiseBaseRecommender                  # noqa: E402
from utils.evaluate import TestEvaluator, ValEvaluator         # noqa: E402
from utils.metrics import metric_candidates                    # noqa: E402
from utils.optimizer import SGD, BaseOptimizer                 # noqa: E402


def fields(cls):
    out = []
    for f in dataclasses.fields(cls):
        has_default = f.default is not dataclasses.MISSING
        out.append({"name": f.name, "has_default": has_default, "default": repr(f.default) if has_default else None})
    return out


def params(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def error_of(fn):
    try:
        fn()
    except Exception as e:                                     # noqa: BLE001
        return {"type": type(e).__name__, "message": str(e)}
    return None


frame = pd.DataFrame({"user": [0, 0, 1], "item": [0, 1, 0], "label": [1, 0, 1], "pscore": [0.5, 0.5, 0.5],
                      "ones_pscore": [1.0, 1.0, 1.0]})
train = {"features": None, "labels": np.zeros(3), "pscores": np.ones(3)}

contract = {
    "fields": {c.__name__: fields(c) for c in (PointwiseBaseRecommender, FactorizationMachines,
                                               LogisticMatrixFactorization, TestEvaluator, ValEvaluator, SGD,
                                               BaseOptimizer)},
    "methods": {
        "FactorizationMachines.fit": params(FactorizationMachines.fit),
        "FactorizationMachines.predict": params(FactorizationMachines.predict),
        "LogisticMatrixFactorization.fit": params(LogisticMatrixFactorization.fit),
        "LogisticMatrixFactorization.predict": params(LogisticMatrixFactorization.predict),
        "TestEvaluator.evaluate": params(TestEvaluator.evaluate),
        "ValEvaluator.evaluate": params(ValEvaluator.evaluate),
        "SGD.update": params(SGD.update),
        "SGD.__call__": params(SGD.__call__),
    },
    "metric_candidates": {name: params(fn) for name, fn in metric_candidates.items()},
    "errors": {
        "TestEvaluator_unknown_metric": error_of(lambda: TestEvaluator(frame, {}, (1, 3), {"DCG", "nope"}, 2)),
        "ValEvaluator_non_dcg_metric": error_of(lambda: ValEvaluator(frame, {}, 3, "Recall")),
    },
    "reference_versions": {"numpy": np.__version__, "pandas": pd.__version__},
}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_contract.json")
with open(out, "w") as f:
    json.dump(contract, f, indent=1, sort_keys=True)
print("wrote", out)
