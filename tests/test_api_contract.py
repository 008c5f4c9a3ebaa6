"""Drop-in boundary: rfm_b200's classes against the reference's API surface recorded by
tests/golden/make_api_contract.py from the live, unmodified reference (dataclass field order and defaults,
method parameter names, the metric registry, error types and messages). CPU only: nothing is computed."""
import dataclasses
import inspect
import json
import os

import numpy as np
import pytest

from conftest import ROOT

CONTRACT = json.load(open(os.path.join(ROOT, "tests", "golden", "api_contract.json")))


def _classes():
    from rfm_b200.base import PointwiseBaseRecommender
    from rfm_b200.evaluate import TestEvaluator, ValEvaluator
    from rfm_b200.fm import FactorizationMachines
    from rfm_b200.mf import LogisticMatrixFactorization
    from rfm_b200.optimizer import SGD, BaseOptimizer
    return {c.__name__: c for c in (PointwiseBaseRecommender, FactorizationMachines, LogisticMatrixFactorization,
                                    TestEvaluator, ValEvaluator, SGD, BaseOptimizer)}


@pytest.mark.parametrize("name", sorted(CONTRACT["fields"]))
def test_dataclass_fields_keep_the_reference_order_and_defaults(name):
    """Positional construction as in main_coat.py:95-118 must keep working: the reference's fields come first,
    in its order, with its defaults; anything this build adds is a trailing field with a default."""
    ours = dataclasses.fields(_classes()[name])
    ref = CONTRACT["fields"][name]
    assert [f.name for f in ours[: len(ref)]] == [f["name"] for f in ref]
    for f, r in zip(ours, ref):
        has_default = f.default is not dataclasses.MISSING
        assert has_default == r["has_default"], f.name
        if has_default:
            assert repr(f.default) == r["default"], f.name
    for f in ours[len(ref):]:
        assert f.default is not dataclasses.MISSING or f.default_factory is not dataclasses.MISSING, \
            "extra field %s.%s needs a default" % (name, f.name)


@pytest.mark.parametrize("qualname", sorted(CONTRACT["methods"]))
def test_method_parameter_names(qualname):
    """Callers use keywords (predict(X=...), evaluate(y_scores=..., estimator=...): src/fm.py:105-109,
    main_coat.py:121-124), so the names are part of the contract."""
    cls, meth = qualname.split(".")
    ours = [p for p in inspect.signature(getattr(_classes()[cls], meth)).parameters if p != "self"]
    assert ours == CONTRACT["methods"][qualname]


def test_metric_registry_and_function_signatures():
    from rfm_b200.metrics import metric_candidates
    assert sorted(metric_candidates) == sorted(CONTRACT["metric_candidates"])
    for name, ref_params in CONTRACT["metric_candidates"].items():
        assert list(inspect.signature(metric_candidates[name]).parameters) == ref_params, name


def test_evaluator_errors_match_the_reference():
    from rfm_b200.evaluate import TestEvaluator, ValEvaluator
    frame = {"user": np.array([0, 0, 1]), "item": np.array([0, 1, 0]), "label": np.array([1, 0, 1]),
             "pscore": np.full(3, 0.5), "ones_pscore": np.ones(3)}
    want = CONTRACT["errors"]["TestEvaluator_unknown_metric"]
    with pytest.raises(ValueError) as e:
        TestEvaluator(frame, {}, (1, 3), {"DCG", "nope"}, 2)
    assert type(e.value).__name__ == want["type"]
    # the message embeds dict_keys([...]) of the registry: same text when the registry has the same order
    assert str(e.value).endswith("metric_name: 'nope'") and str(e.value).startswith("metric_name must be in")
    assert str(e.value) == want["message"]
    want = CONTRACT["errors"]["ValEvaluator_non_dcg_metric"]
    with pytest.raises(ValueError) as e:
        ValEvaluator(frame, {}, 3, "Recall")
    assert str(e.value) == want["message"]


def test_reference_import_paths_resolve_to_this_build():
    """with the package directory first on PYTHONPATH the reference's own import
    statements (src/fm.py:12-13, main_coat.py) bind to rfm_b200. Child process: clean sys.modules / sys.path."""
    import subprocess
    import sys
    from conftest import PKG
    code = ("import src.fm, src.mf, src.base, utils.optimizer, utils.metrics, utils.evaluate\n"
            "mods = [src.fm.FactorizationMachines, src.mf.LogisticMatrixFactorization, utils.optimizer.SGD,\n"
            "        utils.evaluate.TestEvaluator, utils.evaluate.ValEvaluator, src.base.PointwiseBaseRecommender]\n"
            "assert all(m.__module__.startswith('rfm_b200.') for m in mods), [m.__module__ for m in mods]\n"
            "assert utils.metrics.metric_candidates is __import__('rfm_b200.metrics').metrics.metric_candidates\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=PKG)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stderr[-2000:]
