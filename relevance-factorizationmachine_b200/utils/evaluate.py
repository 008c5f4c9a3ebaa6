from rfm_b200.evaluate import TestEvaluator, ValEvaluator, _BaseEvaluator  # noqa: F401
