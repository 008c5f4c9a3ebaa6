from rfm_b200.optimizer import BaseOptimizer, SGD  # noqa: F401
