from rfm_b200.mf import LogisticMatrixFactorization  # noqa: F401
