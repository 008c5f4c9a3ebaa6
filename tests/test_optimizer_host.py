"""Host-side checks of the optimizer holders (reference contract utils/optimizer.py:10-64 + the Adam
extension) against the NumPy specification in oracle/optimizer_oracle.py. No GPU needed."""
import numpy as np

from oracle import optimizer_oracle
from rfm_b200.optimizer import SGD, Adam, BaseOptimizer


def test_sgd_holder_contract():
    p = SGD(params=np.arange(6.0).reshape(3, 2), lr=0.5)
    assert isinstance(p, BaseOptimizer)
    assert p() is p.params and p(1).tolist() == [2.0, 3.0]
    p.update(np.ones((3, 2)), None)
    np.testing.assert_array_equal(p(), np.arange(6.0).reshape(3, 2) - 0.5)
    p.update(2.0, (np.array([0, 2]), 1))                      # (rows, col) tuple index, src/fm.py:183-187
    np.testing.assert_array_equal(p()[:, 1], [0.5 - 1.0, 2.5, 4.5 - 1.0])


def test_adam_holder_matches_the_specification():
    rng = np.random.default_rng(0)
    theta = rng.normal(size=(5, 3))
    h = Adam(params=theta.copy(), lr=0.01, beta1=0.8, beta2=0.95, eps=1e-6, l2=0.1)
    m, v = np.zeros_like(theta), np.zeros_like(theta)
    ref = theta.copy()
    for step in range(1, 6):
        g = rng.normal(size=theta.shape)
        h.update(g, None)
        ref, m, v = optimizer_oracle.adam_step(ref, -g, m, v, step, 0.01, 0.8, 0.95, 1e-6, 0.1)   # d = -g
        np.testing.assert_allclose(h(), ref, rtol=1e-14, atol=1e-16)
    assert h.t == 5 and h.version == 5


def test_adam_holder_indexed_update_touches_only_the_index():
    h = Adam(params=np.ones((4, 2)), lr=0.1)
    h.update(np.array([1.0, -1.0]), 2)
    assert (h()[[0, 1, 3]] == 1.0).all()
    np.testing.assert_allclose(h()[2], [1.0 - 0.1 / (1 + 1e-8), 1.0 + 0.1 / (1 + 1e-8)], rtol=1e-12)


def test_adam_first_step_is_sign_of_gradient():
    """Known property of bias-corrected Adam: the first step is lr * g / (|g| + eps)."""
    theta, d = np.array([1.0, -2.0, 0.5]), np.array([0.3, -4.0, 1e-3])
    new, m, v = optimizer_oracle.adam_step(theta, d, 0 * theta, 0 * theta, 1, 0.01, 0.9, 0.999, 1e-8, 0.0)
    np.testing.assert_allclose(new, theta + 0.01 * np.sign(d), rtol=1e-5)
