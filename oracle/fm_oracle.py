"""ORACLE (test infrastructure, not product code) -- CPU restatement of the reference FM path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module. The shipped package never does.

Parity status: PINNED. ``tests/golden/make_golden.py`` imports the unmodified reference
from ``/root/reference`` in the build container and stores its outputs (losses, parameters,
predictions) as ``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` checks this
restatement against those fixtures (max-abs 1e-13 or better in float64).

What is restated (reference file:line):

* ``sigmoid``            -- ``src/base.py:63-66``  (clip to +-700, then logistic)
* ``ips_logloss``        -- ``src/base.py:37-61``  (eps inside both logs, mean over rows)
* ``fm_predict``         -- ``src/fm.py:114-133``  (O(kn) sum-of-squares identity)
* ``fm_step``            -- ``src/fm.py:80-88,135-187`` as ONE simultaneous full-batch step:
  the residual (``fm.py:80``) and ``V.T @ X.T`` (``fm.py:165``) are taken once, before any
  parameter changes, and column f of V is untouched when iteration f reads it (``fm.py:175``),
  so the per-factor loop equals the closed form below.
  Gradients are sums over the batch, not means (``fm.py:142,153,178-180``); no L2.
* ``legacy_batch``       -- ``sklearn.utils.resample(replace=False, random_state=epoch)`` as
  called at ``src/fm.py:72-79`` / ``src/mf.py:88-95``: ``idx = arange(N);
  RandomState(epoch).shuffle(idx); idx[:B]``; raises ValueError when B > N like sklearn.
* ``fm_fit``             -- ``src/fm.py:55-112`` epoch loop: one minibatch per "epoch",
  post-update batch loss, full val loss every epoch.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix

EPS = 1e-8


def sigmoid(x):
    x = np.clip(x, -700, 700)
    return 1.0 / (1.0 + np.exp(-x))


def ips_logloss(y, p, ps, eps: float = EPS) -> float:
    y = np.asarray(y)
    r = y / ps
    return float(-np.sum(r * np.log(p + eps) + (1 - r) * np.log(1 - p + eps)) / len(y))


def fm_logits(X: csr_matrix, w0, w, V):
    s = X.dot(V)                                        # (R, k)
    q = X.power(2).dot(V ** 2).sum(axis=1)
    return w0 + X.dot(w) + 0.5 * ((s ** 2).sum(axis=1) - np.asarray(q).ravel())


def fm_predict(X: csr_matrix, w0, w, V):
    return sigmoid(fm_logits(X, float(np.asarray(w0).ravel()[0]), w, V))


def fm_grad(X: csr_matrix, y, ps, w0, w, V):
    """Descent direction of one batch (a plain SUM over its rows, so partial batches add up):
    returns (sum_t e_t,  X^T e,  sum_t e_t (x_tj s_t - x_tj^2 v_j))."""
    w0 = float(np.asarray(w0).ravel()[0])
    e = y / ps - fm_predict(X, w0, w, V)                # fm.py:80, pre-update parameters
    S = X.dot(V)                                        # == (V.T @ X.T).T, fm.py:165
    Xe = X.multiply(e[:, None]).tocsr()                 # diag(e) @ X
    a = np.asarray(Xe.sum(axis=0)).ravel()              # X^T e           (fm.py:153)
    c = np.asarray(X.power(2).multiply(e[:, None]).sum(axis=0)).ravel()   # (X∘X)^T e
    G = Xe.T.dot(S) - c[:, None] * V                    # sum_t e_t (x_tj s_t - x_tj^2 v_j)
    return e.sum(), a, G


def fm_step(X: csr_matrix, y, ps, w0, w, V, lr: float):
    """One reference epoch on an already-sampled batch. Returns new (w0, w, V)."""
    g0, a, G = fm_grad(X, y, ps, w0, w, V)
    return np.array([float(np.asarray(w0).ravel()[0]) + lr * g0]), w + lr * a, V + lr * G


def legacy_batch(n_rows: int, batch_size: int, epoch: int) -> np.ndarray:
    if batch_size > n_rows:
        raise ValueError(
            "Cannot sample %d out of arrays with dim %d when replace is False" % (batch_size, n_rows))
    idx = np.arange(n_rows)
    np.random.RandomState(epoch).shuffle(idx)
    return idx[:batch_size]


def fm_init(seed: int, n_features: int, n_factors: int, alpha: float = 2.0):
    """``src/fm.py:31-48``: seeds the GLOBAL legacy RNG, then draws w, V in that order."""
    np.random.seed(seed)
    limit = alpha * np.sqrt(6 / n_features)
    w = np.random.uniform(low=-limit, high=limit, size=n_features)
    limit = alpha * np.sqrt(6 / n_factors)
    V = np.random.uniform(low=-limit, high=limit, size=(n_features, n_factors))
    return np.array([0.0]), w, V


def fm_fit(train, val, n_epochs, batch_size, lr, w0, w, V, sampler=legacy_batch, first_epoch=0):
    X, y, ps = train["features"].tocsr(), np.asarray(train["labels"]), np.asarray(train["pscores"])
    Xv, yv, psv = val["features"].tocsr(), np.asarray(val["labels"]), np.asarray(val["pscores"])
    train_loss, val_loss = [], []
    for epoch in range(first_epoch, first_epoch + n_epochs):
        idx = sampler(X.shape[0], batch_size, epoch)
        bX, by, bps = X[idx], y[idx], ps[idx]
        w0, w, V = fm_step(bX, by, bps, w0, w, V, lr)
        train_loss.append(ips_logloss(by, fm_predict(bX, w0, w, V), bps))
        val_loss.append(ips_logloss(yv, fm_predict(Xv, w0, w, V), psv))
    return (w0, w, V), train_loss, val_loss
