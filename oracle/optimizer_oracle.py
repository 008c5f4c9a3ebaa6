"""TEST INFRASTRUCTURE ONLY -- NumPy specification of the optimizers the reference does not have.

The reference's ``utils/optimizer.py`` ends with SGD (line 64); BASELINE.json's north_star also names
"SGD or Adam with L2". There is nothing in the reference to be in parity with, so this file IS the
specification ("needs its own NumPy spec"), and the CUDA path
(``rfm_fm_train_epoch_opt``) and the host holder (``rfm_b200.optimizer.Adam``) are tested against it.
Parity status: unpinned by construction (no reference implementation exists).

The batch gradient is the reference's own closed form: ``fm_oracle.fm_grad`` returns the ASCENT
direction d (``theta += lr d`` is the step of ``src/fm.py:135-187``), so the loss gradient is ``-d``.
"""
import numpy as np

from . import fm_oracle


def sgd_l2_step(theta, d, lr, l2):
    """theta -= lr (g + l2 theta), g = -d."""
    return theta - lr * (l2 * theta - d)


def adam_step(theta, d, m, v, step, lr, beta1, beta2, eps, l2):
    """One bias-corrected Adam step (Kingma & Ba 2015, Algorithm 1) with coupled L2; returns (theta, m, v)."""
    g = l2 * theta - d
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * (g * g)
    c1 = 1.0 / (1.0 - beta1 ** step)
    c2 = 1.0 / (1.0 - beta2 ** step)
    return theta - lr * (m * c1) / (np.sqrt(v * c2) + eps), m, v


def fm_fit_opt(train, val, n_epochs, batch_size, lr, w0, w, V, kind="adam", l2=0.0, beta1=0.9, beta2=0.999,
               eps=1e-8, sampler=None):
    """The reference epoch loop (src/fm.py:69-102: one minibatch per epoch, post-update batch loss, val loss)
    with the optimizer step replaced. Returns ((w0, w, V), train_loss, val_loss)."""
    sampler = sampler or fm_oracle.legacy_batch
    X, y, ps = train["features"], np.asarray(train["labels"], dtype=np.float64), np.asarray(train["pscores"])
    Xv, yv, psv = val["features"], np.asarray(val["labels"], dtype=np.float64), np.asarray(val["pscores"])
    w0, w, V = float(np.asarray(w0).reshape(-1)[0]), np.array(w, dtype=np.float64), np.array(V, dtype=np.float64)
    state = [[np.zeros_like(np.asarray(p, dtype=np.float64)) for p in (w0, w, V)] for _ in range(2)]
    tl, vl = [], []
    for epoch in range(n_epochs):
        idx = sampler(X.shape[0], batch_size, epoch)
        Xb, yb, pb = X[idx], y[idx], ps[idx]
        d0, dw, dV = fm_oracle.fm_grad(Xb, yb, pb, w0, w, V)
        if kind == "adam":
            w0, state[0][0], state[1][0] = adam_step(w0, d0, state[0][0], state[1][0], epoch + 1, lr, beta1, beta2, eps, l2)
            w, state[0][1], state[1][1] = adam_step(w, dw, state[0][1], state[1][1], epoch + 1, lr, beta1, beta2, eps, l2)
            V, state[0][2], state[1][2] = adam_step(V, dV, state[0][2], state[1][2], epoch + 1, lr, beta1, beta2, eps, l2)
            w0 = float(w0)
        else:
            w0, w, V = float(sgd_l2_step(w0, d0, lr, l2)), sgd_l2_step(w, dw, lr, l2), sgd_l2_step(V, dV, lr, l2)
        tl.append(fm_oracle.ips_logloss(yb, fm_oracle.fm_predict(Xb, w0, w, V), pb))
        vl.append(fm_oracle.ips_logloss(yv, fm_oracle.fm_predict(Xv, w0, w, V), psv))
    return (w0, w, V), tl, vl
