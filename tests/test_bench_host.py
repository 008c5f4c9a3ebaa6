"""The multi-process CPU arm of bench.py computes the same epochs as the scalar oracle port (no GPU)."""
import numpy as np

from oracle import fm_oracle


def test_parallel_cpu_port_equals_scalar_port(monkeypatch):
    import bench
    from rfm_b200.synth import make_coat_shaped
    log = make_coat_shaped(seed=5, n_users=60, n_items=80, n_rated=12, n_test=4)
    monkeypatch.setattr(bench, "K_FACTORS", 8)
    monkeypatch.setattr(bench, "LR", 1e-3)
    B, steps, warmup = 200, 3, 1
    value, done, dt, workers = bench.cpu_port_run_parallel(log, B, steps, warmup, budget_s=1e9, workers=3)
    assert done == steps and workers == 3 and value > 0
    tl, vl, w0, w, V = bench._PAR["last"]
    i0, iw, iV = fm_oracle.fm_init(12345, log.n_features, 8)
    (rw0, rw, rV), rtl, rvl = fm_oracle.fm_fit(log.fm_train, log.fm_val, warmup + steps, B, 1e-3, i0, iw, iV)
    np.testing.assert_allclose(tl, rtl[-1], rtol=1e-10)
    np.testing.assert_allclose(vl, rvl[-1], rtol=1e-10)
    np.testing.assert_allclose(V, rV, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(w, rw, rtol=1e-10, atol=1e-14)
