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


def test_roofline_numerator_is_the_survey_formula():
    """bench.py's roofline.step numerator is SURVEY.md section 8(d)'s bytes per train interaction:
    [4 + 4 + m(4 + s) + 4 + s] + 2 m (k+1) s + 2 (U_b/B)(k+1) s. Checked on a matrix with exactly m = 16 non-zeros
    per row against the survey's own worked example (s = 4, k = 64, U_b/B = 0.275 -> about 8.6 KB) and against the
    per-kernel split DESIGN.md section 4.1 tabulates (the three row/column passes must cover the step's gathers)."""
    import bench
    from scipy.sparse import csr_matrix
    rng = np.random.default_rng(0)
    n_rows, n_cols, m, B = 4000, 1100, 16, 4000
    cols = np.stack([rng.choice(n_cols, size=m, replace=False) for _ in range(n_rows)])
    X = csr_matrix((np.ones(n_rows * m), cols.ravel(), np.arange(0, n_rows * m + 1, m)), shape=(n_rows, n_cols))
    batch = np.arange(B)
    for s, k in ((4, 64), (8, 64), (8, 128)):
        mean_nnz, touched, step, per_kernel = bench.algorithmic_bytes(X, batch, k, s)
        assert mean_nnz == m and touched == np.unique(cols).size == n_cols      # every column is hit: U_b/B = 0.275
        want = (4 + 4 + m * (4 + s) + 4 + s) + 2 * m * (k + 1) * s + 2 * (touched / B) * (k + 1) * s
        assert abs(step - want) < 1e-9
        # the gathers of the step (2 x m(k+1)s) are split over the train and loss row passes, the touched-row
        # read/write sits in the column pass; intermediates (S, E, sort triples) come on top, never below
        assert per_kernel["fm_rows_train"] + per_kernel["fm_cols"] + per_kernel["fm_rows_loss"] >= step
    _, _, step, _ = bench.algorithmic_bytes(X, batch, 64, 4)
    assert 8500 < step < 8700                                                   # SURVEY: "about 8.6 KB/interaction"


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the product arm): one JSON line with the
    product arm's metric / unit / config keys, impl = reference, a cpu_baseline describing this run and an e2e
    object that repeats the value with zero copy bytes. Small --rows / --batch so that it runs in seconds."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--rows", "60000", "--batch", "2048"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_interactions_per_sec"
    assert line["unit"] == "interactions/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and 1 <= cb["cores"] <= os.cpu_count()
    assert "B=2048" in cb["sample"]
    cfg = line["config"]
    assert "workload" in cfg and cfg["train_interactions"] == 60000 and cfg["batch_per_gpu"] == 2048
    assert "model" not in cfg


def test_two_level_bytes_formula_and_step_argument():
    """bench.py's own-bytes model of the two-level step at the KuaiRec shape (row record + 2 passes x 3 gathered rows +
    the entity pass spread over the batch), and the drop-in's `step` argument."""
    import bench
    import pytest
    got = bench.two_level_bytes(65536, 64, 8, 1, 138000, 18030)
    want = (4 + 4 + 4 + 8 + 4 + 8) + 2 * 3 * 65 * 8 + (2 * 138000 + 2 * 18030) * 65 * 8 / 65536
    assert abs(got - want) < 1e-9 and 5000 < got < 6000
    assert got < bench.algorithmic_bytes_of(16.479, 17187, 65536, 64, 8)[0] / 3
    from rfm_b200.fm import FactorizationMachines
    with pytest.raises(ValueError, match="step must be"):
        FactorizationMachines("IPS", 1, 4, 1e-3, 10, 1, 20, step="fast")
    assert FactorizationMachines("IPS", 1, 4, 1e-3, 10, 1, 20).step == "auto"
