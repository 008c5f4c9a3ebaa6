"""world_size-2 gloo test (CPU) of the data-parallel FM step's host logic.

The device work is replaced by the NumPy oracle behind the same four callables the C-ABI binding
provides, so what is tested is the composition: contiguous batch slices, one gradient all-reduce,
identical apply on every rank, loss partial sums -- and that the result equals the single-process
reference step."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import PKG, ROOT, load_golden, golden_csr


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_epochs, out_dir):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    from oracle import fm_oracle, sampler_oracle
    from rfm_b200.dist import DataParallelFM, DistEnv

    g = load_golden("coat_fm_ips_alpha01")
    X = golden_csr(g, "train")
    y, ps = g["train_labels"], g["train_pscores"]
    Xv, yv, psv = golden_csr(g, "val"), g["val_labels"], g["val_pscores"]
    B, lr, k = int(g["B"]), float(g["lr"]), int(g["k"])
    n = X.shape[1]
    params = {"w0": g["w0_init"].copy(), "w": g["w_init"].copy(), "V": g["V_init"].copy()}
    env = DistEnv("gloo")
    H = 4    # header of the gradient buffer: [sum_e, loss_batch, loss_val, pad], as in csrc/fm.cu (GRAD_W_OFF)
    grad = torch.zeros(H + n + n * k, dtype=torch.float64)
    loss = torch.zeros(2, dtype=torch.float64)
    state = {}

    def local_grad(begin, end, epoch):
        idx = sampler_oracle.legacy_batch(X.shape[0], B, epoch)
        state["idx"] = idx
        mine = idx[begin:end]
        g0, a, G = fm_oracle.fm_grad(X[mine], y[mine], ps[mine], params["w0"], params["w"], params["V"])
        grad[:H] = 0.0                      # the device path zeroes the whole buffer before accumulating
        grad[0] = g0
        grad[H:H + n] = torch.from_numpy(a)
        grad[H + n:] = torch.from_numpy(G.reshape(-1))

    def apply(step_lr):
        gnp = grad.numpy()
        params["w0"] = params["w0"] + step_lr * gnp[0]
        params["w"] = params["w"] + step_lr * gnp[H:H + n]
        params["V"] = params["V"] + step_lr * gnp[H + n:].reshape(n, k)

    def term_sum(Xs, ys, pss):
        p = fm_oracle.fm_predict(Xs, params["w0"], params["w"], params["V"])
        r = ys / pss
        return float(-np.sum(r * np.log(p + 1e-8) + (1 - r) * np.log(1 - p + 1e-8)))

    def local_loss_sums(begin, end, vbegin, vend):
        mine = state["idx"][begin:end]
        loss[0] = term_sum(X[mine], y[mine], ps[mine])
        loss[1] = term_sum(Xv[vbegin:vend], yv[vbegin:vend], psv[vbegin:vend])

    dp = DataParallelFM(env, B, Xv.shape[0], lr, local_grad, grad, apply, local_loss_sums, loss)
    tl, vl = [], []
    for epoch in range(n_epochs):
        prev = dp.step(epoch)                  # the previous epoch's global loss sums ride in this all-reduce
        if prev is not None:
            tl.append(float(prev[0]) / B)
            vl.append(float(prev[1]) / Xv.shape[0])
    out = dp.flush().numpy()
    tl.append(out[0] / B)
    vl.append(out[1] / Xv.shape[0])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), tl=tl, vl=vl, **params)
    env.shutdown()


@pytest.mark.parametrize("world", [2, 3])
def test_data_parallel_step_equals_single_process(tmp_path, world):
    n_epochs = 4
    mp.spawn(_worker, args=(world, _free_port(), n_epochs, str(tmp_path)), nprocs=world, join=True)
    g = load_golden("coat_fm_ips_alpha01")
    ranks = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    for r in ranks[1:]:                       # identical apply: every rank ends with the same bits
        for key in ("w0", "w", "V", "tl", "vl"):
            np.testing.assert_array_equal(r[key], ranks[0][key])
    # and the trajectory is the reference's (golden losses are from the unmodified reference)
    np.testing.assert_allclose(ranks[0]["tl"], g["train_loss"][:n_epochs], rtol=1e-10)
    np.testing.assert_allclose(ranks[0]["vl"], g["val_loss"][:n_epochs], rtol=1e-10)


def test_slice_bounds_cover_the_batch_exactly():
    from rfm_b200.dist import slice_bounds
    for n in (1, 2, 7, 500, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            cuts = [slice_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


class _FakeModel:
    """predict() of a fixed linear model: the gloo test below is about the row split and the gather."""

    def __init__(self, w):
        self.w = w

    def predict(self, X):
        return np.asarray(X @ self.w).ravel()


def _predict_worker(rank, world, port, out_dir):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from rfm_b200.dist import DistEnv, sharded_predict
    rng = np.random.default_rng(0)
    X, w = rng.normal(size=(101, 7)), rng.normal(size=7)
    env = DistEnv("gloo")
    out = sharded_predict(_FakeModel(w), X, env)
    np.save(os.path.join(out_dir, "pred%d.npy" % rank), out)
    env.shutdown()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_predict_equals_single_process(tmp_path, world):
    mp.spawn(_predict_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    X, w = rng.normal(size=(101, 7)), rng.normal(size=7)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / ("pred%d.npy" % r)), X @ w)


def test_fit_data_parallel_host_flow_with_stub_device(monkeypatch):
    """The Python flow of FactorizationMachines._fit_data_parallel (loss history assembly from the one-step-late
    loss sums, the final flush, normalisation, phase timings) with the device pieces stubbed: no GPU, no C calls."""
    import types
    import torch
    from rfm_b200 import dist as rdist, fm as fmmod

    class StubTrainer:
        def __init__(self, *a):
            self.closed = False

        def close(self):
            self.closed = True

    class StubDP:
        loss_tensor = torch.zeros(2, dtype=torch.float64)

        def __init__(self):
            self.steps = []

        def step(self, epoch):                 # returns the PREVIOUS epoch's global [batch, val] loss sums
            self.steps.append(epoch)
            return None if epoch == 0 else torch.tensor([10.0 * epoch, 4.0 * epoch], dtype=torch.float64)

        def flush(self):
            return torch.tensor([10.0 * len(self.steps), 4.0 * len(self.steps)], dtype=torch.float64)

    dp = StubDP()
    monkeypatch.setattr(fmmod, "_FmTrainer", StubTrainer)
    monkeypatch.setattr(rdist, "make_fm_dp", lambda *a, **k: dp)
    m = object.__new__(fmmod.FactorizationMachines)
    m.distributed = types.SimpleNamespace(world=2, rank=1, torch=torch)
    m._context = lambda: types.SimpleNamespace(launch_count=lambda: 7)
    m.batch_size, m.n_epochs, m.sampler, m.lr, m.evaluator, m._dev = 10, 3, "feistel", 0.1, None, None
    m._upload_seconds = 0.25
    m.sync_to_host = lambda: None
    rows = types.SimpleNamespace(h2d_bytes=5, n_rows=4)
    train_loss, val_loss = m._fit_data_parallel(rows, rows, 100)
    assert dp.steps == [0, 1, 2]
    assert train_loss == [1.0, 2.0, 3.0] and val_loss == [1.0, 2.0, 3.0]       # sums / batch_size, sums / n_val
    st = m.last_fit_stats
    assert st["h2d_bytes_rows"] == 10 and st["upload_seconds"] == 0.25 and st["gpu_launches"] == 0
    assert set(st["phase_seconds"]) == {"upload", "trainer_create", "dp_connect", "enqueue_epochs",
                                        "drain_and_read_losses", "trainer_destroy", "download_params"}


def test_csr_ranges_tile_the_device_arrays():
    """Host arithmetic of the sharded upload: the ranks' byte ranges partition col / val / targets exactly."""
    from rfm_b200.dist import csr_ranges
    rng = np.random.default_rng(5)
    for n_rows, world, es in ((0, 2, 8), (1, 3, 8), (10, 4, 4), (1000, 8, 8), (7, 8, 4)):
        lens = rng.integers(0, 6, size=n_rows)
        indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(indptr[-1])
        ranges = csr_ranges(indptr, n_rows, world, es)
        assert len(ranges) == world
        for which, (unit, total) in enumerate(((4, nnz), (es, nnz), (es, n_rows))):
            cursor = 0
            for r in range(world):
                off, nbytes = ranges[r][which]
                assert off == cursor and nbytes >= 0 and nbytes % unit == 0
                cursor += nbytes
            assert cursor == total * unit
