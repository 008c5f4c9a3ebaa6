"""CPU-side checks (no GPU needed): the C-ABI library loads, exports every declared symbol,
its host-side samplers match NumPy / the oracle bit for bit, and it refuses to run without a
device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import sampler_oracle
from rfm_b200 import _capi


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rfm_b200.h")).read()
    declared = set(re.findall(r"\b(rfm_[a-z0-9_]+)\s*\(", header))
    declared -= {"rfm_last_error"} - set(_capi.DECLARED_SYMBOLS)
    handle = ctypes.CDLL(_capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, "declared in include/rfm_b200.h but not exported: %s" % missing
    assert set(_capi.DECLARED_SYMBOLS) == declared, set(_capi.DECLARED_SYMBOLS) ^ declared
    assert _capi.lib().rfm_abi_version() == _capi.ABI_VERSION


def test_legacy_sampler_matches_numpy_known_answers():
    g = load_golden("legacy_sampler")
    for key in g.files:
        N, B, ep = (int(s[1:]) for s in key.split("_"))
        np.testing.assert_array_equal(_capi.legacy_batch(N, B, ep), g[key], err_msg=key)


@pytest.mark.parametrize("path", ["array", "prefix"])
@pytest.mark.parametrize("N", [1, 2, 3, 5, 16, 17, 255, 256, 257, 1000, 4097, 70000])
def test_legacy_sampler_matches_numpy_live(N, path, monkeypatch):
    # the library picks the path by size (prefix = backward tracking of the batch positions, for >= 8 M rows and
    # small batches); both must give NumPy's answer everywhere, so each is forced in turn
    monkeypatch.setenv("RFM_LEGACY_SAMPLER_PATH", path)
    for epoch in (0, 1, 2, 499, 123456789):
        ref = np.arange(N)
        np.random.RandomState(epoch).shuffle(ref)
        np.testing.assert_array_equal(_capi.legacy_batch(N, N, epoch), ref)
        # the extremes and a spread of batch sizes, with and without a caller-provided scratch
        scratch = np.empty(N, dtype=np.int32)
        for B in sorted({0, 1, 2, max(1, N // 3), N // 16, N // 16 + 1, N // 50, N - 1}):
            if 0 <= B <= N:
                np.testing.assert_array_equal(_capi.legacy_batch(N, B, epoch), ref[:B], err_msg="B=%d" % B)
                np.testing.assert_array_equal(_capi.legacy_batch(N, B, epoch, scratch), ref[:B], err_msg="B=%d" % B)


@pytest.mark.parametrize("path", ["array", "prefix", ""])
def test_legacy_sampler_large_matches_sklearn_resample(path, monkeypatch):
    """The call the reference makes (src/fm.py:72-79), at a size where the rejection bound crosses several
    powers of two and the tracking path sees tens of thousands of hits; "" = the library's own choice."""
    from sklearn.utils import resample
    if path:
        monkeypatch.setenv("RFM_LEGACY_SAMPLER_PATH", path)
    else:
        monkeypatch.delenv("RFM_LEGACY_SAMPLER_PATH", raising=False)
    N = 1_000_003
    ids = np.arange(N)
    for epoch, B in ((0, 500), (7, 2000), (31, 62500), (499, 65536)):
        want = resample(ids, replace=False, n_samples=B, random_state=epoch)
        np.testing.assert_array_equal(_capi.legacy_batch(N, B, epoch), want)


def test_legacy_sampler_bench_shape_takes_the_prefix_path_and_matches_numpy(monkeypatch):
    """12 M rows, batch 65,536 (the bench shape): the size rule selects backward tracking; one epoch against NumPy."""
    monkeypatch.delenv("RFM_LEGACY_SAMPLER_PATH", raising=False)
    N, B = 12_000_000, 65536
    ref = np.arange(N)
    np.random.RandomState(5).shuffle(ref)
    np.testing.assert_array_equal(_capi.legacy_batch(N, B, 5), ref[:B])


def test_legacy_sampler_rejects_oversized_batch_like_sklearn():
    with pytest.raises(ValueError, match="Cannot sample 11 out of arrays with dim 10"):
        _capi.legacy_batch(10, 11, 0)


@pytest.mark.parametrize("N", [1, 2, 3, 17, 256, 257, 3660, 65536, 100003, 12_000_000])
def test_feistel_host_matches_oracle(N):
    B = min(N, 4096)
    for epoch, seed in ((0, 0), (1, 12345), (77, 2**32 - 1)):
        np.testing.assert_array_equal(_capi.feistel_batch(N, B, epoch, seed),
                                      sampler_oracle.feistel_batch(N, B, epoch, seed))


def test_prefetcher_yields_epochs_in_order():
    from rfm_b200.sampler import LegacyBatchPrefetcher
    pf = LegacyBatchPrefetcher(5000, 100, range(3, 12), n_threads=3)
    try:
        for epoch in range(3, 12):
            np.testing.assert_array_equal(pf.next(), sampler_oracle.legacy_batch(5000, 100, epoch))
    finally:
        pf.close()


def test_no_device_is_an_error_not_a_fallback():
    n = ctypes.c_int()
    _capi.check(_capi.lib().rfm_device_count(ctypes.byref(n)))
    if n.value > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.Context(0)
    from rfm_b200.fm import FactorizationMachines
    m = FactorizationMachines("IPS", 1, 4, 0.1, 2, 0, 5)
    from scipy.sparse import identity
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.predict(X=identity(5, format="csr"))


_NULL_PROBE = r"""
import ctypes, json, sys
from rfm_b200 import _capi
L = _capi.lib()
out = {}
for name, (argtypes, _) in _capi._SIGNATURES.items():
    if name in ("rfm_abi_version", "rfm_last_error", "rfm_device_count", "rfm_legacy_batch", "rfm_feistel_batch",
                "rfm_host_alloc", "rfm_host_free", "rfm_host_register", "rfm_host_unregister"):
        continue
    args = []
    for t in argtypes:
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents"):
            args.append(None)
        elif t is ctypes.c_double:
            args.append(0.0)
        else:
            args.append(0)
    rc = getattr(L, name)(*args)
    out[name] = [rc, L.rfm_last_error().decode("utf-8", "replace")]
print(json.dumps(out))
"""


def test_null_arguments_are_status_codes_never_crashes():
    """a bad call returns a status and a message naming the entry point; it never
    takes the interpreter down. Every handle-taking entry point is called with NULL / zero arguments in a
    child process (so that a segfault would fail this test, not the run). destroy(NULL) is a no-op like free()."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(sys.path))
    p = subprocess.run([sys.executable, "-c", _NULL_PROBE], capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0, "the library crashed on NULL arguments:\n" + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert len(res) >= 50
    for name, (rc, msg) in res.items():
        if name.endswith("_destroy"):
            assert rc == 0, name
        else:
            assert rc == 1, (name, rc, msg)                       # RFM_ERR_INVALID -> ValueError in the shim
            assert msg.startswith(name), (name, msg)


def test_feistel_batches_and_row_gather_match_their_serial_counterparts():
    """Host helpers of the "upload only what the fit samples" idea (round2-prework): the threaded multi-epoch
    Feistel batches equal rfm_feistel_batch epoch by epoch (any slice), and the threaded CSR gather equals scipy's
    X[rows] with labels[rows] / pscores[rows] (repeated rows, empty rows, int64 row pointers)."""
    import scipy.sparse as sp
    N, B = 50_000, 2048
    a = _capi.feistel_batches(N, B, 12345, 3, 4, n_threads=3)
    for e in range(4):
        np.testing.assert_array_equal(a[e], _capi.feistel_batch(N, B, 3 + e, 12345))
    np.testing.assert_array_equal(_capi.feistel_batches(N, B, 12345, 3, 4, begin=100, count=700, n_threads=2),
                                  a[:, 100:800])
    rng = np.random.default_rng(0)
    X = sp.random(N, 300, density=0.01, format="csr", random_state=1, dtype=np.float64)
    X.sum_duplicates()
    y, ps = rng.integers(0, 2, N), rng.uniform(0.1, 1.0, N)
    rows = np.concatenate([a.reshape(-1), a[0, :50]])                       # repeats across "epochs"
    for indptr_dtype in (np.int32, np.int64):
        X.indptr = X.indptr.astype(indptr_dtype)
        for nt in (1, 4):
            sub, sy, sps = _capi.gather_rows(X, y, ps, rows, n_threads=nt)
            ref = X[rows]
            np.testing.assert_array_equal(sub.indptr, ref.indptr)
            np.testing.assert_array_equal(sub.indices, ref.indices)
            np.testing.assert_array_equal(sub.data, ref.data)
            np.testing.assert_array_equal(sy, y[rows])
            np.testing.assert_array_equal(sps, ps[rows])
    with pytest.raises(ValueError, match="outside"):
        _capi.gather_rows(X, y, ps, np.array([0, N]))
