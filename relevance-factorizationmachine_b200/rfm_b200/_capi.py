"""ctypes binding of ``librfm_b200.so`` (C ABI declared in ``include/rfm_b200.h``).

There is deliberately no fallback: if the library is missing or no B200 is visible, every
compute entry point raises ``RuntimeError`` naming what is missing.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librfm_b200.so")

RFM_F32, RFM_F64 = 0, 1
RANK_NCOLS = 12
RANK_COLS = dict(DCG_SUM=0, IPSDCG_SUM=1, ME_SUM=2, ME_COUNT=3, RECALL_SUM=4, MAP_SUM=5, USERS=6, COVERED=7)

_lib = None

_P = c_void_p
_SIGNATURES = {
    "rfm_abi_version": ([], c_int),
    "rfm_last_error": ([], c_char_p),
    "rfm_device_count": ([POINTER(c_int)], c_int),
    "rfm_ctx_create": ([c_int, _P, POINTER(_P)], c_int),
    "rfm_ctx_destroy": ([_P], c_int),
    "rfm_ctx_synchronize": ([_P], c_int),
    "rfm_ctx_launch_count": ([_P, POINTER(c_int64)], c_int),
    "rfm_ctx_timer_start": ([_P], c_int),
    "rfm_ctx_timer_stop_ms": ([_P, POINTER(c_double)], c_int),
    "rfm_ctx_profile_begin": ([_P], c_int),
    "rfm_ctx_profile_end": ([_P, c_char_p, c_size_t], c_int),
    "rfm_host_register": ([_P, c_size_t], c_int),
    "rfm_host_unregister": ([_P], c_int),
    "rfm_host_alloc": ([c_size_t, POINTER(_P)], c_int),
    "rfm_host_free": ([_P], c_int),
    "rfm_legacy_batch": ([c_int64, c_int64, c_uint32, _P, _P], c_int),
    "rfm_feistel_batch": ([c_int64, c_int64, c_uint32, c_uint32, _P], c_int),
    "rfm_feistel_batches": ([c_int64, c_int64, c_uint32, c_uint32, c_int32, c_int64, c_int64, _P, c_int32], c_int),
    "rfm_csr_gather_rows": ([c_int64, _P, c_int, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, c_int32], c_int),
    "rfm_csr_create": ([_P, c_int64, c_int64, _P, c_int, _P, _P, _P, _P, c_int, POINTER(_P)], c_int),
    "rfm_csr_create_range": ([_P, c_int64, c_int64, _P, c_int, _P, _P, _P, _P, c_int, c_int64, c_int64, POINTER(_P)],
                             c_int),
    "rfm_csr_device_ptrs": ([_P, POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P)], c_int),
    "rfm_factored_create": ([_P, c_int64, _P, c_int32, _P, c_int32, _P, c_int32, _P, c_int32, _P, c_int, POINTER(_P)],
                            c_int),
    "rfm_factored_create_range": ([_P, c_int64, _P, c_int32, _P, c_int32, _P, c_int32, _P, c_int32, _P, _P, c_int64, c_int,
                                   c_int64, c_int64, POINTER(_P)], c_int),
    "rfm_rows_device_ptrs": ([_P, POINTER(_P), POINTER(_P), POINTER(_P), POINTER(_P)], c_int),
    "rfm_factored_finalize": ([_P], c_int),
    "rfm_factored_create_item_pscores": ([_P, c_int64, _P, c_int32, _P, c_int32, _P, c_int32, _P, c_int32, _P, c_int64,
                                          c_int, POINTER(_P)], c_int),
    "rfm_factored_generate": ([_P, c_int64, _P, _P, c_int32, c_int, POINTER(_P)], c_int),
    "rfm_rows_materialize": ([_P, POINTER(_P)], c_int),
    "rfm_rows_download": ([_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P], c_int),
    "rfm_csr_set_targets": ([_P, _P], c_int),
    "rfm_csr_destroy": ([_P], c_int),
    "rfm_fm_create": ([_P, c_int64, c_int32, c_int, POINTER(_P)], c_int),
    "rfm_fm_destroy": ([_P], c_int),
    "rfm_fm_set_params": ([_P, _P, _P, _P], c_int),
    "rfm_fm_get_params": ([_P, _P, _P, _P], c_int),
    "rfm_fm_predict": ([_P, _P, _P], c_int),
    "rfm_fm_predict_dev": ([_P, _P, _P], c_int),
    "rfm_fm_logloss": ([_P, _P, POINTER(c_double)], c_int),
    "rfm_fm_trainer_create": ([_P, _P, _P, c_int64, c_int64, POINTER(_P)], c_int),
    "rfm_fm_trainer_destroy": ([_P], c_int),
    "rfm_fm_trainer_set_two_level": ([_P, c_int32, POINTER(c_int32)], c_int),
    "rfm_fm_dp_trace": ([_P, _P], c_int),
    "rfm_fm_train_epoch": ([_P, _P, c_int64, c_double, c_int64], c_int),
    "rfm_fm_train_epoch_sampled": ([_P, c_uint32, c_uint32, c_int64, c_double, c_int64], c_int),
    "rfm_fm_grad_size": ([_P, POINTER(c_int64)], c_int),
    "rfm_fm_grad_ptr_dev": ([_P, POINTER(_P)], c_int),
    "rfm_fm_grad_epoch": ([_P, _P, c_int64], c_int),
    "rfm_fm_grad_epoch_sampled": ([_P, c_uint32, c_uint32, c_int64, c_int64], c_int),
    "rfm_fm_apply_grad": ([_P, c_double], c_int),
    "rfm_fm_train_epoch_opt": ([_P, c_void_p, c_uint32, c_uint32, c_int64, c_int64, c_void_p], c_int),
    "rfm_fm_dp_export": ([_P, c_void_p], c_int),
    "rfm_fm_dp_connect": ([_P, c_int32, c_int32, c_void_p], c_int),
    "rfm_fm_dp_exchange_apply": ([_P, c_double], c_int),
    "rfm_fm_dp_prev_loss_ptr_dev": ([_P, POINTER(c_void_p), POINTER(c_void_p)], c_int),
    "rfm_fm_loss_sums_ptr_dev": ([_P, POINTER(_P)], c_int),
    "rfm_fm_loss_sums": ([_P, _P, c_int64, c_int64, c_int64], c_int),
    "rfm_fm_trainer_losses": ([_P, c_int64, c_int64, _P, _P], c_int),
    "rfm_pairs_create": ([_P, c_int64, _P, _P, _P, c_int, POINTER(_P)], c_int),
    "rfm_pairs_set_targets": ([_P, _P], c_int),
    "rfm_pairs_destroy": ([_P], c_int),
    "rfm_mf_create": ([_P, c_int64, c_int64, c_int32, c_int, POINTER(_P)], c_int),
    "rfm_mf_destroy": ([_P], c_int),
    "rfm_mf_set_params": ([_P, _P, _P, _P, _P, c_double], c_int),
    "rfm_mf_get_params": ([_P, _P, _P, _P, _P], c_int),
    "rfm_mf_predict": ([_P, _P, _P], c_int),
    "rfm_mf_predict_dev": ([_P, _P, _P], c_int),
    "rfm_mf_logloss": ([_P, _P, POINTER(c_double)], c_int),
    "rfm_mf_train_epoch": ([_P, _P, _P, _P, c_int64, c_double, c_double, POINTER(c_double), POINTER(c_double)],
                           c_int),
    "rfm_topk_create": ([_P, c_int64, c_int64, c_int32, POINTER(_P)], c_int),
    "rfm_topk_destroy": ([_P], c_int),
    "rfm_topk_set_factors": ([_P, _P, _P, _P, _P, c_double], c_int),
    "rfm_topk_run": ([_P, c_int32, c_int32, c_int64, c_int64, _P, _P, _P], c_int),
    "rfm_topk_result_ptr_dev": ([_P, POINTER(_P), POINTER(_P)], c_int),
    "rfm_topk_result_host": ([_P, c_int32, POINTER(_P), POINTER(_P)], c_int),
    "rfm_topk_dp_export": ([_P, c_int32, c_void_p], c_int),
    "rfm_topk_dp_connect": ([_P, c_int32, c_int32, c_void_p], c_int),
    "rfm_topk_run_sharded": ([_P, c_int32, c_int32, _P, _P, _P, _P], c_int),
    "rfm_topk_merge_dev": ([_P, c_int64, c_int32, c_int32, _P, _P, _P, _P], c_int),
    "rfm_ranker_create": ([_P, c_int64, _P, _P, _P, _P, c_int64, POINTER(_P)], c_int),
    "rfm_ranker_destroy": ([_P], c_int),
    "rfm_ranker_num_users": ([_P, POINTER(c_int64)], c_int),
    "rfm_ranker_set_user_totals": ([_P, _P], c_int),
    "rfm_ranker_evaluate": ([_P, _P, _P, c_int32, _P, _P, _P], c_int),
    "rfm_ranker_scores_ptr_dev": ([_P, POINTER(_P)], c_int),
    "rfm_ranker_evaluate_dev": ([_P, _P, _P, c_int32, c_int64, c_int64], c_int),
    "rfm_ranker_read_slots": ([_P, c_int64, c_int64, c_int32, _P], c_int),
    "rfm_catalog_eval_create": ([_P, c_int64, c_int64, _P, _P, _P, _P, POINTER(_P)], c_int),
    "rfm_catalog_eval_destroy": ([_P], c_int),
    "rfm_catalog_eval_run": ([_P, _P, c_int32, c_int64, c_int64, _P, c_int32, _P, _P], c_int),
}

# every symbol include/rfm_b200.h declares (tests check the library exports all of them)
DECLARED_SYMBOLS = tuple(_SIGNATURES)


class RfmError(RuntimeError):
    pass


ABI_VERSION = 6        # RFM_ABI_VERSION of include/rfm_b200.h this shim was written against


def lib():
    """Load the shared library once; raise loudly when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "rfm_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if handle.rfm_abi_version() != ABI_VERSION:
            raise RuntimeError("rfm_b200: %s has ABI version %d, this shim needs %d -- rebuild the library"
                               % (LIB_PATH, handle.rfm_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().rfm_last_error().decode("utf-8", "replace")
        if status == 1:
            raise ValueError(msg)
        raise RfmError("rfm_b200 (status %d): %s" % (status, msg))


def ptr(a):
    """Raw pointer of a C-contiguous ndarray (or None)."""
    if a is None:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data_as(c_void_p)


def as_array(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def dtype_code(name) -> int:
    name = np.dtype(name).name if not isinstance(name, str) else name
    if name in ("float64", "f64", "double"):
        return RFM_F64
    if name in ("float32", "f32", "float"):
        return RFM_F32
    raise ValueError("dtype must be 'float64' or 'float32', got %r" % (name,))


class Context:
    """One device + one stream. Shared by every handle of a model."""

    _default = {}

    def __init__(self, device: int = 0, stream: int = 0):
        self.handle = c_void_p()
        n = c_int()
        check(lib().rfm_device_count(byref(n)))
        if n.value == 0:
            raise RuntimeError("rfm_b200: no CUDA device is visible and there is no CPU fallback")
        check(lib().rfm_ctx_create(device, c_void_p(stream) if stream else None, byref(self.handle)))
        self.device = device

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def synchronize(self):
        check(lib().rfm_ctx_synchronize(self.handle))

    def launch_count(self) -> int:
        out = c_int64()
        check(lib().rfm_ctx_launch_count(self.handle, byref(out)))
        return out.value

    def profile_begin(self):
        check(lib().rfm_ctx_profile_begin(self.handle))

    def profile_end(self) -> dict:
        """{kernel name: (launches, total ms)} of every launch since profile_begin."""
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib().rfm_ctx_profile_end(self.handle, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, count, ms = line.split("\t")
            out[name] = (int(count), float(ms))
        return out

    def timer_start(self):
        check(lib().rfm_ctx_timer_start(self.handle))

    def timer_stop_ms(self) -> float:
        out = c_double()
        check(lib().rfm_ctx_timer_stop_ms(self.handle, byref(out)))
        return out.value


class _Handle:
    _destroy = None

    def __init__(self):
        self.handle = c_void_p()
        self._keep = []

    def close(self):
        if self.handle and _lib is not None:
            getattr(_lib, self._destroy)(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def weak_refs(*objs):
    """Weak references to every non-None object, or None when one of them cannot be weakly referenced (the caller
    then skips its cache: an id() alone can be reused after garbage collection)."""
    import weakref
    try:
        return tuple(None if o is None else weakref.ref(o) for o in objs)
    except TypeError:
        return None


def refs_match(refs, *objs) -> bool:
    return refs is not None and all((r is None and o is None) or (r is not None and r() is o) for r, o in zip(refs, objs))


def integer_labels(labels, pscores):
    """(labels as int64, float64 targets or None): the C ABI takes integer labels and divides on the device exactly
    as NumPy would (``y / pscore``, src/fm.py:80); labels with a fractional part keep the reference's semantics by
    sending ``labels / pscores`` computed here in float64 (``rfm_*_set_targets``) instead of being truncated."""
    y = np.asarray(labels)
    if y.dtype.kind in "iub":
        return as_array(y, np.int64), None
    yf = as_array(y, np.float64)
    yi = yf.astype(np.int64)
    if np.array_equal(yi, yf):
        return yi, None
    return np.zeros(yf.shape, dtype=np.int64), np.ascontiguousarray(yf / as_array(pscores, np.float64))


class CsrRows(_Handle):
    """Device copy of a scipy CSR matrix (+ labels / pscores)."""

    _destroy = "rfm_csr_destroy"

    def __init__(self, ctx: Context, X, labels=None, pscores=None, dtype="float64", row_range=None):
        """row_range=(begin, end): copy only those rows from the host (the object keeps the full shape; the
        caller fills the rest on the device, see rfm_b200.dist.sharded_csr_rows)."""
        super().__init__()
        X = X.tocsr()
        if not X.has_canonical_format:
            X = X.copy()
            X.sum_duplicates()
        self.ctx, self.shape, self.dtype = ctx, X.shape, dtype
        indptr = np.ascontiguousarray(X.indptr)
        is64 = indptr.dtype == np.int64
        if not is64:
            indptr = as_array(indptr, np.int32)
        indices = as_array(X.indices, np.int32)
        data = as_array(X.data, np.float64)
        ps = None if pscores is None else as_array(pscores, np.float64)
        if labels is not None and (len(labels) != X.shape[0] or ps is None or ps.shape[0] != X.shape[0]):
            raise ValueError("labels/pscores must have one entry per row")
        y, targets = (None, None) if labels is None else integer_labels(labels, ps)
        self.n_rows = X.shape[0]
        if row_range is None:
            check(lib().rfm_csr_create(ctx.handle, X.shape[0], X.shape[1], ptr(indptr), int(is64), ptr(indices),
                                       ptr(data), ptr(y), ptr(ps), dtype_code(dtype), byref(self.handle)))
            self.h2d_bytes = (indptr.nbytes + indices.nbytes + data.nbytes
                              + (y.nbytes + ps.nbytes if y is not None else 0))
        else:
            begin, end = row_range
            check(lib().rfm_csr_create_range(ctx.handle, X.shape[0], X.shape[1], ptr(indptr), int(is64), ptr(indices),
                                             ptr(data), ptr(y), ptr(ps), dtype_code(dtype), begin, end,
                                             byref(self.handle)))
            nz = int(indptr[end]) - int(indptr[begin])
            self.h2d_bytes = indptr.nbytes + nz * 12 + ((end - begin) * 16 if y is not None else 0)
        self.indptr_host = indptr          # kept for range arithmetic (a view of the caller's array when possible)
        if targets is not None:
            check(lib().rfm_csr_set_targets(self.handle, ptr(targets)))

    def device_ptrs(self):
        """(row_ptr, col, val, targets) device addresses."""
        out = [c_void_p() for _ in range(4)]
        check(lib().rfm_csr_device_ptrs(self.handle, *[byref(p) for p in out]))
        return tuple(p.value for p in out)


class MaterializedRows(_Handle):
    """Stacked CSR assembled on the device from factored rows (``rfm_rows_materialize``)."""

    _destroy = "rfm_csr_destroy"

    def __init__(self, factored_rows):
        super().__init__()
        check(lib().rfm_rows_materialize(factored_rows.handle, byref(self.handle)))
        self.ctx, self.dtype, self.n_rows, self.shape = factored_rows.ctx, factored_rows.dtype, factored_rows.n_rows, \
            factored_rows.shape
        self.h2d_bytes = factored_rows.h2d_bytes        # what reached the device over PCIe


class Optimizer(ctypes.Structure):
    """``rfm_optimizer`` of include/rfm_b200.h."""
    _fields_ = [("kind", c_int32), ("reserved", c_int32), ("lr", c_double), ("l2", c_double), ("beta1", c_double),
                ("beta2", c_double), ("eps", c_double), ("step", c_int64)]


def pin_array(a: np.ndarray) -> bool:
    """Page-lock an ndarray in place (cudaHostRegister). Returns False if the driver refuses."""
    try:
        check(lib().rfm_host_register(ptr(a), a.nbytes))
        return True
    except RfmError:
        return False


def unpin_array(a: np.ndarray) -> None:
    lib().rfm_host_unregister(ptr(a))


def legacy_batch(n_rows: int, batch: int, epoch: int, scratch=None) -> np.ndarray:
    """``resample(replace=False, n_samples=batch, random_state=epoch)`` row ids (host, exact)."""
    out = np.empty(batch, dtype=np.int64)
    check(lib().rfm_legacy_batch(n_rows, batch, epoch & 0xFFFFFFFF, ptr(out), ptr(scratch)))
    return out


def feistel_batches(n_rows: int, batch: int, seed: int, epoch0: int, n_epochs: int, begin: int = 0, count=None,
                    n_threads: int = 0) -> np.ndarray:
    """(n_epochs, count) row ids: positions [begin, begin+count) of each epoch's Feistel batch (host threads)."""
    count = batch - begin if count is None else count
    out = np.empty((n_epochs, count), dtype=np.int64)
    check(lib().rfm_feistel_batches(n_rows, batch, seed & 0xFFFFFFFF, epoch0 & 0xFFFFFFFF, n_epochs, begin, count,
                                    ptr(out), n_threads))
    return out


def gather_rows(X, labels, pscores, rows: np.ndarray, n_threads: int = 0):
    """scipy's ``X[rows]`` (+ labels[rows], pscores[rows]) on host threads; returns a csr_matrix flagged canonical
    (rows of a canonical matrix are canonical) and the two target arrays."""
    from scipy.sparse import csr_matrix
    indptr = np.ascontiguousarray(X.indptr)
    is64 = indptr.dtype == np.int64
    if not is64:
        indptr = as_array(indptr, np.int32)
    indices, data = as_array(X.indices, np.int32), as_array(X.data, np.float64)
    y = None if labels is None else as_array(labels, np.int64)
    ps = None if pscores is None else as_array(pscores, np.float64)
    rows = as_array(rows, np.int64).reshape(-1)
    n_sel = rows.shape[0]
    out_ptr = np.empty(n_sel + 1, dtype=np.int64)
    args = (X.shape[0], ptr(indptr), int(is64), ptr(indices), ptr(data), ptr(y), ptr(ps), ptr(rows), n_sel)
    check(lib().rfm_csr_gather_rows(*args, ptr(out_ptr), None, None, None, None, n_threads))
    nnz = int(out_ptr[-1])
    out_idx, out_val = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.float64)
    out_y = None if y is None else np.empty(n_sel, dtype=np.int64)
    out_ps = None if ps is None else np.empty(n_sel, dtype=np.float64)
    check(lib().rfm_csr_gather_rows(*args, ptr(out_ptr), ptr(out_idx), ptr(out_val), ptr(out_y), ptr(out_ps),
                                    n_threads))
    sub = csr_matrix((out_val, out_idx, out_ptr), shape=(n_sel, X.shape[1]), copy=False)
    if X.has_canonical_format:
        sub.has_canonical_format = True
    return sub, out_y, out_ps


def feistel_batch(n_rows: int, batch: int, epoch: int, seed: int = 0) -> np.ndarray:
    out = np.empty(batch, dtype=np.int64)
    check(lib().rfm_feistel_batch(n_rows, batch, seed & 0xFFFFFFFF, epoch & 0xFFFFFFFF, ptr(out)))
    return out
