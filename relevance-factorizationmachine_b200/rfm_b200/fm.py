"""``FactorizationMachines`` -- drop-in for the reference class (``src/fm.py:16-187``).

Same dataclass fields in the same order (``estimator, n_epochs, n_factors, lr, batch_size,
seed, n_features, alpha=2.0, evaluator=None``), same ``fit(train, val) -> (train_loss,
val_loss)``, ``predict(X)``, parameter holders ``w0 / w / V`` and ``val_metrics``. Trailing
optional fields select the device-side behaviour; their defaults reproduce the reference.

The arithmetic runs in ``librfm_b200.so`` (``csrc/fm.cu``); this file is host plumbing:
legacy-RNG initialisation (``src/fm.py:31-48``), batch order (``src/fm.py:72-79``), uploads,
and the lazy host mirrors of the parameters.
"""
from __future__ import annotations

import os
import time
import weakref
from ctypes import byref, c_double, c_void_p
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import _capi
from ._capi import check, lib, ptr
from .base import EvalChain, PointwiseBaseRecommender
from .optimizer import SGD, Adam
from .sampler import LegacyBatchPrefetcher


class _FmDevice(_capi._Handle):
    _destroy = "rfm_fm_destroy"

    def __init__(self, ctx, n_features, n_factors, dtype):
        super().__init__()
        check(lib().rfm_fm_create(ctx.handle, n_features, n_factors, _capi.dtype_code(dtype), byref(self.handle)))


class _FmTrainer(_capi._Handle):
    _destroy = "rfm_fm_trainer_destroy"

    def __init__(self, model, train, val, max_batch, max_slots):
        super().__init__()
        self._keep = [model, train, val]
        check(lib().rfm_fm_trainer_create(model.handle, train.handle, val.handle if val is not None else None,
                                          max_batch, max_slots, byref(self.handle)))
        self.two_level = False

    def set_two_level(self, mode: int) -> bool:
        """mode 1: on; 2: where the cost model predicts a gain (``rfm_fm_trainer_set_two_level``)."""
        from ctypes import c_int32
        on = c_int32(0)
        check(lib().rfm_fm_trainer_set_two_level(self.handle, mode, byref(on)))
        self.two_level = bool(on.value)
        return self.two_level


@dataclass
class FactorizationMachines(PointwiseBaseRecommender):
    n_features: int
    alpha: float = 2.0
    evaluator: Optional[object] = None
    # ---- extensions (not in the reference; defaults keep reference behaviour) ----
    dtype: str = "float64"        # "float64" = parity mode, "float32" = perf mode
    sampler: str = "legacy"       # "legacy" = RandomState(epoch) order, "feistel" = device sampler
    device: int = 0
    progress: bool = False        # tqdm bar like the reference's (src/fm.py:71)
    distributed: object = None    # rfm_b200.dist.DistEnv: data-parallel fit over its ranks
    optimizer: str = "sgd"        # "sgd" = the reference's step; "adam" (spec: oracle/optimizer_oracle.py)
    l2: float = 0.0               # coupled L2 on w0, w, V (0 = reference behaviour)
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    materialize: str = "auto"     # factored rows -> stacked CSR on the device before a fit: "auto" (long fits), "always", "never"
    step: str = "auto"            # factored rows: "two_level" = per-entity aggregates (csrc/two_level.cuh), "flat" = one
                                  # gathered parameter row per stored non-zero, "auto" = two-level where it gathers >= 1.5 x less
    _dev: object = field(default=None, init=False, repr=False, compare=False)

    def __post_init__(self) -> None:
        if self.sampler not in ("legacy", "feistel"):
            raise ValueError("sampler must be 'legacy' or 'feistel'")
        if self.optimizer not in ("sgd", "adam"):
            raise ValueError("optimizer must be 'sgd' or 'adam'")
        if self.l2 < 0:
            raise ValueError("l2 must be >= 0")
        if self.materialize not in ("auto", "always", "never"):
            raise ValueError("materialize must be 'auto', 'always' or 'never'")
        if self.step not in ("auto", "flat", "two_level"):
            raise ValueError("step must be 'auto', 'flat' or 'two_level'")
        if self.distributed is not None and (self.optimizer != "sgd" or self.l2 != 0):
            raise ValueError("the data-parallel fit implements the reference's SGD step only")
        _capi.dtype_code(self.dtype)

        def holder(params):
            if self.optimizer == "adam":
                return Adam(params=params, lr=self.lr, beta1=self.beta1, beta2=self.beta2, eps=self.adam_eps, l2=self.l2)
            return SGD(params=params, lr=self.lr)

        np.random.seed(self.seed)                                   # global legacy RNG, like the reference
        self.w0 = holder(np.array([0.0]))
        limit = self.alpha * np.sqrt(6 / self.n_features)
        self.w = holder(np.random.uniform(low=-limit, high=limit, size=self.n_features))
        limit = self.alpha * np.sqrt(6 / self.n_factors)
        self.V = holder(np.random.uniform(low=-limit, high=limit, size=(self.n_features, self.n_factors)))
        if self.evaluator is not None:
            self.val_metrics = []
            self.model_name = "FM"
        self._ctx = None
        self._synced = None            # versions of (w0, w, V) the device copy corresponds to
        self._rows_cache: Dict[int, tuple] = {}
        self.last_fit_stats = {}

    # ---- device plumbing ---------------------------------------------------------------------
    def _context(self):
        if self._ctx is None:
            self._ctx = _capi.Context.default(self.device)
            self._dev = _FmDevice(self._ctx, self.n_features, self.n_factors, self.dtype)
            self._pin_params()
        return self._ctx

    def _pin_params(self) -> None:
        """Page-lock the holders' arrays (once, when the model first meets its device): fit() uploads them at its start
        and refreshes them in place at its end, and a pageable 9 MB copy each way costs ~1.5 ms of a 10 ms fit.
        The registration is dropped when the array is garbage collected. RFM_PIN_PARAMS=0 switches it off."""
        if os.environ.get("RFM_PIN_PARAMS", "1") == "0":
            return
        for holder in (self.w, self.V):
            a = holder.params
            if (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.owndata
                    and a.nbytes >= (1 << 20) and _capi.pin_array(a)):
                weakref.finalize(a, _capi.lib().rfm_host_unregister, c_void_p(a.ctypes.data))

    def _host_state(self):
        return (id(self.w0.params), self.w0.version, id(self.w.params), self.w.version,
                id(self.V.params), self.V.version)

    def sync_to_device(self, force: bool = False) -> None:
        """Upload host parameters if a holder was updated on the host since the last sync."""
        self._context()
        if force or self._synced != self._host_state():
            w0 = _capi.as_array(self.w0.params, np.float64).reshape(-1)
            w = _capi.as_array(self.w.params, np.float64)
            V = _capi.as_array(self.V.params, np.float64)
            if w.shape != (self.n_features,) or V.shape != (self.n_features, self.n_factors):
                raise ValueError("parameter arrays changed shape")
            check(lib().rfm_fm_set_params(self._dev.handle, ptr(w0), ptr(w), ptr(V)))
            self._synced = self._host_state()

    def sync_to_host(self) -> None:
        """Refresh the holders' ndarrays in place from the device master copy (straight into the arrays the holders
        already own when those are plain float64 buffers of the right shape: no intermediate copy)."""
        shapes = ((self.w0, (1,)), (self.w, (self.n_features,)), (self.V, (self.n_features, self.n_factors)))
        targets = []
        for holder, shape in shapes:
            p = holder.params
            if (isinstance(p, np.ndarray) and p.shape == shape and p.dtype == np.float64 and p.flags.writeable
                    and p.flags.c_contiguous):
                targets.append(p)
            else:
                holder.params = np.empty(shape)
                targets.append(holder.params)
        check(lib().rfm_fm_get_params(self._dev.handle, ptr(targets[0]), ptr(targets[1]), ptr(targets[2])))
        self._synced = self._host_state()

    def _rows(self, X, labels=None, pscores=None):
        """Device copy of a CSR matrix, cached per Python object so that the evaluator's features
        and the val set are uploaded once, not every epoch. The cache holds weak references to X, labels and
        pscores: a replaced object is seen, an array edited IN PLACE is not -- call ``reset_rows_cache()``."""
        key = (id(X), id(labels), id(pscores))
        hit = self._rows_cache.get(key)
        if hit is not None and _capi.refs_match(hit[0], X, labels, pscores):
            return hit[1]
        rows = self._make_rows(X, labels, pscores)
        refs = _capi.weak_refs(X, labels, pscores)
        if refs is not None:
            self._rows_cache[key] = (refs, rows)
            if len(self._rows_cache) > 8:
                self._rows_cache.pop(next(iter(self._rows_cache)))
        return rows

    def _make_rows(self, X, labels, pscores):
        from .factored import FactoredFeatures, FactoredRows
        if isinstance(X, FactoredFeatures):     # user table + item table + (user, item, ctx) records: assembled on the device
            env = self.distributed
            lab = None if labels is None else np.asarray(labels)
            if (env is not None and lab is not None and lab.dtype in (np.int8, np.int32, np.int64)
                    and env.backend == "nccl" and env.world > 1
                    and X.shape[0] >= int(os.environ.get("RFM_DP_UPLOAD_MIN_ROWS", "1000000"))
                    and os.environ.get("RFM_DP_UPLOAD", "sharded") == "sharded"):
                # data-parallel fit on a large train set: each rank uploads 1/G of the rows, NVLink carries the rest
                from .dist import sharded_factored_rows
                return sharded_factored_rows(self._context(), X, labels, pscores, self.dtype, env)
            return FactoredRows(self._context(), X, labels, pscores, self.dtype)
        if isinstance(X, _capi._Handle):        # rows that already live on the device (rfm_b200.clicks.GeneratedRows)
            if X.dtype != self.dtype:
                raise ValueError("device rows are %s, the model is %s" % (X.dtype, self.dtype))
            return X
        env = self.distributed
        if (env is not None and labels is not None and env.backend == "nccl" and env.world > 1
                and X.shape[0] >= int(os.environ.get("RFM_DP_UPLOAD_MIN_ROWS", "1000000"))
                and os.environ.get("RFM_DP_UPLOAD", "sharded") == "sharded"):
            # data-parallel fit on a large train set: each rank uploads 1/G of the rows, NVLink carries the rest
            from .dist import sharded_csr_rows
            return sharded_csr_rows(self._context(), X, labels, pscores, self.dtype, env)
        return _capi.CsrRows(self._context(), X, labels, pscores, self.dtype)

    def reset_rows_cache(self) -> None:
        self._rows_cache.clear()

    def _maybe_materialize(self, train_rows, val_rows):
        """Factored rows reach the device in a ninth of the bytes, but where everything is cached the row kernels run
        ~10 % faster on the stacked CSR. A long fit therefore assembles the CSR ON THE DEVICE from
        the factored rows (same entries, same order: not a bit changes) and trains on that."""
        from .factored import FactoredRows
        is_fac = lambda r: isinstance(r, FactoredRows) or type(r).__name__ == "GeneratedRows"
        if self.materialize == "never" or not (is_fac(train_rows) and is_fac(val_rows)):
            return train_rows, val_rows
        if self.materialize == "auto" and self.n_epochs < 128:
            return train_rows, val_rows
        try:
            return _capi.MaterializedRows(train_rows), _capi.MaterializedRows(val_rows)
        except (ValueError, _capi.RfmError):
            if self.materialize == "always":
                raise
            return train_rows, val_rows        # too large for 32-bit offsets or for the device: stay factored

    def _make_trainer(self, train_rows, val_rows, max_batch, max_slots):
        """(trainer, train rows, val rows): factored rows train with the two-level step where it is asked for or
        pays (then they stay factored); otherwise a long fit may assemble the stacked CSR on the device first."""
        is_fac = lambda r: getattr(r, "factored", False)
        if is_fac(train_rows) and (self.step == "two_level" or (self.step == "auto" and self.materialize != "always")):
            trainer = _FmTrainer(self._dev, train_rows, val_rows, max_batch, max_slots)
            if trainer.set_two_level(1 if self.step == "two_level" else 2):
                return trainer, train_rows, val_rows
            trainer.close()
        elif self.step == "two_level":
            raise ValueError("step='two_level' needs FactoredFeatures (or generated rows) as train['features']")
        train_rows, val_rows = self._maybe_materialize(train_rows, val_rows)
        return _FmTrainer(self._dev, train_rows, val_rows, max_batch, max_slots), train_rows, val_rows

    # ---- reference API -----------------------------------------------------------------------
    def fit(self, train, val) -> tuple:
        ctx = self._context()
        X = train["features"]
        n_rows = X.shape[0]
        if self.batch_size > n_rows:
            raise ValueError("Cannot sample %d out of arrays with dim %d when replace is False"
                             % (self.batch_size, n_rows))
        t_up = time.perf_counter()
        train_rows = self._rows(X, train["labels"], train["pscores"])
        val_rows = self._rows(val["features"], val["labels"], val["pscores"])
        self.sync_to_device()
        self._upload_seconds = time.perf_counter() - t_up
        if self.distributed is not None:
            return self._fit_data_parallel(train_rows, val_rows, n_rows)
        t_phase = time.perf_counter()
        trainer, train_rows, val_rows = self._make_trainer(train_rows, val_rows, self.batch_size, max(self.n_epochs, 1))
        phases = {"upload": self._upload_seconds, "trainer_create": time.perf_counter() - t_phase}
        epochs = range(self.n_epochs)
        prefetch = LegacyBatchPrefetcher(n_rows, self.batch_size, epochs) if self.sampler == "legacy" else None
        eval_rows, chain = None, None
        if self.evaluator is not None:
            eval_rows = self._rows(self.evaluator.features[self.model_name])
            if EvalChain.supported(self.evaluator):
                chain = EvalChain(self.evaluator, self.estimator, eval_rows.n_rows, self.n_epochs, self.device)
        it = epochs
        if self.progress:
            from tqdm import tqdm
            it = tqdm(epochs)
        launches0 = ctx.launch_count()
        t_phase = time.perf_counter()
        dense_opt = self.optimizer == "adam" or self.l2 != 0          # not the reference's fused SGD step
        try:
            for epoch in it:
                if dense_opt:
                    opt = _capi.Optimizer(kind=1 if self.optimizer == "adam" else 0, lr=self.lr, l2=self.l2,
                                          beta1=self.beta1, beta2=self.beta2, eps=self.adam_eps, step=epoch + 1)
                    idx = prefetch.next() if prefetch is not None else None
                    check(lib().rfm_fm_train_epoch_opt(trainer.handle, ptr(idx), self.seed & 0xFFFFFFFF, epoch,
                                                       self.batch_size, epoch, byref(opt)))
                elif prefetch is not None:
                    idx = prefetch.next()
                    check(lib().rfm_fm_train_epoch(trainer.handle, ptr(idx), self.batch_size, self.lr, epoch))
                else:
                    check(lib().rfm_fm_train_epoch_sampled(trainer.handle, self.seed & 0xFFFFFFFF, epoch,
                                                           self.batch_size, self.lr, epoch))
                if chain is not None:           # predict -> rank -> metric slot, all on the device, no synchronisation
                    check(lib().rfm_fm_predict_dev(self._dev.handle, eval_rows.handle, c_void_p(chain.scores_ptr)))
                    chain.after_epoch(epoch)
                elif eval_rows is not None:
                    scores = np.empty(eval_rows.n_rows)
                    check(lib().rfm_fm_predict(self._dev.handle, eval_rows.handle, ptr(scores)))
                    self.val_metrics.append(self.evaluator.evaluate(y_scores=scores, estimator=self.estimator))
            phases["enqueue_epochs"] = time.perf_counter() - t_phase
            t_phase = time.perf_counter()
            if chain is not None:
                self.val_metrics.extend(chain.finish(self.n_epochs))
            train_loss = np.empty(self.n_epochs)
            val_loss = np.empty(self.n_epochs)
            check(lib().rfm_fm_trainer_losses(trainer.handle, 0, self.n_epochs, ptr(train_loss), ptr(val_loss)))
            phases["drain_and_read_losses"] = time.perf_counter() - t_phase
        finally:
            if prefetch is not None:
                prefetch.close()
        launches = ctx.launch_count() - launches0
        t_phase = time.perf_counter()
        trainer.close()
        phases["trainer_destroy"] = time.perf_counter() - t_phase
        t_phase = time.perf_counter()
        self.sync_to_host()
        phases["download_params"] = time.perf_counter() - t_phase
        self.last_fit_stats = {"gpu_launches": launches, "two_level": getattr(trainer, "two_level", False),
                               "h2d_bytes_rows": train_rows.h2d_bytes + val_rows.h2d_bytes,
                               "upload_seconds": self._upload_seconds,
                               "phase_seconds": {k: round(v, 5) for k, v in phases.items()}}
        return train_loss.tolist(), val_loss.tolist()

    def _fit_data_parallel(self, train_rows, val_rows, n_rows) -> tuple:
        """Same epochs, the batch cut into one contiguous slice per rank, one gradient all-reduce
        per epoch, identical apply everywhere. Every rank returns the same losses/parameters."""
        from . import dist as rdist
        env = self.distributed
        ctx = self._context()
        begin, end = rdist.slice_bounds(self.batch_size, env.world, env.rank)
        phases = {"upload": self._upload_seconds}
        t_phase = time.perf_counter()
        trainer, train_rows, val_rows = self._make_trainer(train_rows, val_rows, max(end - begin, 1), 1)
        phases["trainer_create"] = time.perf_counter() - t_phase
        epochs = range(self.n_epochs)
        prefetch = LegacyBatchPrefetcher(n_rows, self.batch_size, epochs) if self.sampler == "legacy" else None
        source = (lambda epoch: prefetch.next()) if prefetch is not None else (lambda epoch: None)
        t_phase = time.perf_counter()
        dp = rdist.make_fm_dp(self, trainer, env, self.batch_size, val_rows.n_rows, self.lr, source)
        phases["dp_connect"] = time.perf_counter() - t_phase     # exchange region, IPC handles, peer mappings
        torch = env.torch
        hist = torch.zeros((max(self.n_epochs, 1), 2), dtype=torch.float64, device=dp.loss_tensor.device)
        eval_rows = self._rows(self.evaluator.features[self.model_name]) if self.evaluator is not None else None
        chain = None
        if eval_rows is not None and EvalChain.supported(self.evaluator):
            chain = EvalChain(self.evaluator, self.estimator, eval_rows.n_rows, self.n_epochs, self.device)
        launches0 = ctx.launch_count()
        t_phase = time.perf_counter()
        try:
            for epoch in epochs:
                prev = dp.step(epoch)                       # global loss sums of the previous epoch
                if prev is not None:
                    hist[epoch - 1].copy_(prev)
                if chain is not None:
                    check(lib().rfm_fm_predict_dev(self._dev.handle, eval_rows.handle, c_void_p(chain.scores_ptr)))
                    chain.after_epoch(epoch)
                elif eval_rows is not None:
                    scores = np.empty(eval_rows.n_rows)
                    check(lib().rfm_fm_predict(self._dev.handle, eval_rows.handle, ptr(scores)))
                    self.val_metrics.append(self.evaluator.evaluate(y_scores=scores, estimator=self.estimator))
            phases["enqueue_epochs"] = time.perf_counter() - t_phase
            t_phase = time.perf_counter()
            if chain is not None:
                self.val_metrics.extend(chain.finish(self.n_epochs))
            last = dp.flush()
            if last is not None:
                hist[self.n_epochs - 1].copy_(last)
            out = hist.cpu().numpy()
            phases["drain_and_read_losses"] = time.perf_counter() - t_phase
        finally:
            if prefetch is not None:
                prefetch.close()
        launches = ctx.launch_count() - launches0
        t_phase = time.perf_counter()
        trainer.close()
        phases["trainer_destroy"] = time.perf_counter() - t_phase   # closes the peer mappings, frees the region
        t_phase = time.perf_counter()
        self.sync_to_host()
        phases["download_params"] = time.perf_counter() - t_phase
        self.last_fit_stats = {"gpu_launches": launches, "two_level": getattr(trainer, "two_level", False),
                               "h2d_bytes_rows": train_rows.h2d_bytes + val_rows.h2d_bytes,
                               "upload_seconds": self._upload_seconds,
                               "phase_seconds": {k: round(v, 5) for k, v in phases.items()}}
        return (out[: self.n_epochs, 0] / self.batch_size).tolist(), (out[: self.n_epochs, 1] / val_rows.n_rows).tolist()

    def predict(self, X) -> np.ndarray:
        self.sync_to_device()
        rows = self._rows(X)
        out = np.empty(rows.n_rows)
        check(lib().rfm_fm_predict(self._dev.handle, rows.handle, ptr(out)))
        return out

    def logloss(self, data) -> float:
        """``_cross_entropy_loss(labels, predict(features), pscores)`` evaluated on the device."""
        self.sync_to_device()
        rows = self._rows(data["features"], data["labels"], data["pscores"])
        out = c_double()
        check(lib().rfm_fm_logloss(self._dev.handle, rows.handle, byref(out)))
        return out.value
