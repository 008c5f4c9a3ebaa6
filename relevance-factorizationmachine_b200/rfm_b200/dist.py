"""Data-parallel FM training over one 8xB200 box.

One process per GPU (``torchrun``); ``torch.distributed`` is plumbing only: rendezvous and the
all-reduce of the dense gradient ``[sum_e | dw (n) | dV (n x kpad)]`` that ``rfm_fm_grad_epoch``
leaves in device memory. The FM gradient is a plain sum over batch rows (``src/fm.py:142,153,
178-180``), so the step shards naturally:

    every rank: same parameters, same global batch order (the sampler is deterministic)
    rank r    : forward + residual + segmented column reduction over ITS slice of the batch
    all ranks : all-reduce(sum) of the gradient buffer             <- the one exchange step
    every rank: identical dense apply  w0 += lr*g0, w += lr*dw, V += lr*dV
    losses    : per-rank partial sums over the batch slice / a val slice; they ride along in the NEXT
                step's gradient all-reduce (two spare header slots), so a step has ONE collective

With one rank this is exactly the single-GPU step; with G ranks results differ from it only by the
association of the cross-rank sum. MF's sequential per-sample semantics do not shard ("replicas only").

``DataParallelFM`` holds the algorithm with the device work behind four callables so that the
composition can be exercised on CPU with the ``gloo`` backend and the NumPy oracle
(tests/test_dist_cpu.py); ``make_fm_stepper`` binds it to the C ABI for real runs.
"""
from __future__ import annotations

import os
from ctypes import byref, c_int64, c_void_p

import numpy as np


def slice_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced slices of range(n): the first n % world ranks get one extra element."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


class DistEnv:
    """Thin wrapper over torch.distributed (NCCL on GPUs, gloo on CPU for tests)."""

    def __init__(self, backend: str, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = device
        if not dist.is_initialized():
            kwargs = {}
            if backend == "nccl" and device is not None:
                kwargs["device_id"] = torch.device("cuda", device)
            dist.init_process_group(backend=backend, **kwargs)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.backend = backend

    def _scalar(self, value, op):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64,
                              device="cuda:%d" % self.device if self.backend == "nccl" else "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, value):
        return self._scalar(value, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, value):
        return self._scalar(value, self.dist.ReduceOp.SUM)

    def barrier(self):
        if self.backend == "nccl":
            self.torch.cuda.synchronize(self.device)
        self.dist.barrier()
        if self.backend == "nccl":
            self.torch.cuda.synchronize(self.device)

    def all_reduce_sum(self, tensor):
        self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)

    def shutdown(self):
        if self.dist.is_initialized():
            self.dist.destroy_process_group()


def init(local_rank: int) -> DistEnv:
    """NCCL process group for this rank's GPU; reads RANK/WORLD_SIZE/MASTER_* from the env."""
    import torch
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    return DistEnv("nccl", local_rank)


class _DeviceArray:
    """Exposes library-owned device memory to torch via __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def device_tensor(ptr: int, n: int, dtype: str, device: int):
    import torch
    typestr = "<f8" if dtype == "float64" else "<f4"
    return torch.as_tensor(_DeviceArray(ptr, n, typestr), device="cuda:%d" % device)


def csr_ranges(indptr, n_rows: int, world: int, es: int):
    """Per rank: the byte ranges (offset, length) of its row slice inside the device arrays col (int32 per
    non-zero), val (es bytes per non-zero) and targets (es bytes per row). Pure host arithmetic."""
    out = []
    for r in range(world):
        rb, re = slice_bounds(n_rows, world, r)
        z0, z1 = int(indptr[rb]), int(indptr[re])
        out.append(((z0 * 4, (z1 - z0) * 4), (z0 * es, (z1 - z0) * es), (rb * es, (re - rb) * es)))
    return out


def sharded_csr_rows(ctx, X, labels, pscores, dtype: str, env: DistEnv):
    """The replicated device CSR every rank trains on, built with 1/G of the PCIe traffic: rank r copies only rows
    ``slice_bounds(N, G, r)`` from the host (``rfm_csr_create_range``), then every slice is broadcast from its
    owner into the same offsets on all ranks over NVLink (3 NCCL broadcasts per rank: col, val, targets). The
    result is byte-identical to ``CsrRows(ctx, X, labels, pscores, dtype)`` on every rank, so the sampler, the
    kernels and the results do not change. All ranks must call this with the same matrix."""
    from . import _capi
    if env.backend != "nccl":
        raise RuntimeError("sharded_csr_rows needs the NCCL backend (the slices are exchanged on the device)")
    torch = env.torch
    n = X.shape[0]
    begin, end = slice_bounds(n, env.world, env.rank)
    err = None
    try:
        rows = _capi.CsrRows(ctx, X, labels, pscores, dtype, row_range=(begin, end))
    except Exception as e:                      # noqa: BLE001 -- every rank must learn about it before the collectives
        rows, err = None, e
    if env.max_over_ranks(0.0 if err is None else 1.0) > 0:
        raise err if err is not None else RuntimeError("sharded_csr_rows: the upload failed on another rank")
    ctx.synchronize()
    _, col, val, targets = rows.device_ptrs()
    es = 8 if dtype == "float64" else 4
    dev = "cuda:%d" % env.device
    for src, ranges in enumerate(csr_ranges(rows.indptr_host, n, env.world, es)):
        for base, (off, nbytes) in zip((col, val, targets), ranges):
            if nbytes:
                t = torch.as_tensor(_DeviceArray(base + off, nbytes, "|u1"), device=dev)
                env.dist.broadcast(t, src=src)
    torch.cuda.synchronize(env.device)
    return rows


def sharded_factored_rows(ctx, X, labels, pscores, dtype: str, env: DistEnv):
    """Factored rows for a data-parallel fit with 1/G of the PCIe traffic per rank: rank r copies only rows
    ``slice_bounds(N, G, r)`` of the per-interaction arrays from the host (``rfm_factored_create_range``; the small
    per-entity tables go to every rank), then every slice of (user, item, context, target) is broadcast from its
    owner into the same offsets on all ranks over NVLink. Byte-identical to ``FactoredRows(ctx, X, labels, pscores,
    dtype)`` on every rank: eight ranks pulling the same 0.2 GB through the host's memory system was what bounded the
    8-GPU fit end to end. All ranks must call this with the same data."""
    from .factored import FactoredRows
    if env.backend != "nccl":
        raise RuntimeError("sharded_factored_rows needs the NCCL backend (the slices are exchanged on the device)")
    torch = env.torch
    n = X.shape[0]
    begin, end = slice_bounds(n, env.world, env.rank)
    err = None
    try:
        rows = FactoredRows(ctx, X, labels, pscores, dtype, row_range=(begin, end))
    except Exception as e:                      # noqa: BLE001 -- every rank must learn about it before the collectives
        rows, err = None, e
    if env.max_over_ranks(0.0 if err is None else 1.0) > 0:
        raise err if err is not None else RuntimeError("sharded_factored_rows: the upload failed on another rank")
    ctx.synchronize()
    user, item, cvals, targets = rows.device_ptrs()
    es = 8 if dtype == "float64" else 4
    arrays = [(user, 4), (item, 4)]
    if cvals and rows.n_ctx:
        arrays.append((cvals, rows.n_ctx * es))
    if rows.has_targets:
        arrays.append((targets, es))
    dev = "cuda:%d" % env.device
    for src in range(env.world):
        b, e = slice_bounds(n, env.world, src)
        for base, per_row in arrays:
            if e > b:
                t = torch.as_tensor(_DeviceArray(base + b * per_row, (e - b) * per_row, "|u1"), device=dev)
                env.dist.broadcast(t, src=src)
    torch.cuda.synchronize(env.device)
    rows.finalize()
    return rows


class DataParallelFM:
    """The DP step, independent of where the arithmetic runs.

    local_grad(begin, end, epoch) -> None   fill the gradient buffer for global-batch slice [begin, end)
    grad_tensor                              torch tensor aliasing that buffer (all-reduced in place)
    apply(lr) -> None                        params += lr * grad, identically on every rank
    local_loss_sums(begin, end, vbegin, vend) -> None   fill loss_tensor[0:2] with this rank's partial sums
    """

    def __init__(self, env: DistEnv, global_batch: int, n_val: int, lr: float, local_grad, grad_tensor, apply,
                 local_loss_sums, loss_tensor):
        self.env, self.global_batch, self.n_val, self.lr = env, global_batch, n_val, lr
        self.local_grad, self.grad_tensor, self.apply = local_grad, grad_tensor, apply
        self.local_loss_sums, self.loss_tensor = local_loss_sums, loss_tensor
        if global_batch < env.world:
            raise ValueError("global batch (%d) must be at least the number of ranks (%d)" % (global_batch, env.world))
        self.begin, self.end = slice_bounds(global_batch, env.world, env.rank)
        self.vbegin, self.vend = slice_bounds(n_val, env.world, env.rank)

    def step(self, epoch: int):
        """One reference epoch over the global batch, with ONE collective: the gradient all-reduce. The
        loss sums of the PREVIOUS step ride along in two spare slots of the gradient buffer's header
        ([sum_e, loss_batch, loss_val, pad | dw | dV]), so this returns the previous step's global
        [batch loss sum, val loss sum] (None on the first step); call ``flush()`` after the last step."""
        self.local_grad(self.begin, self.end, epoch)
        pending = getattr(self, "_pending", False)
        if pending:
            self.grad_tensor[1:3].copy_(self.loss_tensor)        # previous step's local sums (dtype-converting copy)
        self.env.all_reduce_sum(self.grad_tensor)
        prev = self.grad_tensor[1:3].to(self.loss_tensor.dtype).clone() if pending else None
        self.apply(self.lr)
        self.local_loss_sums(self.begin, self.end, self.vbegin, self.vend)
        self._pending = True
        return prev

    def flush(self):
        """Global loss sums of the last step (its own small all-reduce)."""
        if not getattr(self, "_pending", False):
            return None
        self.env.all_reduce_sum(self.loss_tensor)
        self._pending = False
        return self.loss_tensor.clone()


class NvlinkDataParallelFM(DataParallelFM):
    """The same step with the exchange done by the library's own kernel over NVLink peer memory
    (rfm_fm_dp_exchange_apply: barrier, rank-ordered reduce of this rank's slice in place, barrier, apply from
    the slice owners) instead of {NCCL all-reduce, apply}. torch.distributed only carries the IPC handles."""

    def __init__(self, env, global_batch, n_val, lr, local_grad, exchange_apply, local_loss_sums, loss_tensor,
                 prev_loss_tensor, status_tensor):
        super().__init__(env, global_batch, n_val, lr, local_grad, None, None, local_loss_sums, loss_tensor)
        self.exchange_apply, self.prev_loss_tensor, self.status_tensor = exchange_apply, prev_loss_tensor, status_tensor

    STATUS_EVERY = 64      # steps between reads of the exchange kernel's status word (a read drains the stream)

    def _check_status(self):
        status = int(self.status_tensor.item())
        if status:
            raise RuntimeError("rfm_fm_dp_exchange_apply: a cross-GPU barrier timed out (status %d); the parameters "
                               "of this fit are not trustworthy from that step on" % status)

    def step(self, epoch: int):
        self.local_grad(self.begin, self.end, epoch)
        pending = getattr(self, "_pending", False)
        self.exchange_apply(self.lr)                 # carries the previous step's local loss sums in the header
        prev = self.prev_loss_tensor.clone() if pending else None
        self.local_loss_sums(self.begin, self.end, self.vbegin, self.vend)
        self._pending = True
        self._steps = getattr(self, "_steps", 0) + 1
        if self._steps % self.STATUS_EVERY == 0:     # a timed-out barrier skips its reduce/apply: stop here, not after the fit
            self._check_status()
        return prev

    def flush(self):
        self._check_status()
        out = super().flush()
        self._check_status()
        return out


def make_fm_dp(model, trainer, env: DistEnv, global_batch: int, n_val: int, lr: float, batch_source,
               exchange: str = "auto"):
    """Bind DataParallelFM to the C ABI.

    batch_source(epoch) -> None | np.ndarray: None selects the device Feistel sampler; an int64
    array is the epoch's GLOBAL batch order (legacy sampler), identical on every rank.
    exchange: "nvlink" (the library's fused peer-memory kernel; one node, at most 8 ranks), "nccl"
    (torch.distributed all-reduce + apply), or "auto" (nvlink on an NCCL group of <= 8 ranks unless
    RFM_DP_EXCHANGE=nccl).
    """
    from . import _capi
    from ._capi import check, lib, ptr
    if exchange == "auto":
        exchange = os.environ.get("RFM_DP_EXCHANGE", "nvlink" if env.backend == "nccl" and env.world <= 8 else "nccl")
    if exchange not in ("nvlink", "nccl"):
        raise ValueError("exchange must be 'nvlink', 'nccl' or 'auto'")
    n = c_int64()
    check(lib().rfm_fm_grad_size(trainer.handle, byref(n)))
    gp, lp = c_void_p(), c_void_p()
    if exchange == "nvlink":
        from ctypes import c_ubyte
        handle = (c_ubyte * 64)()
        check(lib().rfm_fm_dp_export(trainer.handle, handle))
        gathered = [None] * env.world
        env.dist.all_gather_object(gathered, bytes(handle))
        every = (c_ubyte * (64 * env.world)).from_buffer_copy(b"".join(gathered))
        check(lib().rfm_fm_dp_connect(trainer.handle, env.rank, env.world, every))
        env.barrier()                                   # every rank has mapped every region
    else:
        check(lib().rfm_fm_grad_ptr_dev(trainer.handle, byref(gp)))
    check(lib().rfm_fm_loss_sums_ptr_dev(trainer.handle, byref(lp)))
    grad_tensor = device_tensor(gp.value, n.value, model.dtype, env.device) if exchange == "nccl" else None
    loss_tensor = device_tensor(lp.value, 2, "float64", env.device)
    state = {}

    def local_grad(begin, end, epoch):
        rows = batch_source(epoch)
        state["rows"] = rows
        if rows is None:
            check(lib().rfm_fm_grad_epoch_sampled(trainer.handle, model.seed & 0xFFFFFFFF, epoch, begin, end - begin))
        else:
            mine = np.ascontiguousarray(rows[begin:end], dtype=np.int64)
            check(lib().rfm_fm_grad_epoch(trainer.handle, ptr(mine), end - begin))

    def apply(step_lr):
        check(lib().rfm_fm_apply_grad(trainer.handle, step_lr))

    def local_loss_sums(begin, end, vbegin, vend):
        check(lib().rfm_fm_loss_sums(trainer.handle, None, end - begin, vbegin, vend))

    if exchange == "nvlink":
        pp, sp = c_void_p(), c_void_p()
        check(lib().rfm_fm_dp_prev_loss_ptr_dev(trainer.handle, byref(pp), byref(sp)))
        import torch
        status = torch.as_tensor(_DeviceArray(sp.value, 1, "<u4"), device="cuda:%d" % env.device)

        def exchange_apply(step_lr):
            check(lib().rfm_fm_dp_exchange_apply(trainer.handle, step_lr))

        return NvlinkDataParallelFM(env, global_batch, n_val, lr, local_grad, exchange_apply, local_loss_sums,
                                    loss_tensor, device_tensor(pp.value, 2, "float64", env.device), status)
    return DataParallelFM(env, global_batch, n_val, lr, local_grad, grad_tensor, apply, local_loss_sums, loss_tensor)


def item_shard(n_items: int, world: int, rank: int, tile: int = 256):
    """Contiguous item range of a rank, cut at multiples of the scoring kernel's item tile."""
    n_tiles = (n_items + tile - 1) // tile
    begin, end = slice_bounds(n_tiles, world, rank)
    return min(begin * tile, n_items), min(end * tile, n_items)


def connect_scorer(scorer, env: DistEnv, k_cap: int = 120):
    """One-time set-up of a scorer's NVLink exchange (collective): allocate the exchange region, gather the CUDA-IPC
    handles through torch.distributed, map every peer's region."""
    from ._capi import check, lib
    from ctypes import c_ubyte
    if getattr(scorer, "_xchg", None) is not None:
        if scorer._xchg[0] != env.world or scorer._xchg[1] < k_cap:
            raise ValueError("scorer is already connected to %d ranks with k_cap %d" % scorer._xchg)
        return
    if env.world > 8:
        raise ValueError("the peer-memory exchange serves one node (at most 8 ranks)")
    handle = (c_ubyte * 64)()
    check(lib().rfm_topk_dp_export(scorer.handle, k_cap, handle))
    gathered = [None] * env.world
    env.dist.all_gather_object(gathered, bytes(handle))
    every = (c_ubyte * (64 * env.world)).from_buffer_copy(b"".join(gathered))
    check(lib().rfm_topk_dp_connect(scorer.handle, env.rank, env.world, every))
    env.barrier()                                       # every rank has mapped every region
    scorer._xchg = (env.world, k_cap)
    scorer._xchg_env = env


def sharded_topk(scorer, env: DistEnv, K: int, mode: str = "tensor", gather: bool = True, exchange: str = "auto",
                 copy: bool = True):
    """Item-sharded full-catalog top-K: every rank holds all users' factors and ranks its own
    slice of the catalog on its GPU.

    exchange="nvlink" (default on an NCCL group of <= 8 ranks): the library's own peer-memory exchange -- global
    collect thresholds from the union of the ranks' sampled group maxima, then rank r merges the sorted per-rank
    lists of the users it owns, ``slice_bounds(n_users, world, rank)``, reading the peers' lists directly
    (``rfm_topk_run_sharded``). exchange="nccl": two all-gathers of the per-rank lists and a merge of every user on
    every rank (the round-1 path, kept for comparison).

    gather=True: every rank returns the same global ``(items, scores)`` for all users (the owned ranges are
    all-gathered). gather=False (nvlink only): returns ``(user_begin, user_end, items, scores)`` for the users this
    rank owns -- what the metric reductions downstream need."""
    from . import _capi
    from ._capi import check, lib, ptr
    if env.backend != "nccl":
        raise RuntimeError("sharded_topk needs the NCCL backend (results stay on the device)")
    if exchange == "auto":
        exchange = os.environ.get("RFM_TOPK_EXCHANGE", "nvlink" if env.world <= 8 else "nccl")
    if exchange not in ("nvlink", "nccl"):
        raise ValueError("exchange must be 'nvlink', 'nccl' or 'auto'")
    torch = env.torch
    n_users = scorer.n_users
    dev = "cuda:%d" % env.device
    if exchange == "nvlink":
        connect_scorer(scorer, env, max(K, 1))
        ub, ue = slice_bounds(n_users, env.world, env.rank)
        rng = (c_int64 * 2)()
        stats = (c_int64 * 4)()
        if copy:
            items = np.empty((ue - ub, K), dtype=np.int32)
            scores = np.empty((ue - ub, K), dtype=np.float64)
            check(lib().rfm_topk_run_sharded(scorer.handle, K, 0 if mode == "tensor" else 1, ptr(items), ptr(scores),
                                             rng, stats))
        else:
            import ctypes
            check(lib().rfm_topk_run_sharded(scorer.handle, K, 0 if mode == "tensor" else 1, None, None, rng, stats))
            pi, ps = c_void_p(), c_void_p()
            check(lib().rfm_topk_result_host(scorer.handle, K, byref(pi), byref(ps)))
            n = (ue - ub) * K
            items = np.ctypeslib.as_array(ctypes.cast(pi, ctypes.POINTER(ctypes.c_int32)), shape=(n,)).reshape(ue - ub, K)
            scores = np.ctypeslib.as_array(ctypes.cast(ps, ctypes.POINTER(ctypes.c_double)), shape=(n,)).reshape(ue - ub, K)
        assert (rng[0], rng[1]) == (ub, ue)
        scorer.last_stats = {"tensor_core_path": bool(stats[0]), "users_ranked_exactly": int(stats[1]),
                             "candidates": int(stats[2]), "sample_stride": int(stats[3])}
        if not gather:
            return ub, ue, items, scores
        width = -(-n_users // env.world)
        mine_i = torch.full((width, K), -1, dtype=torch.int32, device=dev)
        mine_s = torch.full((width, K), float("-inf"), dtype=torch.float64, device=dev)
        mine_i[: ue - ub] = torch.from_numpy(np.ascontiguousarray(items)).to(dev)
        mine_s[: ue - ub] = torch.from_numpy(np.ascontiguousarray(scores)).to(dev)
        all_i = torch.empty((env.world, width, K), dtype=torch.int32, device=dev)
        all_s = torch.empty((env.world, width, K), dtype=torch.float64, device=dev)
        env.dist.all_gather_into_tensor(all_i, mine_i)
        env.dist.all_gather_into_tensor(all_s, mine_s)
        out_i = np.empty((n_users, K), dtype=np.int32)
        out_s = np.empty((n_users, K), dtype=np.float64)
        hi, hs = all_i.cpu().numpy(), all_s.cpu().numpy()
        for r in range(env.world):
            b, e = slice_bounds(n_users, env.world, r)
            out_i[b:e] = hi[r, : e - b]
            out_s[b:e] = hs[r, : e - b]
        return out_i, out_s
    begin, end = item_shard(scorer.n_items, env.world, env.rank)
    if end > begin:
        check(lib().rfm_topk_run(scorer.handle, K, 0 if mode == "tensor" else 1, begin, end, None, None, None))
        ip, sp = c_void_p(), c_void_p()
        check(lib().rfm_topk_result_ptr_dev(scorer.handle, byref(ip), byref(sp)))
        mine_items = torch.as_tensor(_DeviceArray(ip.value, n_users * K, "<i4"), device=dev)
        mine_scores = torch.as_tensor(_DeviceArray(sp.value, n_users * K, "<f8"), device=dev)
    else:
        mine_items = torch.full((n_users * K,), -1, dtype=torch.int32, device=dev)
        mine_scores = torch.full((n_users * K,), float("-inf"), dtype=torch.float64, device=dev)
    all_items = torch.empty((env.world, n_users * K), dtype=torch.int32, device=dev)
    all_scores = torch.empty((env.world, n_users * K), dtype=torch.float64, device=dev)
    env.dist.all_gather_into_tensor(all_items, mine_items)
    env.dist.all_gather_into_tensor(all_scores, mine_scores)
    torch.cuda.current_stream(env.device).synchronize()
    items = np.empty((n_users, K), dtype=np.int32)
    scores = np.empty((n_users, K), dtype=np.float64)
    check(lib().rfm_topk_merge_dev(scorer.ctx.handle, n_users, K, env.world, c_void_p(all_items.data_ptr()),
                                   c_void_p(all_scores.data_ptr()), ptr(items), ptr(scores)))
    return items, scores


def sharded_predict(model, X, env: DistEnv) -> np.ndarray:
    """Row-sharded ``predict`` (independent rows): every rank scores a contiguous slice of
    the rows of ``X`` with its replica of the parameters and the slices are all-gathered; every rank returns
    the full vector, equal to the single-process ``model.predict(X)`` bit for bit (a row's score does not
    depend on which rank computes it)."""
    torch = env.torch
    n = X.shape[0]
    begin, end = slice_bounds(n, env.world, env.rank)
    width = -(-n // env.world) if n else 0                 # slices differ by at most one row: pad to the widest
    if env.backend == "nccl" and type(model).__name__ == "FactorizationMachines" and n > 0:
        # device path: the slice's scores are written by rfm_fm_predict_dev straight into the all-gather's send buffer,
        # the gathered [world][width] block comes back in ONE copy into page-locked memory
        from ctypes import c_void_p
        from ._capi import check, lib
        dev = "cuda:%d" % env.device
        model.sync_to_device()
        send = torch.zeros(max(width, 1), dtype=torch.float64, device=dev)
        if end > begin:
            rows = model._rows(X[begin:end])
            check(lib().rfm_fm_predict_dev(model._dev.handle, rows.handle, c_void_p(send.data_ptr())))
            model._ctx.synchronize()                       # the library's stream -> torch's stream
        recv = torch.empty(env.world * max(width, 1), dtype=torch.float64, device=dev)
        env.dist.all_gather_into_tensor(recv, send)
        host = torch.empty(recv.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(recv, non_blocking=True)
        torch.cuda.synchronize(env.device)
        flat = host.numpy()
        if n % env.world == 0:
            return flat[:n]                                # slices are contiguous: no reassembly
        out = np.empty(n)
        for r in range(env.world):
            b, e = slice_bounds(n, env.world, r)
            out[b:e] = flat[r * width: r * width + (e - b)]
        return out
    mine = model.predict(X=X[begin:end]) if end > begin else np.empty(0)
    dev = "cuda:%d" % env.device if env.backend == "nccl" else "cpu"
    buf = torch.zeros(max(width, 1), dtype=torch.float64, device=dev)
    buf[: end - begin] = torch.from_numpy(np.ascontiguousarray(mine, dtype=np.float64)).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(env.world)]
    env.dist.all_gather(gathered, buf)
    out = np.empty(n)
    for r, t in enumerate(gathered):
        b, e = slice_bounds(n, env.world, r)
        out[b:e] = t[: e - b].cpu().numpy()
    return out
