#!/usr/bin/env python
"""bench.py -- train interactions/s of the fused IPS-FM epoch on synthetic KuaiRec-big-shaped data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one reference epoch (src/fm.py:71-102): one minibatch of B interactions through
forward, IPS residual, the simultaneous w0/w/V update, the post-update batch loss and the val loss.

ours:       value     = K*B*N / device time of K steps (CUDA events), dataset resident in HBM,
                        batches drawn on the device (Feistel sampler, perf mode), float64.
            e2e       = the same metric through the public API, FactorizationMachines.fit(train, val)
                        on HOST (pinned) arrays: dataset upload, K epochs, loss read-back, parameter
                        download, wall clock around the call.
            roofline  = dominant kernel's algorithmic bytes / its CUDA-event time (separate profiled
                        pass of the same steps), against MEASURED_PEAKS.json's HBM copy bandwidth.
            cpu_baseline = the CPU oracle (NumPy/SciPy port of the reference step + the reference's
                        own sampler) on a bounded sample of the same workload, on every host core
                        (each epoch's batch split over forked workers; single-core figure alongside).
            scoring   = scored user-item pairs/s of full-catalog top-K (second half of the metric),
                        with its tensor-core roofline.
            clocks    = nvidia-smi samples from the warm-up to a run of identical untimed steps.
reference:  the CPU oracle port on every host core timed alone, same config / metric / unit (the
            reference is pure Python and cannot travel to the GPU box).
Under torchrun (N > 1): one rank per GPU, B per GPU (weak scaling), the library's NVLink exchange
kernel between ranks (RFM_DP_EXCHANGE=nccl for the torch.distributed all-reduce), time = max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "relevance-factorizationmachine_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOAD = "kuairec_big"
METRIC = "train_interactions_per_sec"
UNIT = "interactions/s"
N_USERS, N_ITEMS, N_TRAIN, N_VAL, K_FACTORS = 7176, 10728, 12_000_000, 2000, 64
LR = 9e-6   # conf/setting/kuairec.yaml:59 (FM, IPS)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="interactions per step per GPU")
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--rows", type=int, default=N_TRAIN, help="train interactions in the job")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
    ap.add_argument("--workload", default="kuairec_big", choices=["kuairec_big", "stress"],
                    help="kuairec_big = BASELINE configs[2] (the headline); stress = configs[4]-shaped: 1M users x 1M "
                         "items, k=128, 8 non-zeros per row, parameter table far larger than L2")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": ("IPS-FM, synthetic KuaiRec big_matrix shape (BASELINE.json configs[2])" if WORKLOAD == "kuairec_big"
                     else "IPS-FM stress, BASELINE.json configs[4] shape on one GPU's share: 1M users x 1M items, "
                          "k=128, 8 non-zeros per row"),
        "n_users": N_USERS, "n_items": N_ITEMS, "train_interactions": args.rows, "val_rows": N_VAL,
        "n_factors": K_FACTORS, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "lr": LR, "parallelism": "dp%d" % world if world > 1 else "single",
        "l2": "inputs larger than L2: each step gathers a fresh random batch from the resident "
              "train CSR; the parameter table V is legitimately L2-resident across steps",
    }


def make_stress_data(rows, seed):
    """configs[4]-shaped rows: [user id | 3 user-side one-hots | item id | 3 item-side one-hots], m = 8,
    n = 2,000,000 + 60 columns. Sorted columns per row, float64 values, vectorised."""
    from scipy.sparse import csr_matrix
    from rfm_b200.synth import SyntheticLog
    rng = np.random.default_rng(seed)
    U = I = 1_000_000
    groups = (4, 8, 12, 6, 10, 20)
    t0 = time.perf_counter()

    def rows_of(n_rows):
        u = rng.integers(0, U, n_rows)
        i = rng.integers(0, I, n_rows)
        cols = np.empty((n_rows, 8), dtype=np.int32)
        cols[:, 0] = u
        off = U
        for g_idx in range(3):                       # user-side one-hots are a function of the user
            cols[:, 1 + g_idx] = off + (u * (g_idx + 3) + g_idx) % groups[g_idx]
            off += groups[g_idx]
        cols[:, 4] = off + i
        off += I
        for g_idx in range(3, 6):
            cols[:, 1 + g_idx + 1] = off + (i * (g_idx + 2) + g_idx) % groups[g_idx]
            off += groups[g_idx]
        X = csr_matrix((np.ones(n_rows * 8), cols.ravel(), np.arange(0, 8 * n_rows + 1, 8, dtype=np.int64)
                        if n_rows * 8 >= 2**31 else np.arange(0, 8 * n_rows + 1, 8, dtype=np.int32)),
                       shape=(n_rows, off))
        X.has_sorted_indices = True
        y = (rng.random(n_rows) < 0.3).astype(np.int64)
        ps = rng.uniform(0.3, 1.0, n_rows)
        return {"features": X, "labels": y, "pscores": ps}, off

    train, n = rows_of(rows)
    val, _ = rows_of(N_VAL)
    log = SyntheticLog(U, I, n, train, val, None, None, {}, None, np.zeros((0, 2), np.int64), {})
    return log, time.perf_counter() - t0


def make_data(rows, seed, rank=0):
    if WORKLOAD == "stress":
        return make_stress_data(rows, seed)
    from rfm_b200.synth import make_kuairec_shaped
    t0 = time.perf_counter()
    log = make_kuairec_shaped(seed=seed + rank, n_users=N_USERS, n_items=N_ITEMS, n_train=rows, n_val=N_VAL,
                              build_mf=False, build_eval=False)
    return log, time.perf_counter() - t0


def algorithmic_bytes(X, batch_rows, k, s):
    """bytes one interaction must move, per pass of the hot path."""
    m = X.nnz / X.shape[0]
    sub = X[batch_rows]
    touched = np.unique(sub.indices).size
    B = len(batch_rows)
    stream = 4 + 4 + m * (4 + s) + 4 + s
    step = stream + 2 * m * (k + 1) * s + 2 * (touched / B) * (k + 1) * s
    per_kernel = {
        # row pass: batch stream + gather of V rows and w + write of s_t, e_t and the sort triples
        "fm_rows_train": 8 + 16 + m * (4 + s) + s + m * (k + 1) * s + k * s + s + m * (8 + s),
        # column pass: sorted triples + e_t and s_t gathers + read/write of each touched row
        "fm_cols": m * (8 + s) + m * s + m * k * s + 2 * (touched / B) * (k + 1) * s,
        # loss pass: batch stream + gather of V rows and w
        "fm_rows_loss": 8 + 16 + m * (4 + s) + s + m * (k + 1) * s,
    }
    return m, touched, step, per_kernel


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def count(self):
        """Samples written so far."""
        if self.proc is None:
            return 0
        try:
            with open(self.tmp.name) as f:
                return sum(1 for line in f if line.count(",") >= 8)
        except OSError:
            return 0

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---- the CPU arm -----------------------------------------------------------------------------------
def cpu_port_run(log, batch, steps, warmup, budget_s=25.0):
    """Oracle port on the host: reference sampler + step + post-update loss + val loss per epoch
    (oracle/fm_oracle.py::fm_fit == src/fm.py:55-112). Returns (interactions/s, description)."""
    from oracle import fm_oracle
    w0, w, V = fm_oracle.fm_init(12345, log.n_features, K_FACTORS)
    t0 = time.perf_counter()
    (w0, w, V), _, _ = fm_oracle.fm_fit(log.fm_train, log.fm_val, 1, batch, LR, w0, w, V)
    one = time.perf_counter() - t0
    if warmup + steps > 1 and one * (warmup + steps) > budget_s:
        steps = max(1, int(budget_s / one) - warmup)
    if warmup > 1:
        (w0, w, V), _, _ = fm_oracle.fm_fit(log.fm_train, log.fm_val, warmup - 1, batch, LR, w0, w, V, first_epoch=1)
    t0 = time.perf_counter()
    fm_oracle.fm_fit(log.fm_train, log.fm_val, steps, batch, LR, w0, w, V, first_epoch=warmup)
    dt = time.perf_counter() - t0
    return steps * batch / dt, steps, dt


# ---- the same port on every host core -----------------------------------------------------------------
# The gradient of a batch is a plain sum over its rows (oracle/fm_oracle.py::fm_grad), and so are the loss sums, so
# the port splits every epoch's batch over forked workers that share the CSR arrays copy-on-write and the parameters
# through shared memory; the legacy shuffles of the coming epochs (0.8 s each, sequential by nature) are computed
# ahead by the same pool, as the product's own host prefetcher does. Same arithmetic as the scalar port.
_PAR = {}


def _par_shuffle(epoch):
    from oracle import fm_oracle
    return fm_oracle.legacy_batch(_PAR["X"].shape[0], _PAR["B"], epoch)


def _par_grad(job):
    from oracle import fm_oracle
    wid, idx = job
    g0, a, G = fm_oracle.fm_grad(_PAR["X"][idx], _PAR["y"][idx], _PAR["ps"][idx], _PAR["w0"], _PAR["w"], _PAR["V"])
    out = _PAR["out"][wid]
    n = a.shape[0]
    out[0] = g0
    out[1:1 + n] = a
    out[1 + n:] = G.ravel()


def _par_loss(job):
    from oracle import fm_oracle
    which, idx = job
    X, y, ps = (_PAR["X"], _PAR["y"], _PAR["ps"]) if which == 0 else (_PAR["Xv"], _PAR["yv"], _PAR["psv"])
    if len(idx) == 0:
        return which, 0.0
    p = fm_oracle.fm_predict(X[idx], _PAR["w0"], _PAR["w"], _PAR["V"])
    return which, fm_oracle.ips_logloss(y[idx], p, ps[idx]) * len(idx)


def cpu_port_run_parallel(log, batch, steps, warmup, budget_s=25.0, workers=None):
    """(interactions/s, steps, seconds, workers) of the port on `workers` processes (default: every host core)."""
    import multiprocessing as mp
    from oracle import fm_oracle
    workers = workers or max(1, min(os.cpu_count() or 1, 64))
    n, k = log.n_features, K_FACTORS
    w0, w, V = fm_oracle.fm_init(12345, n, k)

    def shared(arr):
        raw = mp.RawArray("d", int(arr.size))
        view = np.frombuffer(raw, dtype=np.float64).reshape(arr.shape)
        view[...] = arr
        return view

    X = log.fm_train["features"].tocsr()
    Xv = log.fm_val["features"].tocsr()
    _PAR.update(X=X, y=np.asarray(log.fm_train["labels"], dtype=np.float64), ps=np.asarray(log.fm_train["pscores"]),
                Xv=Xv, yv=np.asarray(log.fm_val["labels"], dtype=np.float64), psv=np.asarray(log.fm_val["pscores"]),
                B=batch, w0=shared(np.asarray(w0, dtype=np.float64).reshape(1)), w=shared(w), V=shared(V),
                out=shared(np.zeros((workers, 1 + n + n * k))))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        def epoch(e, idx):
            pool.map(_par_grad, list(enumerate(np.array_split(idx, workers))))
            g = _PAR["out"].sum(axis=0)                       # partial sums in worker order
            _PAR["w0"][0] += LR * g[0]
            _PAR["w"] += LR * g[1:1 + n]
            _PAR["V"] += LR * g[1 + n:].reshape(n, k)
            jobs = [(0, part) for part in np.array_split(idx, workers)]
            jobs += [(1, part) for part in np.array_split(np.arange(Xv.shape[0]), workers)]
            sums = [0.0, 0.0]
            for which, v in pool.map(_par_loss, jobs):
                sums[which] += v
            return sums[0] / batch, sums[1] / Xv.shape[0]

        t0 = time.perf_counter()
        tl, vl = epoch(0, pool.apply(_par_shuffle, (0,)))
        one = time.perf_counter() - t0
        if warmup + steps > 1 and one * (warmup + steps) > budget_s:
            steps = max(1, int(budget_s / one) - warmup)
        for e in range(1, warmup):
            epoch(e, pool.apply(_par_shuffle, (e,)))
        t0 = time.perf_counter()
        ahead = [pool.apply_async(_par_shuffle, (warmup + e,)) for e in range(steps)]     # the reference's sampler
        for e in range(steps):
            tl, vl = epoch(warmup + e, ahead[e].get())
        dt = time.perf_counter() - t0
    assert np.isfinite(tl) and np.isfinite(vl)
    _PAR["last"] = (tl, vl, _PAR["w0"].copy(), _PAR["w"].copy(), _PAR["V"].copy())     # for tests/test_bench_host.py
    return steps * batch / dt, steps, dt, workers


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    log, gen_s = make_data(args.rows, 2024)
    value, steps, dt, cores = cpu_port_run_parallel(log, args.batch, args.steps, max(args.warmup, 1), budget_s=150.0)
    sample = ("%d epochs of B=%d (one GPU's share of the step) on the %d-row train set, reference sampler included, "
              "%.1f s on %d processes" % (steps, args.batch, args.rows, dt, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (NumPy/SciPy, single-threaded) and is not present on the GPU box; "
                "this is its CPU restatement oracle/fm_oracle.py (~8x faster per core than the reference's own "
                "per-factor loop, BASELINE.md section 4) run on every host core: each epoch's batch is split over "
                "forked workers (the gradient and the losses are sums over rows)",
    }
    print(json.dumps(line))


# ---- full-catalog scoring (second half of BASELINE.json's metric) ------------------------------------
def measure_scoring(device, dist, world, peaks, peak_kind):
    """Scored user-item pairs/s of rank-all-items-for-all-users: bf16 tcgen05 GEMM prune + exact float64
    top-K (csrc/score.cu). Two shapes: BASELINE configs[3]'s evaluation grid (1,411 x 3,327, k=64, top-9;
    latency-bound: 0.6 GFLOP) and a grid large enough for the tensor pipe to matter. With N GPUs the
    catalog is item-sharded and the per-rank lists are all-gathered and merged."""
    from rfm_b200.score import TopKScorer
    out = {"metric": "scored_user_item_pairs_per_sec", "unit": "pairs/s"}
    rng = np.random.default_rng(11)
    shapes = (("eval_grid_1411x3327", 1411, 3327, 64, 9, 20), ("large_32768x262144", 32768, 262144, 64, 9, 5),
              ("large_k128_32768x262144", 32768, 262144, 128, 9, 5),
              ("large_k128_top100_16384x131072", 16384, 131072, 128, 100, 5))
    for name, U, I, k, K, reps in shapes:
        A = rng.normal(size=(U, k)) * 0.3
        C = rng.normal(size=(I, k)) * 0.3
        beta = rng.normal(size=I) * 0.2
        sc = TopKScorer(A, C, None, beta, 0.0, device=device)
        ctx = sc.ctx

        def call():
            if dist is not None:
                from rfm_b200 import dist as rdist
                return rdist.sharded_topk(sc, dist, K)
            return sc.topk(K, copy=False)      # views of the library's page-locked result buffers

        for _ in range(3):
            call()
        if dist is not None:
            dist.barrier()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            dt = dist.max_over_ranks(dt)
        ctx.profile_begin()
        call()
        prof = ctx.profile_end()
        fk = sum(v[0] * v[1] for k2, v in prof.items() if k2 in ("score_sample", "score_collect"))
        i_local = I // world
        upad, ipad = -(-U // 128) * 128, -(-i_local // 256) * 256
        entry = {"users": U, "items": I, "n_factors": k, "top_k": K, "value": U * I / dt, "ms_per_call": dt * 1e3,
                 "includes": "operands resident; per call: sampled threshold pass, collect pass, exact re-score, "
                             "D2H of (items, scores)"
                             + ("; all-gather + merge across ranks" if world > 1 else ""),
                 "users_ranked_exactly": sc.last_stats.get("users_ranked_exactly"),
                 "candidates_per_user": round(sc.last_stats.get("candidates", 0) / U, 2),
                 "sample_stride": sc.last_stats.get("sample_stride"),
                 "kernels_ms": {k2: [v[0], round(v[1], 4)] for k2, v in prof.items()}}
        if fk > 0:
            tf = 2.0 * upad * ipad * 64 * -(-k // 64) / (fk * 1e-3) / 1e12     # one pass over the grid is algorithmic
            entry["roofline"] = {"bound": "tensor", "kernel": "score_pass_kernel (sample + collect)", "achieved": tf,
                                 "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": tf / peaks["bf16_tflops_sustained"], "traffic": None,
                                 "peak_kind": peak_kind + " (cuBLAS bf16, sustained)",
                                 "note": "per rank; algorithmic flops = 2 U I k (one pass) over the time of both "
                                         "tensor passes (the sampled threshold pass is overhead, not counted as work)"}
        out[name] = entry
        sc.close()
    return out


# ---- our arm ----------------------------------------------------------------------------------------
def run_ours(args):
    from ctypes import byref
    from rfm_b200 import _capi
    from rfm_b200._capi import check, lib, ptr
    from rfm_b200.fm import FactorizationMachines, _FmTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d ...)"
                             % (args.gpus, args.gpus))
    dist = None
    if world > 1:
        from rfm_b200 import dist as rdist
        dist = rdist.init(local_rank)

    # data-parallel runs replicate the train set on every GPU (2.6 GB of 180 GB): same seed everywhere
    log, gen_s = make_data(args.rows, 2024)
    X = log.fm_train["features"]
    pinned = []
    host_arrays = (X.indptr, X.indices, X.data, log.fm_train["labels"], log.fm_train["pscores"])
    for a in host_arrays:
        if _capi.pin_array(a):
            pinned.append(a)
    log.pinned_note = "%d of %d train arrays page-locked (cudaHostRegister)" % (len(pinned), len(host_arrays))
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    s = 8 if args.dtype == "float64" else 4
    dtype_tag = "f64" if args.dtype == "float64" else "f32"

    model = FactorizationMachines("IPS", K, K_FACTORS, LR, B, 12345, log.n_features, dtype=args.dtype,
                                  sampler="feistel", device=local_rank)
    ctx = model._context()
    train_rows = model._rows(X, log.fm_train["labels"], log.fm_train["pscores"])
    val_rows = model._rows(log.fm_val["features"], log.fm_val["labels"], log.fm_val["pscores"])
    model.sync_to_device()
    trainer = _FmTrainer(model._dev, train_rows, val_rows, B, W + 2 * K + 8)

    if dist is not None:
        from rfm_b200 import dist as rdist
        model.batch_size = B * world                       # weak scaling: B per GPU, global batch B*world
        dp = rdist.make_fm_dp(model, trainer, dist, B * world, N_VAL, LR, lambda epoch: None)

        def stepper(epoch, slot):
            dp.step(epoch)
    else:
        def stepper(epoch, slot):
            check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, epoch, B, LR, slot))

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    clocks = ClockSampler(local_rank)                      # nvidia-smi -lms 100 from before the warm-up on
    for e in range(W):
        stepper(e, e)
    barrier()
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for e in range(K):
        stepper(W + e, W + e)
    ms = ctx.timer_stop_ms()
    barrier()
    launches = ctx.launch_count() - launches0
    # The timed region lasts ~K x 0.25 ms, shorter than nvidia-smi's sampling period, so the same step loop keeps
    # running (untimed, same stream, same data) until the sampler has seen the GPU under this load a few times.
    seen0, t_obs, extra = clocks.count(), time.perf_counter(), 0
    n_obs = 1 if dist is None else 0        # ranks must issue the same number of collective steps: fixed count there
    while (dist is None and clocks.count() < seen0 + 3 and time.perf_counter() - t_obs < 3.0) or (dist is not None and n_obs < 1):
        for e in range(200 if dist is None else 1200):
            stepper(W + K + extra, W + 2 * K + 7)
            extra += 1
        ctx.synchronize()
        n_obs += 1
    clk = clocks.stop()
    clk["note"] = ("nvidia-smi -lms 100 from the warm-up to %d identical untimed steps right after the timed region "
                   "(the timed region itself is shorter than one sampling period)" % extra)
    if dist is not None:
        ms = dist.max_over_ranks(ms)
        launches = int(dist.sum_over_ranks(launches))
    value = K * B * world / (ms * 1e-3)

    # per-kernel share of the step: CUDA events around every launch, separate pass of the same steps
    ctx.profile_begin()
    for e in range(K):
        stepper(W + K + e, W + K + e)
    prof = ctx.profile_end()
    tl = np.empty(W + 2 * K)
    vl = np.empty(W + 2 * K)
    if dist is None:
        check(lib().rfm_fm_trainer_losses(trainer.handle, 0, W + 2 * K, ptr(tl), ptr(vl)))
    else:
        last = dp.flush().cpu().numpy()
        tl[:] = last[0] / (B * world)
        vl[:] = last[1] / N_VAL
    assert np.all(np.isfinite(tl)) and np.all(np.isfinite(vl)), "non-finite loss in the timed region"

    if rank != 0:
        if not args.no_e2e:
            measure_e2e(args, log, local_rank, dist, world)
        if not args.no_scoring:
            measure_scoring(local_rank, dist, world, *measured_peaks())
        dist.shutdown()
        return

    sample_rows = _capi.feistel_batch(X.shape[0], B, W, 12345)
    m, touched, step_bytes, per_kernel = algorithmic_bytes(X, sample_rows, K_FACTORS, s)
    total_prof_ms = sum(v[1] for v in prof.values())
    top = max((k for k in prof if k in per_kernel), key=lambda k: prof[k][1])
    top_ms = prof[top][1] / prof[top][0]
    peaks, peak_kind = measured_peaks()
    achieved = per_kernel[top] * B / (top_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and B == (65536 if WORKLOAD == "kuairec_big" else 1 << 20) and args.dtype == "float64":
        with open(tpath) as f:     # DRAM bytes per launch from the committed ncu --set full capture of this workload
            traffic = json.load(f).get(WORKLOAD, {}).get(top, {}).get("dram_bytes")
    roofline = {
        "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "algorithmic_bytes_per_launch": per_kernel[top] * B,
        "peak_kind": peak_kind + " (HBM copy, burst)",
        "algorithmic_bytes_per_interaction": per_kernel[top], "avg_launch_ms": top_ms,
        "share_of_step": prof[top][1] / total_prof_ms,
        "timing": "cudaEvent pair around each launch, separate pass of the same %d steps" % K,
        "note": ("V (%.1f MB) and S (%.1f MB) are L2-resident at this shape, so the algorithmic gather traffic is "
                 "served by L2, not HBM (SURVEY.md H7); achieved may therefore exceed the HBM peak"
                 % (log.n_features * K_FACTORS * s / 1e6, B * K_FACTORS * s / 1e6)) if WORKLOAD != "stress" else
                ("V (%.1f GB) and S (%.1f GB) exceed L2: id-column gathers and the s_t gather are HBM traffic"
                 % (log.n_features * K_FACTORS * s / 1e9, B * K_FACTORS * s / 1e9)),
        "step": {"algorithmic_bytes_per_interaction": step_bytes,
                 "achieved": step_bytes * B * world / (ms / K * 1e-3) / 1e9 / world,
                 "frac": step_bytes * B / (ms / K * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "kernels_ms_per_step": {k: round(v[1] / K, 5) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
    }

    e2e = None if args.no_e2e else measure_e2e(args, log, local_rank, dist, world)
    scoring = None if args.no_scoring else measure_scoring(local_rank, dist, world, peaks, peak_kind)
    cpu = None
    if not args.no_cpu_baseline:
        v1, st1, dt1 = cpu_port_run(log, B, 8, 1, budget_s=8.0)
        v, st, dt, cores = cpu_port_run_parallel(log, B, 160, 1, budget_s=15.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "host_cores_available": os.cpu_count(),
               "sample": "%d epochs of B=%d on the %d-row train set, reference sampler included, %.1f s on %d "
                         "processes" % (st, B, X.shape[0], dt, cores),
               "single_core": {"value": v1, "sample": "%d epochs, %.1f s" % (st1, dt1)}}
    cfg = workload_config(args, world)
    cfg["l2"] = cfg["l2"].replace("train CSR", "%.1f GB train CSR" % (train_rows.h2d_bytes / 1e9))
    if WORKLOAD == "stress":
        cfg["l2"] = ("inputs larger than L2: %.1f GB CSR, %.1f GB parameter table and %.1f GB of per-batch s_t rows, "
                     "all far beyond the 126 MB L2" % (train_rows.h2d_bytes / 1e9,
                                                        log.n_features * K_FACTORS * s / 1e9, B * K_FACTORS * s / 1e9))
    cfg.update(sampler="feistel (device, perf mode)", mean_nnz_per_row=round(m, 3),
               touched_columns_per_step=int(touched), n_features=log.n_features, data_gen_s=round(gen_s, 1))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_tag, "data": "synthetic", "config": cfg, "clocks": clk, "e2e": e2e,
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "scoring": scoring,
        "final_train_loss": float(tl[W + K - 1]), "final_val_loss": float(vl[W + K - 1]),
    }
    print(json.dumps(line))
    for a in pinned:
        _capi.unpin_array(a)
    if dist is not None:
        dist.shutdown()


def measure_e2e(args, log, device, dist, world):
    """Public API on host arrays: FactorizationMachines.fit(train, val) for K epochs. Everything a
    user pays is inside the timed region: CSR upload from pinned host memory, trainer set-up, K
    epochs, loss read-back, parameter download."""
    from rfm_b200.fm import FactorizationMachines
    B, K = args.batch, args.steps
    out = {}
    for sampler in ("feistel", "legacy"):
        n_ep = K if sampler == "feistel" else min(K, 16)
        if dist is not None and sampler == "legacy":
            continue
        warm = FactorizationMachines("IPS", 2, K_FACTORS, LR, B * world, 12345, log.n_features, dtype=args.dtype,
                                     sampler=sampler, device=device, distributed=dist)
        warm.fit(log.fm_train, log.fm_val)
        del warm
        model = FactorizationMachines("IPS", n_ep, K_FACTORS, LR, B * world, 12345, log.n_features,
                                      dtype=args.dtype, sampler=sampler, device=device, distributed=dist)
        if dist is not None:
            dist.barrier()
        model._context().synchronize()
        t0 = time.perf_counter()
        tl, vl = model.fit(log.fm_train, log.fm_val)
        model._context().synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            dt = dist.max_over_ranks(dt)
        rows_bytes = model.last_fit_stats["h2d_bytes_rows"]
        upload_s = model.last_fit_stats.get("upload_seconds")
        out[sampler] = {
            "value": n_ep * B * world / dt, "epochs": n_ep, "seconds": dt, "upload_seconds": upload_s,
            "phase_seconds": model.last_fit_stats.get("phase_seconds"),
            "h2d_bytes_per_step": rows_bytes / n_ep + (B * 8 if sampler == "legacy" else 0),
            "d2h_bytes_per_step": 16 + (1 + log.n_features * (K_FACTORS + 1)) * 8 / n_ep,
        }
    main = out["feistel"]
    res = {"value": main["value"], "unit": UNIT, "h2d_bytes_per_step": main["h2d_bytes_per_step"],
           "d2h_bytes_per_step": main["d2h_bytes_per_step"], "seconds": main["seconds"], "epochs": main["epochs"],
           "upload_seconds": main["upload_seconds"], "host_memory": getattr(log, "pinned_note", None),
           "phase_seconds": main["phase_seconds"],
           "api": "FactorizationMachines(sampler='feistel').fit(train, val) on pinned host arrays; includes the "
                  "one-time CSR upload, amortised over the epochs of this call"}
    if "legacy" in out:
        res["legacy_sampler"] = dict(out["legacy"], note="reference batch order (RandomState(epoch) shuffle of all "
                                     "N ids on host threads, SURVEY.md F14): host-bound by design")
    return res


if __name__ == "__main__":
    a = parse()
    if a.workload == "stress":
        WORKLOAD, N_USERS, N_ITEMS, K_FACTORS = "stress", 1_000_000, 1_000_000, 128
        if a.rows == N_TRAIN:
            a.rows = 20_000_000
        if a.batch == 65536:
            a.batch = 1 << 20
        a.no_cpu_baseline = a.no_scoring = True
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
