#!/usr/bin/env python
"""bench.py -- train interactions/s of the fused IPS-FM epoch on synthetic KuaiRec-big-shaped data, and the rest of
BASELINE.json's metric (scored user-item pairs/s) next to it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one reference epoch (src/fm.py:71-102): one minibatch of B interactions through forward, IPS residual,
the simultaneous w0/w/V update, the post-update batch loss and the val loss.

ours:
  value        K*B*N / device time of K steps (CUDA events), rows resident in HBM in the factored form (SURVEY 8 f3;
               config.input_format), batches drawn on the device (Feistel sampler, perf mode), float64.
  e2e          the same metric through the public API on HOST (pinned) arrays, FactorizationMachines.fit(train, val):
               row upload, K epochs, loss read-back, parameter download, wall clock around the call. Headline:
               factored input + device sampler; e2e.hstacked_csr = the reference's own stacked-CSR input;
               e2e.legacy_sampler = the reference's batch order (RandomState(epoch) shuffle, host-bound by design);
               e2e.scoring / e2e.mf = the same for full-catalog ranking and for MF.
  roofline     frac = SURVEY 8(d) whole-step algorithmic bytes / step time / measured HBM copy peak, l2_assisted says
               whether the gathers are served by L2 at this shape (they are at configs[2]); roofline.stress = the same
               fraction on the configs[4]-shaped stress rows where nothing fits L2 (the HBM claim);
               roofline.scoring = tensor-core fractions of the scoring passes; per-kernel CUDA-event times alongside.
  cpu_baseline the CPU oracle (NumPy/SciPy port of the reference step + the reference's own sampler) on a bounded
               sample of the same workload, on every host core; cpu_baseline.scoring / .mf likewise.
  clocks       nvidia-smi samples from the warm-up to a run of identical untimed steps.
reference:     the CPU oracle port on every host core timed alone, same config / metric / unit (the reference is pure
               Python and cannot travel to the GPU box).
Under torchrun (N > 1): one rank per GPU, B per GPU (weak scaling), the library's NVLink exchange kernel between
ranks; before anything is timed the kuairec_small golden is fitted data-parallel with both samplers and checked
(dp_parity; a mismatch fails the run); time = max over ranks; scoring is item-sharded over the ranks.
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
    ap.add_argument("--input", default="factored", choices=["factored", "csr"],
                    help="row format RESIDENT IN HBM for the device-timed steps (value): factored = user table + item "
                         "table + (user, item, ctx) records (SURVEY 8 f3; what the e2e uploads), csr = the reference's "
                         "stacked matrix; the other formats / steps are timed next to it (roofline.other_input_formats)")
    ap.add_argument("--step", default="auto", choices=["auto", "flat", "two_level"],
                    help="factored rows: two_level = per-entity aggregates (csrc/two_level.cuh), flat = one gathered "
                         "parameter row per stored non-zero, auto = what FactorizationMachines.fit picks")
    ap.add_argument("--e2e-input", default="factored", choices=["factored", "csr"],
                    help="row format the headline e2e uploads from host memory; the other one is e2e.hstacked_csr / "
                         "e2e.factored")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
    ap.add_argument("--no-stress", action="store_true")
    ap.add_argument("--no-mf", action="store_true")
    ap.add_argument("--no-search", action="store_true", help="skip the 500-epoch search section (e2e.epoch_search)")
    ap.add_argument("--stress-rows", type=int, default=20_000_000)
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] section that a default 8-GPU run carries")
    ap.add_argument("--c5-rows", type=int, default=0, help="interactions of the c5 log (default: 125 M per GPU)")
    ap.add_argument("--c5-users", type=int, default=1_000_000)
    ap.add_argument("--c5-items", type=int, default=1_000_000)
    ap.add_argument("--workload", default="kuairec_big", choices=["kuairec_big", "stress", "c5"],
                    help="kuairec_big = BASELINE configs[2] (the headline; the stress shape runs as a section of it); "
                         "stress = only the configs[4]-shaped section: 1M users x 1M items, k=128, 8 non-zeros per row")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "IPS-FM, synthetic KuaiRec big_matrix shape (BASELINE.json configs[2])",
        "n_users": N_USERS, "n_items": N_ITEMS, "train_interactions": args.rows, "val_rows": N_VAL,
        "n_factors": K_FACTORS, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "lr": LR, "parallelism": "dp%d" % world if world > 1 else "single",
        "l2": "inputs larger than L2: each step gathers a fresh random batch from the resident "
              "train rows; the parameter table V is legitimately L2-resident across steps",
    }


STRESS_USERS = STRESS_ITEMS = 1_000_000
STRESS_K = 128
STRESS_GROUPS = (4, 8, 12, 6, 10, 20)


def make_stress_data(rows, seed):
    """configs[4]-shaped rows, [user id | 3 user-side one-hots | item id | 3 item-side one-hots] (m = 8,
    n = 2,000,000 + 60 columns), in the factored form: two 1 M-row side tables and one (user, item) pair per row."""
    from scipy.sparse import csr_matrix
    from rfm_b200.factored import FactoredFeatures
    rng = np.random.default_rng(seed)
    U, I = STRESS_USERS, STRESS_ITEMS
    t0 = time.perf_counter()

    def side(n, groups, mult0):
        cols = np.empty((n, 3), dtype=np.int32)
        ids = np.arange(n, dtype=np.int64)
        off = 0
        for j, g in enumerate(groups):
            cols[:, j] = off + (ids * (j + mult0) + j) % g
            off += g
        return csr_matrix((np.ones(n * 3), cols.ravel(), np.arange(0, 3 * n + 1, 3, dtype=np.int32)), shape=(n, off))

    ut, it = side(U, STRESS_GROUPS[:3], 3), side(I, STRESS_GROUPS[3:], 5)
    blocks = [("id", "user", U), ("table", "user", ut), ("id", "item", I), ("table", "item", it)]

    def rows_of(n_rows):
        u = rng.integers(0, U, n_rows).astype(np.int32)
        i = rng.integers(0, I, n_rows).astype(np.int32)
        y = (rng.random(n_rows) < 0.3).astype(np.int8)
        ps = rng.uniform(0.3, 1.0, n_rows)
        return {"features": FactoredFeatures(blocks, u, i), "labels": y, "pscores": ps}

    train, val = rows_of(rows), rows_of(N_VAL)
    return train, val, train["features"].shape[1], time.perf_counter() - t0


def make_data(rows, seed, rank=0):
    from rfm_b200.synth import make_kuairec_shaped
    t0 = time.perf_counter()
    log = make_kuairec_shaped(seed=seed + rank, n_users=N_USERS, n_items=N_ITEMS, n_train=rows, n_val=N_VAL,
                              build_mf=False, build_eval=False)
    return log, time.perf_counter() - t0


def factored_dicts(log):
    """The train / val dicts with the rows in the factored form (what the reference's preparer holds before hstack),
    in the compact form the format allows: int32 ids, float64 context, int8 labels and the propensities as the
    per-ITEM table they are gathered from = 17 bytes per interaction (0.20 GB at 12 M rows, against 2.6 GB of stacked
    CSR + 0.19 GB of labels and per-row pscores)."""
    from rfm_b200.factored import PerItem
    from rfm_b200.synth import factored_from_tables
    out = []
    for d in (log.fm_train, log.fm_val):
        assert np.array_equal(log.tables["item_pscore"][d["items"]], d["pscores"])
        out.append({"features": factored_from_tables(log.tables, d["users"].astype(np.int32), d["items"].astype(np.int32),
                                                     d["ctx"]),
                    "labels": d["labels"].astype(np.int8), "pscores": PerItem(log.tables["item_pscore"])})
    return out


def pin_host_arrays(arrays):
    from rfm_b200 import _capi
    return [a for a in arrays if isinstance(a, np.ndarray) and a.nbytes >= (1 << 20) and _capi.pin_array(a)]


def algorithmic_bytes(X, batch_rows, k, s):
    """bytes one interaction must move, per pass of the hot path."""
    m = X.nnz / X.shape[0]
    sub = X[batch_rows]
    touched = np.unique(sub.indices).size
    return (m, touched) + algorithmic_bytes_of(m, touched, len(batch_rows), k, s)


def algorithmic_bytes_of(m, touched, B, k, s):
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
    return step, per_kernel


def two_level_bytes(B, k, s, n_ctx, entity_entries, n_features):
    """Bytes one interaction must move in the two-level step, counted the way section 8(d) counts the flat one: the
    row record, the forward gathers of both passes (user row, item row, context columns), and per STEP the entity
    pass (entity_entries parameter rows gathered forward, as many rows of per-entity sums gathered backward, every
    parameter row read and written once), spread over the B interactions of the step."""
    stream = 4 + 4 + 4 + n_ctx * s + 4 + s
    return stream + 2 * (2 + n_ctx) * (k + 1) * s + (2 * entity_entries + 2 * n_features) * (k + 1) * s / B


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
SCORING_SHAPES = (("eval_grid_1411x3327", 1411, 3327, 64, 9, 20), ("large_32768x262144", 32768, 262144, 64, 9, 5),
                  ("large_k128_32768x262144", 32768, 262144, 128, 9, 5),
                  ("large_k128_top100_16384x131072", 16384, 131072, 128, 100, 5))


def measure_scoring(device, dist, world, peaks, peak_kind):
    """Scored user-item pairs/s of rank-all-items-for-all-users: bf16 tcgen05 GEMM prune + exact float64
    top-K (csrc/score.cu). Shapes: BASELINE configs[3]'s evaluation grid (1,411 x 3,327, k=64, top-9;
    latency-bound: 0.6 GFLOP) and grids large enough for the tensor pipe to matter. With N GPUs the catalog is
    item-sharded: global thresholds and the user-partitioned K-way merge run over NVLink peer memory and every rank
    ends with the lists of the users it owns in host memory. The same call on one GPU is timed next to it
    (single_gpu_ms), so the line carries its own speed-up and scaling efficiency."""
    from rfm_b200.score import TopKScorer
    out = {"metric": "scored_user_item_pairs_per_sec", "unit": "pairs/s"}
    rng = np.random.default_rng(11)
    for name, U, I, k, K, reps in SCORING_SHAPES:
        A = rng.normal(size=(U, k)) * 0.3
        C = rng.normal(size=(I, k)) * 0.3
        beta = rng.normal(size=I) * 0.2
        sc = TopKScorer(A, C, None, beta, 0.0, device=device)
        ctx = sc.ctx

        def single():
            return sc.topk(K, copy=False)      # views of the library's page-locked result buffers

        def call():
            if dist is not None:
                from rfm_b200 import dist as rdist
                return rdist.sharded_topk(sc, dist, K, gather=False, copy=False)
            return single()

        def timed(fn, sync_ranks):
            for _ in range(3):
                fn()
            if sync_ranks:
                dist.barrier()
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            ctx.synchronize()
            dt = (time.perf_counter() - t0) / reps
            return dist.max_over_ranks(dt) if sync_ranks else dt

        single_dt = timed(single, False) if dist is not None else None
        dt = timed(call, dist is not None)
        ctx.profile_begin()
        call()
        prof = ctx.profile_end()
        fk = sum(v[1] for k2, v in prof.items() if k2 in ("score_sample", "score_collect"))       # v = (launches, total ms)
        i_local = I // world
        upad, ipad = -(-U // 128) * 128, -(-i_local // 256) * 256
        entry = {"users": U, "items": I, "n_factors": k, "top_k": K, "value": U * I / dt, "ms_per_call": dt * 1e3,
                 "includes": "operands resident; per call: sampled threshold pass, collect pass, exact re-score, "
                             "D2H of (items, scores)"
                             + ("; item-sharded: threshold exchange + user-partitioned K-way merge over NVLink peer "
                                "memory, every rank reads back the users it owns" if world > 1 else ""),
                 "users_ranked_exactly": sc.last_stats.get("users_ranked_exactly"),
                 "candidates_per_user": round(sc.last_stats.get("candidates", 0) / U, 2),
                 "sample_stride": sc.last_stats.get("sample_stride"),
                 "kernels_ms": {k2: [v[0], round(v[1], 4)] for k2, v in prof.items()}}
        if single_dt is not None:
            entry.update(single_gpu_ms=single_dt * 1e3, speedup_vs_1gpu=single_dt / dt,
                         scoring_efficiency=single_dt / dt / world)
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


def eval_grid_inputs(log_small=None):
    """BASELINE configs[3]'s evaluation grid: 1,411 users x 3,327 items of the KuaiRec-shaped catalog, an FM at its
    reference initialisation, ~46 held-out rows per user (labels), item exposures."""
    from rfm_b200.synth import make_kuairec_shaped
    from oracle import fm_oracle
    log = make_kuairec_shaped(seed=2024, n_users=N_USERS, n_items=N_ITEMS, n_train=50_000, n_val=100)
    w0, w, V = fm_oracle.fm_init(12345, log.n_features, K_FACTORS, alpha=0.1)
    frame = log.test_frame
    users = np.unique(frame["user"])
    items = np.unique(frame["item"])
    return log, (w0, w, V), frame, users, items


def scoring_cpu_and_e2e(device, budget_s=12.0):
    """The reference's way of ranking a full grid -- FactorizationMachines.predict on the Cartesian-product rows
    (src/fm.py:114-133) + TestEvaluator.evaluate (utils/evaluate.py:80-127), restated by the oracle -- timed on a
    bounded sample of users; and this build's end-to-end path on the same grid from HOST arrays: per-side FM
    factors -> upload + bf16 conversion -> tensor-core top-K -> device metric reductions -> metrics on the host."""
    from oracle import fm_oracle, metrics_oracle
    from rfm_b200.evaluate import FullCatalogEvaluator
    from rfm_b200.score import TopKScorer, fm_side
    from rfm_b200.synth import csr_from_tables
    log, (w0, w, V), frame, users, items = eval_grid_inputs()
    U, I = users.size, items.size
    K = [1, 3, 5, 7, 9]
    used = {"DCG", "CatalogCoverage"}
    u_of = {int(u): j for j, u in enumerate(users)}
    i_of = {int(i): j for j, i in enumerate(items)}
    lab_u = np.array([u_of[int(u)] for u in frame["user"]])
    lab_i = np.array([i_of[int(i)] for i in frame["item"]])
    theta = np.zeros(I)
    theta[lab_i] = frame["pscore"]
    theta[theta == 0] = 0.5
    # ---- CPU: Cartesian rows of a sample of users, oracle predict + oracle TestEvaluator
    n_sample = 120
    t0 = time.perf_counter()
    su = np.arange(n_sample)
    uu = np.repeat(users[su], I)
    ii = np.tile(items, n_sample)
    X = csr_from_tables(uu, ii, log.tables, np.zeros(uu.size))
    X.has_sorted_indices = False
    X.sort_indices()
    build_s = time.perf_counter() - t0
    lab = np.zeros((n_sample, I), dtype=np.int64)
    sel = lab_u < n_sample
    lab[lab_u[sel], lab_i[sel]] = frame["label"][sel]
    cart = {"user": np.repeat(su, I), "item": np.tile(np.arange(I), n_sample), "label": lab.ravel(),
            "pscore": np.tile(theta, n_sample), "ones_pscore": np.ones(n_sample * I)}
    t0 = time.perf_counter()
    scores = fm_oracle.fm_predict(X, w0, w, V)
    t_pred = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = metrics_oracle.test_evaluate(cart, scores, K, used, I)
    t_eval = time.perf_counter() - t0
    cpu = {"value": n_sample * I / (t_pred + t_eval), "unit": "pairs/s", "cores": 1, "kind": "port",
           "sample": "%d of %d users x %d items: oracle fm_predict on the Cartesian CSR (%.2f s) + oracle TestEvaluator "
                     "(%.2f s); building the Cartesian rows (%.2f s) not counted" % (n_sample, U, I, t_pred, t_eval, build_s)}
    # ---- ours, end to end from host arrays, whole grid
    # per-side tables: the user side holds the user's id and side features (+ the context column at its fixed value),
    # the item side the item's id and side features; both as rows over the global columns
    from scipy.sparse import csr_matrix
    t = log.tables

    def side_rows(ptr, col, val, ids):
        lens = (ptr[ids + 1] - ptr[ids]).astype(np.int64)
        indptr = np.zeros(ids.size + 1, dtype=np.int64)
        np.cumsum(lens, out=indptr[1:])
        take = np.concatenate([np.arange(ptr[e], ptr[e + 1]) for e in ids])
        return csr_matrix((val[take], col[take], indptr), shape=(ids.size, t["n_features"]))

    ut = side_rows(t["u_ptr"], t["u_col"], t["u_val"], users)
    it = side_rows(t["i_ptr"], t["i_col"], t["i_val"], items)
    ev = FullCatalogEvaluator({"user": lab_u, "item": lab_i, "label": frame["label"]}, theta, K, used, U, I)

    scorer = []

    def ours():
        A, alpha = fm_side(ut, w, V)
        C, beta = fm_side(it, w, V)
        if not scorer:                       # device buffers, tensor maps, staging areas: once per (users, items, k)
            scorer.append(TopKScorer(A, C, alpha, beta, float(w0[0]), device=device))
        else:
            scorer[0].update(A, C, alpha, beta, float(w0[0]))
        return ev.evaluate(scorer[0])

    res = ours()
    # parity of the sampled users' share is not comparable (coverage is global); DCG of the sample is checked in tests
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        res = ours()
    dt = (time.perf_counter() - t0) / reps
    e2e = {"value": U * I / dt, "unit": "pairs/s", "ms_per_call": dt * 1e3, "grid": "%d x %d, k=%d, K=%s" % (U, I, K_FACTORS, K),
           "h2d_bytes_per_step": (U + I) * (K_FACTORS + 1) * 8, "d2h_bytes_per_step": len(K) * 12 * 8 + len(K) * I * 4,
           "api": "fm_side (host, per-side FM factors) -> TopKScorer.update (upload, bf16 operands) -> "
                  "FullCatalogEvaluator.evaluate (tcgen05 top-K, device label look-up + DCG/ME/coverage reductions) "
                  "-> metrics dict on the host",
           "dcg_at_9": float(res["DCG"][-1]), "coverage_at_9": float(res["CatalogCoverage"][-1])}
    scorer[0].close()
    return cpu, e2e


# ---- epoch search (utils/search_params.py:79-152, src/fm.py:104-110) ------------------------------------------------
def measure_epoch_search(device, with_cpu, budget_s=10.0):
    """The reference's real workflow: 500 epochs of B = 2,000 at its tuned k = 400 on the KuaiRec small_matrix shape
    (1,411 users x 3,327 items), with the val DCG@5 of ~65 k held-out rows computed after EVERY epoch
    (logs/kuairec/main_kuairec.log:14-15: 657 s on the author's CPU). Here: one
    FactorizationMachines(evaluator=ValEvaluator).fit on host arrays; train step, both losses, predict on the eval rows,
    per-user ranking and the metric all stay on the device, one read-back at the end."""
    from rfm_b200.evaluate import ValEvaluator
    from rfm_b200.fm import FactorizationMachines
    from rfm_b200.synth import make_kuairec_shaped
    n_users, n_items, n_train, B, k, epochs = 1411, 3327, 1_000_000, 2000, 400, 500
    log = make_kuairec_shaped(seed=2024, n_users=n_users, n_items=n_items, n_train=n_train, n_val=N_VAL,
                              eval_users=n_users, eval_items=n_items, eval_rows_per_user=46, build_mf=False)
    out = {}
    for sampler in ("feistel", "legacy"):
        def run(n_ep):
            ev = ValEvaluator(interaction_df=log.test_frame, features={"FM": log.fm_test_features}, k=5, metric_name="DCG")
            m = FactorizationMachines("IPS", n_ep, k, LR, B, 12345, log.n_features, evaluator=ev, sampler=sampler,
                                      device=device)
            m._context().synchronize()
            t0 = time.perf_counter()
            tl, vl = m.fit(log.fm_train, log.fm_val)
            return time.perf_counter() - t0, tl, m
        run(3)
        dt, tl, m = run(epochs)
        assert np.all(np.isfinite(tl)) and len(m.val_metrics) == epochs and np.all(np.isfinite(m.val_metrics))
        out[sampler] = {"seconds": dt, "seconds_per_epoch": dt / epochs, "value": epochs * B / dt, "unit": UNIT,
                        "final_val_dcg_at_5": float(m.val_metrics[-1]), "gpu_launches": m.last_fit_stats.get("gpu_launches")}
    res = dict(out["legacy"], epochs=epochs, batch=B, n_factors=k, eval_rows=int(log.fm_test_features.shape[0]),
               train_rows=n_train, sampler="legacy (the reference's batch order)", device_sampler=out["feistel"],
               reference_log={"seconds": 657.0, "seconds_per_epoch": 1.31, "hardware": "author's CPU (unknown), real KuaiRec",
                              "source": "logs/kuairec/main_kuairec.log:14-15 (BASELINE.md section 3)"},
               api="FactorizationMachines(evaluator=ValEvaluator(k=5, 'DCG')).fit(train, val): 500 x (IPS-FM step, batch "
                   "+ val loss, predict on the eval rows, per-user ranking, IPS-DCG@5), stacked CSR on host arrays")
    cpu = None
    if with_cpu:
        from oracle import fm_oracle, metrics_oracle
        mm = FactorizationMachines("IPS", 1, k, LR, B, 12345, log.n_features)
        w0, w, V = mm.w0().copy(), mm.w().copy(), mm.V().copy()
        n_ep, t0 = 0, time.perf_counter()
        while n_ep < 2 or (time.perf_counter() - t0 < budget_s and n_ep < 20):
            (w0, w, V), _, _ = fm_oracle.fm_fit(log.fm_train, log.fm_val, 1, B, LR, w0, w, V, first_epoch=n_ep)
            metrics_oracle.val_evaluate(log.test_frame, fm_oracle.fm_predict(log.fm_test_features, w0, w, V), 5, "IPS")
            n_ep += 1
        dtc = time.perf_counter() - t0
        cpu = {"seconds_per_epoch": dtc / n_ep, "value": n_ep * B / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d epochs (oracle fm_fit + fm_predict on the eval rows + val_evaluate), %.1f s; 500 epochs "
                         "would take %.0f s" % (n_ep, dtc, 500 * dtc / n_ep)}
    return res, cpu


# ---- MF (src/mf.py:68-134) -----------------------------------------------------------------------------
def measure_mf(device, with_cpu, budget_s=8.0):
    """IPS-MF at the KuaiRec big_matrix shape (7,176 x 10,728, 2 M train pairs, k = 64, B = 65,536): the reference's
    strictly sequential per-sample SGD executed as a wavefront schedule (csrc/mf.cu), reference batch order.
    End to end through LogisticMatrixFactorization.fit (upload, epochs, both losses, parameter download)."""
    from rfm_b200.mf import LogisticMatrixFactorization
    from rfm_b200.synth import make_kuairec_shaped
    log = make_kuairec_shaped(seed=2024, n_users=N_USERS, n_items=N_ITEMS, n_train=2_000_000, n_val=N_VAL,
                              build_eval=False)
    B, epochs, k, lr, reg = 65536, 10, K_FACTORS, 9e-6, 1e-4
    LogisticMatrixFactorization("IPS", 2, k, lr, B, 12345, log.n_users, log.n_items, reg, device=device).fit(
        log.mf_train, log.mf_val)
    m = LogisticMatrixFactorization("IPS", epochs, k, lr, B, 12345, log.n_users, log.n_items, reg, device=device)
    t0 = time.perf_counter()
    tl, vl = m.fit(log.mf_train, log.mf_val)
    dt = time.perf_counter() - t0
    assert np.all(np.isfinite(tl)) and np.all(np.isfinite(vl))
    out = {"value": epochs * B / dt, "unit": UNIT, "epochs": epochs, "batch": B, "seconds": dt, "n_factors": k,
           "train_pairs": int(log.mf_train["features"].shape[0]), "sampler": "legacy (reference batch order)",
           "gpu_launches": m.last_fit_stats.get("gpu_launches"),
           "h2d_bytes_per_step": log.mf_train["features"].nbytes * 2 / epochs + B * 8,
           "d2h_bytes_per_step": 16 + (log.n_users + log.n_items) * (k + 1) * 8 / epochs,
           "api": "LogisticMatrixFactorization.fit(train, val) on host arrays (replicas only: MF does not shard)"}
    cpu = None
    if with_cpu:
        from oracle import mf_oracle
        P, Q, bu, bi = mf_oracle.mf_init(12345, log.n_users, log.n_items, k)
        t0 = time.perf_counter()
        mf_oracle.mf_fit(log.mf_train, log.mf_val, 1, B, lr, reg, P, Q, bu, bi)
        one = time.perf_counter() - t0
        n_ep = max(1, min(4, int(budget_s / max(one, 1e-3))))
        t0 = time.perf_counter()
        mf_oracle.mf_fit(log.mf_train, log.mf_val, n_ep, B, lr, reg, P, Q, bu, bi)
        dtc = time.perf_counter() - t0
        cpu = {"value": n_ep * B / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d epochs of B=%d, oracle/mf_oracle.py (C inner loop) incl. the reference sampler and both "
                         "losses, %.1f s" % (n_ep, B, dtc)}
    return out, cpu


# ---- data-parallel parity, where the driver can see it ---------------------------------------------------------------
def dp_parity(dist, device):
    """Before anything is timed on N > 1 GPUs: the kuairec_small golden (the unmodified reference's trajectory) fitted
    data-parallel over this process group with the reference sampler, and the device sampler against the single-GPU
    fit; every rank must hold the same bits afterwards. A mismatch fails the run."""
    from scipy.sparse import csr_matrix
    from rfm_b200.fm import FactorizationMachines
    g = np.load(os.path.join(ROOT, "tests", "golden", "kuairec_small_fm_ips.npz"))

    def csr(prefix):
        return csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]),
                          shape=tuple(int(v) for v in g[prefix + "_shape"]))

    train = {"features": csr("train"), "labels": g["train_labels"], "pscores": g["train_pscores"]}
    val = {"features": csr("val"), "labels": g["val_labels"], "pscores": g["val_pscores"]}
    kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=train["features"].shape[1],
              alpha=float(g["alpha"]), device=device)
    torch = dist.torch
    worst, identical = 0.0, True
    for sampler in ("legacy", "feistel"):
        m = FactorizationMachines(distributed=dist, sampler=sampler, **kw)
        tl, vl = m.fit(train, val)
        if sampler == "legacy":
            ref = (g["train_loss"], g["val_loss"], g["V"], g["w"])
        else:
            one = FactorizationMachines(sampler=sampler, **kw)
            rtl, rvl = one.fit(train, val)
            ref = (np.array(rtl), np.array(rvl), one.V(), one.w())
        for mine, want in zip((np.array(tl), np.array(vl), m.V(), m.w()), ref):
            err = np.max(np.abs(mine - want) / np.maximum(np.abs(want), 1e-13))
            worst = max(worst, float(err))
        mine = torch.from_numpy(m.V().copy()).to("cuda:%d" % device)
        lo, hi = mine.clone(), mine.clone()
        dist.dist.all_reduce(lo, op=dist.dist.ReduceOp.MIN)
        dist.dist.all_reduce(hi, op=dist.dist.ReduceOp.MAX)
        identical = identical and bool(torch.equal(lo, hi))
    worst = dist.max_over_ranks(worst)
    out = {"max_rel_err": worst, "ranks_bit_identical": identical, "tolerance": 1e-9, "world": dist.world,
           "what": "kuairec_small golden (unmodified reference) fitted data-parallel: legacy sampler vs the golden, "
                   "device sampler vs the single-GPU fit; losses, w, V"}
    if not (worst <= 1e-9 and identical):
        raise SystemExit("dp_parity FAILED: %s" % json.dumps(out))
    return out


# ---- our arm ----------------------------------------------------------------------------------------
def timed_steps(ctx, dist, stepper, W, K, device, observe=True):
    """W warm-up steps, K timed steps (CUDA events on the launching stream, barrier + synchronize on both sides, max
    over ranks), nvidia-smi clocks sampled around a run of identical untimed steps, then a profiled pass of K steps."""
    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    clocks = ClockSampler(device)                      # nvidia-smi -lms 100 from before the warm-up on
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
    if not observe:                         # long steps: the timed region itself spans several sampling periods
        n_obs = 1
    while observe and (dist is None and clocks.count() < seen0 + 3 and time.perf_counter() - t_obs < 3.0) or (observe and dist is not None and n_obs < 1):
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
    ctx.profile_begin()                                # per-kernel share of the step: CUDA events around every launch
    for e in range(K):
        stepper(W + K + e, W + K + e)
    prof = ctx.profile_end()
    return ms, launches, clk, prof


def step_roofline(prof, K, ms, B, world, m, touched, k, s, peaks, peak_kind, l2_assisted, note, traffic_key):
    step_bytes, per_kernel = algorithmic_bytes_of(m, touched, B, k, s)
    total_prof_ms = sum(v[1] for v in prof.values())
    if not any(kk in per_kernel for kk in prof):      # two-level step: its row passes gather 2 + n_ctx rows per interaction
        per_kernel = {"fm_vrows_train": 8 + 16 + 3 * (k + 1) * s + k * s + s + 3 * (8 + s),
                      "fm_vrows_loss": 8 + 16 + 3 * (k + 1) * s,
                      "fm_rows_train": 8 + 16 + 3 * (k + 1) * s + k * s + s + 3 * (8 + s),
                      "fm_rows_loss": 8 + 16 + 3 * (k + 1) * s}
    top = max((kk for kk in prof if kk in per_kernel), key=lambda kk: prof[kk][1])
    top_ms = prof[top][1] / prof[top][0]
    achieved = step_bytes * B / (ms / K * 1e-3) / 1e9
    traffic, kernel_traffic = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:     # DRAM bytes per launch from the committed ncu --set full captures of this workload
            tj = json.load(f).get(traffic_key, {})
        per = {kk: v.get("dram_bytes") for kk, v in tj.items() if isinstance(v, dict) and v.get("dram_bytes")}
        if per:
            kernel_traffic = per
            traffic = float(sum(v * max(1, round(prof[kk][0] / K)) if kk in prof else v for kk, v in per.items()))
    return {
        "bound": "hbm", "scope": "whole step: SURVEY.md section 8(d) algorithmic bytes per interaction x B / step time",
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": traffic, "traffic_per_kernel": kernel_traffic, "l2_assisted": l2_assisted,
        "algorithmic_bytes_per_interaction": step_bytes, "peak_kind": peak_kind + " (HBM copy, burst)", "note": note,
        "dominant_kernel": {"kernel": top, "avg_launch_ms": top_ms, "share_of_step": prof[top][1] / total_prof_ms,
                            "algorithmic_bytes_per_interaction": per_kernel[top],
                            "achieved": per_kernel[top] * B / (top_ms * 1e-3) / 1e9},
        "kernels_ms_per_step": {kk: round(v[1] / K, 5) for kk, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        "timing": "step: cudaEvent pair around %d steps; kernels: cudaEvent pair around each launch, separate pass" % K,
    }


def measure_stress(args, device, peaks, peak_kind):
    """The HBM-bound shape (BASELINE configs[4] on one GPU's share): n = 2 M features, k = 128, B = 2^20, 8 non-zeros
    per row -- a 2 GB parameter table and 1 GB of per-batch s_t rows, nothing L2-resident."""
    from rfm_b200 import _capi
    from rfm_b200._capi import check, lib
    from rfm_b200.fm import FactorizationMachines, _FmTrainer
    train, val, n_features, gen_s = make_stress_data(args.stress_rows, 2024)
    B, K, W = 1 << 20, max(3, min(args.steps, 10)), 3
    model = FactorizationMachines("IPS", K, STRESS_K, LR, B, 12345, n_features, dtype=args.dtype, sampler="feistel",
                                  device=device)
    ctx = model._context()
    fac_train = model._rows(train["features"], train["labels"], train["pscores"])
    fac_val = model._rows(val["features"], val["labels"], val["pscores"])
    model.sync_to_device()
    # the rows reach the device factored (0.3 GB for 20 M interactions); the resident steps run on the stacked CSR
    # assembled from them on the device (rfm_rows_materialize), like `value`; the factored steps are timed next to it
    t0 = time.perf_counter()
    csr_train, csr_val = _capi.MaterializedRows(fac_train), _capi.MaterializedRows(fac_val)
    ctx.synchronize()
    materialize_s = time.perf_counter() - t0
    results = {}
    for fmt, (train_rows, val_rows) in (("csr", (csr_train, csr_val)), ("factored", (fac_train, fac_val))):
        trainer = _FmTrainer(model._dev, train_rows, val_rows, B, W + 2 * K + 8)

        def stepper(epoch, slot):
            check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, epoch, B, LR, slot))

        results[fmt] = timed_steps(ctx, None, stepper, W, K, device, observe=(fmt == "csr"))
        trainer.close()
    ms, launches, clk, prof = results["csr"]
    s = 8 if args.dtype == "float64" else 4
    u = train["features"].users[:B].astype(np.int64)
    i = train["features"].items[:B].astype(np.int64)
    touched = np.unique(u).size + np.unique(i).size + sum(STRESS_GROUPS)
    roof = step_roofline(prof, K, ms, B, 1, 8.0, touched, STRESS_K, s, peaks, peak_kind, False,
                         "V (%.1f GB) and S (%.1f GB) exceed L2: id-column gathers and the s_t gather are HBM traffic"
                         % (n_features * STRESS_K * s / 1e9, B * STRESS_K * s / 1e9), "stress_csr")
    roof.update(value=K * B / (ms * 1e-3), value_unit=UNIT, ms_per_step=ms / K, steps=K, warmup=W, batch=B,
                n_factors=STRESS_K, n_features=n_features, train_interactions=args.stress_rows,
                input_format="factored rows uploaded (%.2f GB), stacked CSR assembled from them on the device in %.1f ms"
                             % (fac_train.h2d_bytes / 1e9, materialize_s * 1e3),
                other_input_format={"input_format": "factored", "ms_per_step": results["factored"][0] / K,
                                    "value": K * B / (results["factored"][0] * 1e-3)},
                data_gen_s=round(gen_s, 1), clocks=clk,
                workload="IPS-FM stress, BASELINE.json configs[4] shape on one GPU's share: 1M users x 1M items, "
                         "k=128, 8 non-zeros per row")
    return roof


def measure_c5(args, device, dist, world, peaks, peak_kind):
    """BASELINE configs[4]: 1 M users x 1 M items, 10^9 interactions, k = 128, full-catalog top-100 sharded over 8
    GPUs. The log is generated ON EVERY GPU by the device-side click model (rfm_b200.clicks, SURVEY 8 f4: counter-
    based, so every rank holds the same 10^9 factored rows, 16 bytes each) and trained data-parallel with the
    reference's step (global batch 2^21 per GPU, one global Feistel permutation, NVLink reduce + apply); then every
    item is ranked for every user, item-sharded, top-100."""
    from rfm_b200._capi import check, lib, ptr
    from rfm_b200.clicks import ClickModel, GeneratedRows
    from rfm_b200.fm import FactorizationMachines, _FmTrainer
    from rfm_b200.score import TopKScorer
    from scipy.sparse import csr_matrix
    U, I, k = args.c5_users, args.c5_items, STRESS_K
    rows_total = args.c5_rows or 125_000_000 * world
    B = min(1 << 21, max(1024, rows_total // (4 * world)))
    K, W = max(3, min(args.steps, 6)), 3
    t0 = time.perf_counter()
    model_c = ClickModel(U, I, seed=2024)

    def side(n, groups, mult0):
        cols = np.empty((n, 3), dtype=np.int32)
        ids = np.arange(n, dtype=np.int64)
        off = 0
        for j, g in enumerate(groups):
            cols[:, j] = off + (ids * (j + mult0) + j) % g
            off += g
        return csr_matrix((np.ones(n * 3), cols.ravel(), np.arange(0, 3 * n + 1, 3, dtype=np.int32)), shape=(n, off))

    blocks = [("id", "user", U), ("table", "user", side(U, STRESS_GROUPS[:3], 3)), ("id", "item", I),
              ("table", "item", side(I, STRESS_GROUPS[3:], 5))]
    n_features = U + I + sum(STRESS_GROUPS)
    fm = FactorizationMachines("IPS", K, k, LR, B * world, 12345, n_features, dtype=args.dtype, sampler="feistel",
                               device=device)
    ctx = fm._context()
    host_s = time.perf_counter() - t0
    ctx.synchronize()
    t0 = time.perf_counter()
    train = GeneratedRows(ctx, model_c, rows_total, blocks, row0=0, dtype=args.dtype)
    val = GeneratedRows(ctx, model_c, N_VAL, blocks, row0=rows_total, dtype=args.dtype)
    ctx.synchronize()
    gen_s = time.perf_counter() - t0
    fm.sync_to_device()
    trainer = _FmTrainer(fm._dev, train, val, B, W + 2 * K + 8)
    dp = None
    if dist is not None:
        from rfm_b200 import dist as rdist
        dp = rdist.make_fm_dp(fm, trainer, dist, B * world, N_VAL, LR, lambda epoch: None)

        def stepper(epoch, slot):
            dp.step(epoch)
    else:
        def stepper(epoch, slot):
            check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, epoch, B, LR, slot))
    ms, launches, clk, prof = timed_steps(ctx, dist, stepper, W, K, device, observe=False)
    if dist is None:
        tl, vl = np.empty(W + K), np.empty(W + K)
        check(lib().rfm_fm_trainer_losses(trainer.handle, 0, W + K, ptr(tl), ptr(vl)))
        losses = (float(tl[-1]), float(vl[-1]))
    else:
        last = dp.flush().cpu().numpy()
        losses = (float(last[0] / (B * world)), float(last[1] / N_VAL))
    assert np.all(np.isfinite(losses)), "non-finite loss in the c5 section"
    s = 8 if args.dtype == "float64" else 4
    d = train.download(0, min(B, rows_total))
    touched = np.unique(d["users"]).size + np.unique(d["items"]).size + sum(STRESS_GROUPS)
    roof = step_roofline(prof, K, ms, B, world, 8.0, touched, k, s, peaks, peak_kind, False,
                         "V (%.1f GB) and S (%.1f GB) exceed L2; per rank" % (n_features * k * s / 1e9, B * k * s / 1e9),
                         "c5")
    train_part = dict(roof, value=K * B * world / (ms * 1e-3), value_unit=UNIT, ms_per_step=ms / K, steps=K, warmup=W,
                      batch_per_gpu=B, global_batch=B * world, interactions=rows_total, n_features=n_features,
                      n_factors=k, rows_bytes_per_gpu=rows_total * (8 + s), generate_seconds=gen_s, host_setup_seconds=host_s,
                      final_train_loss=losses[0], final_val_loss=losses[1], gpu_launches=launches,
                      data="every rank generated all %d rows on its GPU (Philox click model), %.1f s" % (rows_total, gen_s))
    trainer.close()
    del train, val, trainer
    # ---- scoring half: rank every item for every user, item-sharded, top-100
    rng = np.random.default_rng(11)
    A = rng.normal(size=(U, k)) * 0.3
    C = rng.normal(size=(I, k)) * 0.3
    beta = rng.normal(size=I) * 0.2
    sc = TopKScorer(A, C, None, beta, 0.0, device=device)
    Ktop = 100

    def call():
        if dist is not None:
            from rfm_b200 import dist as rdist
            return rdist.sharded_topk(sc, dist, Ktop, gather=False, copy=False)
        return sc.topk(Ktop, copy=False)

    call()
    if dist is not None:
        dist.barrier()
    ctx.synchronize()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        out = call()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if dist is not None:
        dt = dist.max_over_ranks(dt)
    ctx.profile_begin()
    call()
    sprof = ctx.profile_end()
    fk = sum(v[1] for k2, v in sprof.items() if k2 in ("score_sample", "score_collect"))      # v = (launches, total ms)
    i_local = -(-I // world)
    tf = 2.0 * (-(-U // 128) * 128) * (-(-i_local // 256) * 256) * 128 / (fk * 1e-3) / 1e12 if fk > 0 else None
    # spot check: the first owned users' lists against a float64 argsort over the whole catalog
    items_own = out[2] if dist is not None else out[0]
    first_user = out[0] if dist is not None else 0
    chk = min(4, items_own.shape[0])
    S = A[first_user:first_user + chk] @ C.T + beta[None, :]
    ref = np.argsort(S, axis=1, kind="stable")[:, ::-1][:, :Ktop]
    exact = bool(np.array_equal(items_own[:chk], ref.astype(np.int32)))
    score_part = {"users": U, "items": I, "n_factors": k, "top_k": Ktop, "value": U * I / dt, "unit": "pairs/s",
                  "ms_per_call": dt * 1e3, "users_ranked_exactly": sc.last_stats.get("users_ranked_exactly"),
                  "candidates_per_user": round(sc.last_stats.get("candidates", 0) / U, 2),
                  "kernels_ms": {k2: [v[0], round(v[1], 3)] for k2, v in sprof.items()},
                  "spot_check_vs_float64_argsort": exact,
                  "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                               "frac": tf / peaks["bf16_tflops_sustained"] if tf else None,
                               "note": "per rank; 2 U I_local k flops (one pass) over the time of both tensor passes"}}
    assert exact, "c5 scoring spot check failed"
    sc.close()
    return {"workload": "BASELINE.json configs[4]: %d users x %d items, %d interactions, k=%d, top-%d, %d GPU(s)"
                        % (U, I, rows_total, k, Ktop, world), "train": train_part, "scoring": score_part}


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
    peaks, peak_kind = measured_peaks()
    parity = dp_parity(dist, local_rank) if dist is not None else None

    if args.workload == "c5":
        c5 = measure_c5(args, local_rank, dist, world, peaks, peak_kind)
        if rank == 0:
            tr = c5["train"]
            print(json.dumps({"metric": METRIC, "value": tr["value"], "unit": UNIT, "n_gpus": world, "steps": tr["steps"],
                              "warmup": tr["warmup"], "ms_per_step": tr["ms_per_step"], "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.dtype == "float64" else "f32",
                              "data": "synthetic", "config": {"workload": c5["workload"]}, "clocks": None,
                              "roofline": dict(tr, scoring=c5["scoring"]), "gpu_launches": tr["gpu_launches"],
                              "dp_parity": parity}))
        if dist is not None:
            dist.shutdown()
        return
    if args.workload == "stress":
        roof = measure_stress(args, local_rank, peaks, peak_kind)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": roof["value"], "unit": UNIT, "n_gpus": 1, "steps": roof["steps"],
                              "warmup": roof["warmup"], "ms_per_step": roof["ms_per_step"], "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.dtype == "float64" else "f32",
                              "data": "synthetic", "config": {"workload": roof["workload"]}, "clocks": roof["clocks"],
                              "roofline": roof, "gpu_launches": None}))
        return

    # data-parallel runs replicate the train set on every GPU: same seed everywhere
    log, gen_s = make_data(args.rows, 2024)
    X = log.fm_train["features"]
    ftrain, fval = factored_dicts(log)
    pinned = pin_host_arrays([X.indptr, X.indices, X.data, log.fm_train["labels"], log.fm_train["pscores"],
                              ftrain["features"].users, ftrain["features"].items, ftrain["labels"]]
                             + [b[1] for b in ftrain["features"].blocks if b[0] == "ctx"])
    log.pinned_note = "%d train arrays page-locked (cudaHostRegister)" % len(pinned)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    s = 8 if args.dtype == "float64" else 4
    dtype_tag = "f64" if args.dtype == "float64" else "f32"
    train_in, val_in = (ftrain, fval) if args.input == "factored" else (log.fm_train, log.fm_val)

    def device_run(train_d, val_d, step="flat"):
        model = FactorizationMachines("IPS", K, K_FACTORS, LR, B, 12345, log.n_features, dtype=args.dtype,
                                      sampler="feistel", device=local_rank)
        ctx = model._context()
        train_rows = model._rows(train_d["features"], train_d["labels"], train_d["pscores"])
        val_rows = model._rows(val_d["features"], val_d["labels"], val_d["pscores"])
        model.sync_to_device()
        trainer = _FmTrainer(model._dev, train_rows, val_rows, B, W + 2 * K + 8)
        if step != "flat" and getattr(train_rows, "factored", False):
            trainer.set_two_level(1 if step == "two_level" else 2)
        dp = None
        if dist is not None:
            from rfm_b200 import dist as rdist
            model.batch_size = B * world                       # weak scaling: B per GPU, global batch B*world
            dp = rdist.make_fm_dp(model, trainer, dist, B * world, N_VAL, LR, lambda epoch: None)

            def stepper(epoch, slot):
                dp.step(epoch)
        else:
            def stepper(epoch, slot):
                check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, epoch, B, LR, slot))
        ms, launches, clk, prof = timed_steps(ctx, dist, stepper, W, K, local_rank)
        tl = np.empty(W + 2 * K)
        vl = np.empty(W + 2 * K)
        if dist is None:
            check(lib().rfm_fm_trainer_losses(trainer.handle, 0, W + 2 * K, ptr(tl), ptr(vl)))
        else:
            last = dp.flush().cpu().numpy()
            tl[:] = last[0] / (B * world)
            vl[:] = last[1] / N_VAL
        assert np.all(np.isfinite(tl)) and np.all(np.isfinite(vl)), "non-finite loss in the timed region"
        trainer.close()
        return dict(ms=ms, launches=launches, clk=clk, prof=prof, tl=float(tl[W + K - 1]), vl=float(vl[W + K - 1]),
                    rows_bytes=train_rows.h2d_bytes, two_level=trainer.two_level)

    main = device_run(train_in, val_in, args.step if args.input == "factored" else "flat")
    others = []
    for fmt, step in (("csr", "flat"), ("factored", "flat"), ("factored", "two_level")):
        if fmt == args.input and (fmt == "csr" or (step == "two_level") == main["two_level"]):
            continue                                         # that is the headline run itself
        o = device_run(*((log.fm_train, log.fm_val) if fmt == "csr" else (ftrain, fval)), step)
        others.append({"input_format": fmt, "step": "two_level" if o["two_level"] else "flat",
                       "ms_per_step": o["ms"] / K, "value": K * B * world / (o["ms"] * 1e-3),
                       "final_train_loss": o["tl"],
                       "kernels_ms_per_step": {kk: round(v[1] / K, 5) for kk, v in
                                               sorted(o["prof"].items(), key=lambda kv: -kv[1][1])}})
    ms, launches, clk, prof = main["ms"], main["launches"], main["clk"], main["prof"]
    value = K * B * world / (ms * 1e-3)

    run_c5 = world == 8 and not args.no_c5
    if rank != 0:
        if not args.no_e2e:
            measure_e2e(args, log, ftrain, fval, local_rank, dist, world)
            measure_sharded_predict(args, log, ftrain, local_rank, dist, world)
        if not args.no_scoring:
            measure_scoring(local_rank, dist, world, peaks, peak_kind)
        if run_c5:
            del log, X, ftrain, fval, train_in, val_in
            measure_c5(args, local_rank, dist, world, peaks, peak_kind)
        dist.shutdown()
        return

    sample_rows = _capi.feistel_batch(X.shape[0], B, W, 12345)
    m, touched, step_bytes, per_kernel = algorithmic_bytes(X, sample_rows, K_FACTORS, s)
    roofline = step_roofline(prof, K, ms, B, world, m, touched, K_FACTORS, s, peaks, peak_kind, True,
                             "V (%.1f MB) and S (%.1f MB) are L2-resident at this shape, so the algorithmic gather "
                             "traffic is served by L2, not HBM (SURVEY.md H7): frac is the section-8(d) fraction, "
                             "L2-assisted; the HBM-bound claim is roofline.stress"
                             % (log.n_features * K_FACTORS * s / 1e6, B * K_FACTORS * s / 1e6),
                             "kuairec_two_level" if main["two_level"] else
                             ("kuairec_factored" if args.input == "factored" else "kuairec_big"))
    roofline["other_input_formats"] = others
    roofline["step"] = "two_level" if main["two_level"] else "flat"
    if main["two_level"]:
        ff = ftrain["features"]
        n_ctx = sum(b[1].shape[1] for b in ff.blocks if b[0] == "ctx")
        entries = n_ctx + sum(b[2] for b in ff.blocks if b[0] == "id") + sum(b[2].nnz for b in ff.blocks if b[0] == "table")
        own = two_level_bytes(B, K_FACTORS, s, n_ctx, entries, log.n_features)
        own_achieved = own * B / (ms / K * 1e-3) / 1e9
        roofline["two_level"] = {
            "algorithmic_bytes_per_interaction": own, "achieved": own_achieved, "frac": own_achieved / peaks["hbm_gbs"],
            "entity_entries": int(entries),
            "note": "the two-level step (csrc/two_level.cuh) computes the same update from per-user / per-item "
                    "aggregates: it gathers 2 + n_ctx parameter rows per interaction and pass instead of m = %.1f, plus "
                    "one pass over the entity tables per step, i.e. %.0f bytes per interaction where section 8(d) "
                    "counts %.0f. roofline.frac keeps the section-8(d) numerator (the contract's definition), so it "
                    "measures the step against the bytes the FLAT formulation must move and may exceed what a "
                    "byte-for-byte implementation could reach; this object holds the fraction in the step's own bytes"
                    % (m, own, roofline["algorithmic_bytes_per_interaction"])}
    e2e = None if args.no_e2e else measure_e2e(args, log, ftrain, fval, local_rank, dist, world)
    if e2e is not None and dist is not None:
        e2e["sharded_predict"] = measure_sharded_predict(args, log, ftrain, local_rank, dist, world)
    scoring = None if args.no_scoring else measure_scoring(local_rank, dist, world, peaks, peak_kind)
    if scoring is not None:
        roofline["scoring"] = {name: {"pairs_per_s": v["value"], "ms_per_call": v["ms_per_call"],
                                      "tensor_frac": (v.get("roofline") or {}).get("frac"),
                                      "scoring_efficiency": v.get("scoring_efficiency")}
                               for name, v in scoring.items() if isinstance(v, dict)}
    if world == 1 and not args.no_stress:
        roofline["stress"] = measure_stress(args, local_rank, peaks, peak_kind)
    if run_c5:      # BASELINE configs[4] rides in the 8-GPU line (roofline.c5): 10^9 interactions, 1 M x 1 M top-100
        roofline["c5"] = measure_c5(args, local_rank, dist, world, peaks, peak_kind)
    cpu = None
    if not args.no_cpu_baseline:
        v1, st1, dt1 = cpu_port_run(log, B, 8, 1, budget_s=8.0)
        v, st, dt, cores = cpu_port_run_parallel(log, B, 160, 1, budget_s=15.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "host_cores_available": os.cpu_count(),
               "sample": "%d epochs of B=%d on the %d-row train set (stacked CSR, the reference's input), reference "
                         "sampler included, %.1f s on %d processes" % (st, B, X.shape[0], dt, cores),
               "single_core": {"value": v1, "sample": "%d epochs, %.1f s" % (st1, dt1)}}
    if world == 1 and not args.no_scoring and (not args.no_cpu_baseline or not args.no_e2e):
        s_cpu, s_e2e = scoring_cpu_and_e2e(local_rank)
        if cpu is not None:
            cpu["scoring"] = s_cpu
        if e2e is not None:
            e2e["scoring"] = s_e2e
    if world == 1 and not args.no_search:
        srch, srch_cpu = measure_epoch_search(local_rank, cpu is not None)
        if e2e is not None:
            e2e["epoch_search"] = srch
        if cpu is not None and srch_cpu is not None:
            cpu["epoch_search"] = srch_cpu
    if world == 1 and not args.no_mf:
        mf, mf_cpu = measure_mf(local_rank, cpu is not None)
        if e2e is not None:
            e2e["mf"] = mf
        if cpu is not None and mf_cpu is not None:
            cpu["mf"] = mf_cpu
    cfg = workload_config(args, world)
    cfg["l2"] = cfg["l2"].replace("train rows", "%.2f GB of train rows" % (main["rows_bytes"] / 1e9))
    cfg.update(sampler="feistel (device, perf mode)", mean_nnz_per_row=round(m, 3),
               touched_columns_per_step=int(touched), n_features=log.n_features, data_gen_s=round(gen_s, 1),
               input_format=("factored rows resident in HBM: user table + item table + (user, item, ctx) records "
                             "(SURVEY.md 8 f3), %s step" % ("two-level" if main["two_level"] else "flat")
                             if args.input == "factored" else
                             "stacked CSR (the reference's own input) resident in HBM; e2e uploads the %s form"
                             % args.e2e_input))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_tag, "data": "synthetic", "config": cfg, "clocks": clk, "e2e": e2e,
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "scoring": scoring,
        "final_train_loss": main["tl"], "final_val_loss": main["vl"],
    }
    if parity is not None:
        line["dp_parity"] = parity
        line["config"]["dp_parity"] = parity
    print(json.dumps(line))
    for a in pinned:
        _capi.unpin_array(a)
    if dist is not None:
        dist.shutdown()


def measure_sharded_predict(args, log, ftrain, device, dist, world):
    """Row-sharded predict (independent rows): every rank scores 1/N of the rows with its replica
    of the parameters, the score slices are all-gathered. Timed from host rows to host scores on every rank, next to
    the single-process predict of the same rows in the same run; bit-equality of the two is asserted."""
    from rfm_b200 import dist as rdist
    from rfm_b200.fm import FactorizationMachines
    n = min(4_000_000, ftrain["features"].shape[0])
    X = ftrain["features"][:n]
    m = FactorizationMachines("IPS", 1, K_FACTORS, LR, args.batch, 12345, log.n_features, dtype=args.dtype, device=device)
    m.sync_to_device()

    def timed(fn, reps=3):
        m.reset_rows_cache()
        out = np.array(fn())             # warm-up and the result that is compared (a copy: the sharded call returns a
        m.reset_rows_cache()             # view of a page-locked buffer, which must go back to torch's cache)
        fn()                             # steady state from here on: that buffer (a 14 ms cudaHostAlloc) is reused
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            m.reset_rows_cache()         # every call uploads its rows again, as a fresh call would
            fn()
        return dist.max_over_ranks((time.perf_counter() - t0) / reps), out

    t_single, ref = timed(lambda: m.predict(X=X))
    t_shard, got = timed(lambda: rdist.sharded_predict(m, X, dist))
    assert np.array_equal(ref, got), "row-sharded predict differs from the single-process predict"
    return {"rows": int(n), "unit": "rows/s", "value": n / t_shard, "ms_per_call": t_shard * 1e3,
            "single_gpu_ms": t_single * 1e3, "speedup_vs_1gpu": t_single / t_shard, "bit_identical": True,
            "api": "rfm_b200.dist.sharded_predict(model, FactoredFeatures, env): host rows -> each rank uploads and scores "
                   "its slice (rfm_fm_predict_dev into the send buffer) -> NCCL all-gather -> one copy into page-locked "
                   "host memory on every rank; steady state (the first call also allocates that buffer)"}


def measure_e2e(args, log, ftrain, fval, device, dist, world):
    """Public API on host arrays: FactorizationMachines.fit(train, val) for K epochs. Everything a
    user pays is inside the timed region: row upload from pinned host memory, trainer set-up, K
    epochs, loss read-back, parameter download."""
    from rfm_b200.fm import FactorizationMachines
    B, K = args.batch, args.steps
    out = {}
    variants = [("factored", "feistel", ftrain, fval), ("csr", "feistel", log.fm_train, log.fm_val),
                ("factored", "legacy", ftrain, fval)]
    for fmt, sampler, train, val in variants:
        n_ep = K if sampler == "feistel" else min(K, 16)
        if dist is not None and sampler == "legacy":
            continue
        warm = FactorizationMachines("IPS", 2, K_FACTORS, LR, B * world, 12345, log.n_features, dtype=args.dtype,
                                     sampler=sampler, device=device, distributed=dist)
        warm.fit(train, val)
        del warm
        model = FactorizationMachines("IPS", n_ep, K_FACTORS, LR, B * world, 12345, log.n_features,
                                      dtype=args.dtype, sampler=sampler, device=device, distributed=dist)
        if dist is not None:
            dist.barrier()
        model._context().synchronize()
        t0 = time.perf_counter()
        tl, vl = model.fit(train, val)
        model._context().synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            dt = dist.max_over_ranks(dt)
        rows_bytes = model.last_fit_stats["h2d_bytes_rows"]
        out[(fmt, sampler)] = {
            "value": n_ep * B * world / dt, "epochs": n_ep, "seconds": dt,
            "upload_seconds": model.last_fit_stats.get("upload_seconds"),
            "phase_seconds": model.last_fit_stats.get("phase_seconds"),
            "h2d_bytes_per_step": rows_bytes / n_ep + (B * 8 if sampler == "legacy" else 0),
            "d2h_bytes_per_step": 16 + (1 + log.n_features * (K_FACTORS + 1)) * 8 / n_ep,
        }
    head_fmt = args.e2e_input
    main = out[(head_fmt, "feistel")]
    apis = {"factored": "FactorizationMachines(sampler='feistel').fit(train, val), train['features'] a FactoredFeatures "
                        "(the blocks the reference's preparer stacks + one (user, item, ctx) record per interaction) on "
                        "pinned host arrays; includes the one-time upload, amortised over the epochs of this call",
            "csr": "FactorizationMachines(sampler='feistel').fit(train, val) on the reference's stacked CSR in pinned "
                   "host arrays; includes the one-time CSR upload, amortised over the epochs of this call"}
    res = {"value": main["value"], "unit": UNIT, "h2d_bytes_per_step": main["h2d_bytes_per_step"],
           "d2h_bytes_per_step": main["d2h_bytes_per_step"], "seconds": main["seconds"], "epochs": main["epochs"],
           "upload_seconds": main["upload_seconds"], "host_memory": getattr(log, "pinned_note", None),
           "phase_seconds": main["phase_seconds"], "input_format": head_fmt, "api": apis[head_fmt]}
    alt = "csr" if head_fmt == "factored" else "factored"
    res["hstacked_csr" if alt == "csr" else "factored"] = dict(out[(alt, "feistel")], api=apis[alt])
    if ("factored", "legacy") in out:
        res["legacy_sampler"] = dict(out[("factored", "legacy")], input_format="factored",
                                     note="reference batch order (RandomState(epoch) shuffle of all N ids on host "
                                          "threads, SURVEY.md F14): host-bound by design")
    return res


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
