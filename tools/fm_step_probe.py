"""Device-timed FM steps at the KuaiRec shape for both row formats, with the per-kernel split. Not a test.
    python tools/fm_step_probe.py [rows] [steps] [dtype]"""
import sys, time
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")
import numpy as np
import bench
from rfm_b200._capi import check, lib
from rfm_b200.fm import FactorizationMachines, _FmTrainer
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dtype = sys.argv[3] if len(sys.argv) > 3 else "float64"
log, _ = bench.make_data(rows, 2024)
ftrain, fval = bench.factored_dicts(log)
B, W = 65536, 5
for name, tr, va in (("csr", log.fm_train, log.fm_val), ("factored", ftrain, fval)):
    m = FactorizationMachines("IPS", K, 64, bench.LR, B, 12345, log.n_features, dtype=dtype, sampler="feistel")
    ctx = m._context()
    trr, var = m._rows(tr["features"], tr["labels"], tr["pscores"]), m._rows(va["features"], va["labels"], va["pscores"])
    m.sync_to_device()
    t = _FmTrainer(m._dev, trr, var, B, W + 2 * K + 8)
    step = lambda e, s: check(lib().rfm_fm_train_epoch_sampled(t.handle, 12345, e, B, bench.LR, s))
    for e in range(W):
        step(e, e)
    ctx.synchronize(); ctx.timer_start()
    for e in range(K):
        step(W + e, W + e)
    ms = ctx.timer_stop_ms()
    ctx.profile_begin()
    for e in range(K):
        step(W + K + e, W + K + e)
    prof = ctx.profile_end()
    print(name, dtype, "ms/step %.4f" % (ms / K), {k: round(v[1] / K * 1e3, 1) for k, v in prof.items()}, flush=True)
    t.close()
