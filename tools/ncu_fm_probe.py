"""A few FM steps per (shape, row format) for an ncu capture: KuaiRec-big shape with stacked-CSR rows, the same with
factored rows, then the stress shape (factored). Prints PHASE markers so the launch list can be cut. Not a test.
    python tools/ncu_fm_probe.py [rows_kuairec] [rows_stress]"""
import sys
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")
import numpy as np
import bench
from rfm_b200._capi import check, lib
from rfm_b200.fm import FactorizationMachines, _FmTrainer

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
rows_stress = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
W, K = 0, 1


def run(name, train, val, n_features, k, B, materialize=False):
    from rfm_b200 import _capi
    m = FactorizationMachines("IPS", K, k, bench.LR, B, 12345, n_features, sampler="feistel")
    ctx = m._context()
    trr = m._rows(train["features"], train["labels"], train["pscores"])
    var = m._rows(val["features"], val["labels"], val["pscores"])
    if materialize:
        trr, var = _capi.MaterializedRows(trr), _capi.MaterializedRows(var)
    m.sync_to_device()
    t = _FmTrainer(m._dev, trr, var, B, W + K + 2)
    l0 = ctx.launch_count()
    for e in range(W + K):
        check(lib().rfm_fm_train_epoch_sampled(t.handle, 12345, e, B, bench.LR, e))
    ctx.synchronize()
    print("PHASE", name, "launches", ctx.launch_count() - l0, flush=True)
    t.close()


log, _ = bench.make_data(rows, 2024)
ftrain, fval = bench.factored_dicts(log)
run("kuairec_csr", log.fm_train, log.fm_val, log.n_features, 64, 65536)
run("kuairec_factored", ftrain, fval, log.n_features, 64, 65536)
train, val, n_features, _ = bench.make_stress_data(rows_stress, 2024)
run("stress_factored", train, val, n_features, 128, 1 << 20)
run("stress_csr", train, val, n_features, 128, 1 << 20, materialize=True)
