"""Two epochs of the two-level FM step at the KuaiRec-big shape (3 M rows by default) for an ncu capture. Not a test.
    python tools/ncu_two_level_probe.py [rows] [dtype]"""
import sys
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")
import bench
from rfm_b200._capi import check, lib
from rfm_b200.fm import FactorizationMachines, _FmTrainer

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
dtype = sys.argv[2] if len(sys.argv) > 2 else "float64"
log, _ = bench.make_data(rows, 2024)
ftrain, fval = bench.factored_dicts(log)
B = 65536
m = FactorizationMachines("IPS", 2, 64, bench.LR, B, 12345, log.n_features, sampler="feistel", dtype=dtype)
ctx = m._context()
trr = m._rows(ftrain["features"], ftrain["labels"], ftrain["pscores"])
var = m._rows(fval["features"], fval["labels"], fval["pscores"])
m.sync_to_device()
t = _FmTrainer(m._dev, trr, var, B, 4)
assert t.set_two_level(1)
ctx.synchronize()
print("PHASE setup done", flush=True)
for e in range(2):
    check(lib().rfm_fm_train_epoch_sampled(t.handle, 12345, e, B, bench.LR, e))
    ctx.synchronize()
    print("PHASE epoch", e, flush=True)
t.close()
