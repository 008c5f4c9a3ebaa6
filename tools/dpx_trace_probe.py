"""Phase times of the data-parallel exchange kernel (RFM_DPX_TRACE=1) at the bench shape. Run under torchrun:
    RFM_DPX_TRACE=1 python -m torch.distributed.run --nproc-per-node N tools/dpx_trace_probe.py [rows]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "relevance-factorizationmachine_b200"))
import bench  # noqa: E402

os.environ["RFM_DPX_TRACE"] = "1"
from rfm_b200 import dist as rdist  # noqa: E402
from rfm_b200._capi import check, lib, ptr  # noqa: E402
from rfm_b200.fm import FactorizationMachines, _FmTrainer  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
env = rdist.init(local_rank)
log, _ = bench.make_data(rows, 2024)
ftrain, fval = bench.factored_dicts(log)
B = 65536
model = FactorizationMachines("IPS", 4, 64, bench.LR, B, 12345, log.n_features, sampler="feistel", device=local_rank)
ctx = model._context()
tr = model._rows(ftrain["features"], ftrain["labels"], ftrain["pscores"])
va = model._rows(fval["features"], fval["labels"], fval["pscores"])
model.sync_to_device()
trainer = _FmTrainer(model._dev, tr, va, B, 8)
trainer.set_two_level(2)
model.batch_size = B * world
dp = rdist.make_fm_dp(model, trainer, env, B * world, 2000, bench.LR, lambda epoch: None)
acc = []
for e in range(40):
    dp.step(e)
    if e >= 10:
        ctx.synchronize()
        st = np.zeros(8, dtype=np.uint64)
        check(lib().rfm_fm_dp_trace(trainer.handle, ptr(st)))
        acc.append(np.diff(st[:5].astype(np.int64)) / 1e3)
acc = np.array(acc)
med = np.median(acc, axis=0)
print("rank %d world %d  us: barrier0 %.1f  reduce %.1f  barrier1 %.1f  apply %.1f  total %.1f" %
      (env.rank, world, med[0], med[1], med[2], med[3], med.sum()), flush=True)
dp.flush()
trainer.close()
env.shutdown()
