"""Where the time of rfm_b200.dist.sharded_predict goes (run under torchrun). Not a test."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "relevance-factorizationmachine_b200"))
import bench  # noqa: E402
from rfm_b200 import _capi, dist as rdist  # noqa: E402
from rfm_b200.fm import FactorizationMachines  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
env = rdist.init(local_rank)
log, _ = bench.make_data(4_000_000, 2024)
ftrain, fval = bench.factored_dicts(log)
ff = ftrain["features"]
bench.pin_host_arrays([ff.users, ff.items] + [b[1] for b in ff.blocks if b[0] == "ctx"])
m = FactorizationMachines("IPS", 1, 64, bench.LR, 65536, 12345, log.n_features, device=local_rank)
m.sync_to_device()
torch = env.torch
n = ff.shape[0]
begin, end = rdist.slice_bounds(n, env.world, env.rank)
for rep in range(3):
    m.reset_rows_cache()
    env.barrier()
    t = [time.perf_counter()]
    Xs = ff[begin:end]
    t.append(time.perf_counter())
    rows = m._rows(Xs)
    m._ctx.synchronize()
    t.append(time.perf_counter())
    send = torch.zeros(end - begin, dtype=torch.float64, device="cuda:%d" % local_rank)
    from ctypes import c_void_p
    _capi.check(_capi.lib().rfm_fm_predict_dev(m._dev.handle, rows.handle, c_void_p(send.data_ptr())))
    m._ctx.synchronize()
    t.append(time.perf_counter())
    recv = torch.empty(env.world * (end - begin), dtype=torch.float64, device="cuda:%d" % local_rank)
    env.dist.all_gather_into_tensor(recv, send)
    torch.cuda.synchronize(local_rank)
    t.append(time.perf_counter())
    host = torch.empty(recv.shape, dtype=torch.float64, pin_memory=True)
    t.append(time.perf_counter())
    host.copy_(recv, non_blocking=True)
    torch.cuda.synchronize(local_rank)
    t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("rank %d rep %d ms: slice %.2f  rows(upload) %.2f  predict %.2f  all_gather %.2f  pinned alloc %.2f  d2h %.2f  total %.2f"
          % (env.rank, rep, *d, d.sum()), flush=True)
    t0 = time.perf_counter()
    m.reset_rows_cache()
    full = m.predict(X=ff)
    print("rank %d rep %d single-process predict %.2f ms" % (env.rank, rep, (time.perf_counter() - t0) * 1e3), flush=True)
    del host
env.shutdown()
