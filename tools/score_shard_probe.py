"""Item-sharded scoring over N GPUs (torchrun): per-shape ms/call, single-GPU ms, kernel split. Not a test.
    python -m torch.distributed.run --nproc-per-node N tools/score_shard_probe.py"""
import os, sys, json
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")
import bench
from rfm_b200 import dist as rdist
local = int(os.environ.get("LOCAL_RANK", "0"))
env = rdist.init(local) if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None
peaks, kind = bench.measured_peaks()
out = bench.measure_scoring(local, env, env.world if env else 1, peaks, kind)
if env is None or env.rank == 0:
    for k, v in out.items():
        if isinstance(v, dict):
            print(k, "ms %.4f" % v["ms_per_call"], "single %s" % v.get("single_gpu_ms"), "eff %s" % v.get("scoring_efficiency"),
                  {a: b[1] for a, b in v["kernels_ms"].items()}, flush=True)
if env is not None:
    env.shutdown()
