import sys, time, json
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")  # run from the repository root
import numpy as np
from rfm_b200.score import TopKScorer
from rfm_b200 import _capi
rng = np.random.default_rng(0)
out = {}
for name, U, I, k, K in (("c4_eval_grid", 1411, 3327, 64, 9), ("large_k64", 32768, 262144, 64, 9), ("large_k128_top100", 16384, 131072, 128, 100)):
    A = rng.normal(size=(U, k)) * 0.3; C = rng.normal(size=(I, k)) * 0.3; beta = rng.normal(size=I) * 0.2
    sc = TopKScorer(A, C, None, beta, 0.0)
    ctx = sc.ctx
    sc.topk(K)
    ctx.profile_begin(); t0 = time.perf_counter(); items, scores = sc.topk(K); dt = time.perf_counter() - t0; prof = ctx.profile_end()
    fk = (1, sum(v[0] * v[1] for k2, v in prof.items() if k2 in ("score_sample", "score_collect")))
    upad, ipad, kpad = -(-U // 128) * 128, -(-I // 256) * 256, -(-k // 64) * 64
    out[name] = dict(users=U, items=I, k=k, K=K, pairs_per_s_e2e=U * I / dt, call_ms=dt * 1e3,
                     filter_ms=fk[1], filter_tflops=2.0 * upad * ipad * kpad / (fk[1] * 1e-3) / 1e12 if fk[1] else None,
                     pairs_per_s_filter=U * I / (fk[1] * 1e-3) if fk[1] else None,
                     kernels_ms={k2: [v[0], round(v[1], 4)] for k2, v in prof.items()}, stats=sc.last_stats)
print(json.dumps(out))
