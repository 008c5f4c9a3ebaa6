"""Kernel times of one large scoring call (timing experiments; not a test)."""
import sys, json
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")  # run from the repository root
import numpy as np
from rfm_b200.score import TopKScorer
U, I, k, K = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (32768, 262144, 64, 9)))
rng = np.random.default_rng(0)
sc = TopKScorer(rng.normal(size=(U, k)) * 0.3, rng.normal(size=(I, k)) * 0.3, None, rng.normal(size=I) * 0.2, 0.0)
sc.topk(K)
sc.ctx.profile_begin(); sc.topk(K); prof = sc.ctx.profile_end()
print(json.dumps({k2: [v[0], round(v[1], 4)] for k2, v in prof.items()}), sc.last_stats)
