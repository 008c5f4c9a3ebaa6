"""Single large scoring call for ncu captures of score_filter_kernel (not a test)."""
import sys
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")  # run from the repository root
import numpy as np
from rfm_b200.score import TopKScorer
U, I, k, K = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (16384, 131072, 64, 9)))
rng = np.random.default_rng(0)
sc = TopKScorer(rng.normal(size=(U, k)) * 0.3, rng.normal(size=(I, k)) * 0.3, None, rng.normal(size=I) * 0.2, 0.0)
items, scores = sc.topk(K)
print("ok", sc.last_stats, float(scores[0, 0]))
