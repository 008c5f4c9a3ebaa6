import os, sys
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")
import numpy as np
os.environ.setdefault("MASTER_ADDR","127.0.0.1"); os.environ.setdefault("MASTER_PORT","29533")
os.environ.setdefault("RANK","0"); os.environ.setdefault("WORLD_SIZE","1"); os.environ.setdefault("LOCAL_RANK","0")
from rfm_b200 import dist as rdist
from rfm_b200.score import TopKScorer
env = rdist.init(int(os.environ["LOCAL_RANK"]))
rng = np.random.default_rng(21)
for n_u, n_i, k, K in [(500,3000,64,9),(300,5000,128,100),(130,300,32,9),(2000,70000,64,9)]:
    A, C, beta = rng.normal(size=(n_u, k)) * 0.4, rng.normal(size=(n_i, k)) * 0.4, rng.normal(size=n_i) * 0.2
    sc = TopKScorer(A, C, None, beta, 0.0, device=env.device)
    fi, fs = sc.topk(K)
    st_full = dict(sc.last_stats)
    it, scs = rdist.sharded_topk(sc, env, K)
    np.testing.assert_array_equal(it, fi); np.testing.assert_array_equal(scs, fs)
    print(env.rank, (n_u,n_i,k,K), "ok", st_full, sc.last_stats, flush=True)
    sc.close()
env.shutdown()
