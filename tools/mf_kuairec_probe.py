"""MF fit at the KuaiRec big_matrix shape (sequential per-sample SGD semantics -> wavefront levels). Not a test."""
import sys, time
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")  # run from the repository root
import numpy as np
from rfm_b200.mf import LogisticMatrixFactorization
from rfm_b200.synth import make_kuairec_shaped
log = make_kuairec_shaped(seed=2024, n_train=2_000_000, n_val=2000)
print("train pairs", log.mf_train["features"].shape)
for B in (2000, 65536):
    epochs = 12
    m = LogisticMatrixFactorization("IPS", epochs, 64, 9e-6, B, 12345, log.n_users, log.n_items, 1e-4)
    t0 = time.perf_counter(); m.fit(log.mf_train, log.mf_val); dt = time.perf_counter() - t0
    print("MF B=%d: %.2f ms/epoch, %.2f M interactions/s (incl. upload, legacy sampler)" % (B, dt * 1e3 / epochs, epochs * B / dt / 1e6), m.last_fit_stats)
