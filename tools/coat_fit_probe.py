import sys, time
sys.path.insert(0, "relevance-factorizationmachine_b200"); sys.path.insert(0, ".")
import numpy as np
from rfm_b200.mf import LogisticMatrixFactorization
from rfm_b200.fm import FactorizationMachines
from rfm_b200.synth import make_coat_shaped
log = make_coat_shaped(seed=1)
print("coat-shaped: train rows", log.mf_train["features"].shape, "n_features", log.n_features)
for epochs in (100, 400):
    m = LogisticMatrixFactorization("IPS", epochs, 16, 0.01, 500, 12345, log.n_users, log.n_items, 1e-4)
    t0 = time.perf_counter(); m.fit(log.mf_train, log.mf_val); dt = time.perf_counter() - t0
    print("MF  epochs", epochs, "fit %.1f ms  (%.1f us/epoch)" % (dt * 1e3, dt * 1e6 / epochs), m.last_fit_stats)
    f = FactorizationMachines("IPS", epochs, 32, 0.001, 500, 12345, log.n_features, alpha=0.1)
    t0 = time.perf_counter(); f.fit(log.fm_train, log.fm_val); dt = time.perf_counter() - t0
    print("FM  epochs", epochs, "fit %.1f ms  (%.1f us/epoch)" % (dt * 1e3, dt * 1e6 / epochs), f.last_fit_stats.get("phase_seconds"))
