"""Step time of the FM trainer at the benchmark shape (BASELINE configs[2]: 12 M rows, B = 65,536, k = 64) on the
stacked CSR, on factored rows with the flat step and with the two-level step (csrc/two_level.cuh); per-kernel times
from CUDA events around every launch. Usage: python tools/two_level_probe.py [--rows N] [--batch B] [--dtype float64]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "relevance-factorizationmachine_b200"))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12_000_000)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--factors", type=int, default=64)
    args = ap.parse_args()
    from rfm_b200._capi import check, lib, ptr
    from rfm_b200.fm import FactorizationMachines, _FmTrainer
    log, gen_s = bench.make_data(args.rows, 2024)
    ftrain, fval = bench.factored_dicts(log)
    B, K, W = args.batch, args.steps, 5
    out = {"rows": args.rows, "batch": B, "dtype": args.dtype, "k": args.factors}
    finals = {}
    for name, train_d, val_d, two in (("csr_flat", log.fm_train, log.fm_val, False), ("factored_flat", ftrain, fval, False),
                                      ("factored_two_level", ftrain, fval, True)):
        model = FactorizationMachines("IPS", K, args.factors, bench.LR, B, 12345, log.n_features, dtype=args.dtype,
                                      sampler="feistel")
        ctx = model._context()
        train_rows = model._rows(train_d["features"], train_d["labels"], train_d["pscores"])
        val_rows = model._rows(val_d["features"], val_d["labels"], val_d["pscores"])
        model.sync_to_device()
        trainer = _FmTrainer(model._dev, train_rows, val_rows, B, W + 2 * K + 8)
        if two:
            assert trainer.set_two_level(1)

        def stepper(epoch, slot):
            check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, epoch, B, bench.LR, slot))

        ms, launches, clk, prof = bench.timed_steps(ctx, None, stepper, W, K, 0, observe=False)
        tl, vl = np.empty(W + 2 * K), np.empty(W + 2 * K)
        check(lib().rfm_fm_trainer_losses(trainer.handle, 0, W + 2 * K, ptr(tl), ptr(vl)))
        trainer.close()
        model.sync_to_host()
        finals[name] = (tl.copy(), vl.copy(), model.V().copy())
        out[name] = {"ms_per_step": ms / K, "interactions_per_s": K * B / (ms * 1e-3), "launches_per_step": launches / K,
                     "kernels_us": {k: round(v[1] / K * 1e3, 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                     "final_train_loss": float(tl[W + 2 * K - 1]), "final_val_loss": float(vl[W + 2 * K - 1])}
        model.reset_rows_cache()
    a, b = finals["factored_flat"], finals["factored_two_level"]
    out["two_level_vs_flat"] = {"train_loss_max_rel": float(np.max(np.abs(a[0] - b[0]) / np.abs(a[0]))),
                                "val_loss_max_rel": float(np.max(np.abs(a[1] - b[1]) / np.abs(a[1]))),
                                "V_max_abs": float(np.max(np.abs(a[2] - b[2]))), "V_scale": float(np.max(np.abs(a[2])))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
