#!/usr/bin/env python
"""ncu --page raw --csv of tools/ncu_two_level_probe.py -> profiles/r2y_two_level_full.txt (per-launch summary of the
LAST captured epoch) and the "kuairec_two_level" entry of profiles/ncu_traffic.json (bench.py's roofline.traffic).
    python profiles/two_level_traffic.py gpurun_out/r2y_tl_raw.csv"""
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def short(name):
    """kernel function name -> the name bench.py's per-kernel profile uses"""
    base = name.split("(")[0]
    args = re.findall(r"<(.*)>", base)
    targs = [a.strip() for a in args[0].split(",")] if args else []
    if "fm_vrows_kernel" in base:
        return "fm_vrows_train" if targs[3] == "0" else "fm_vrows_loss"
    if "fm_cols_kernel" in base:
        return "fm_cols_level2" if targs[4] in ("1", "true") else "fm_cols_level1"
    if "fm_fixup_kernel" in base:
        return "fm_fixup_level1" if targs[2] == "2" else "fm_fixup_level2"
    if "fm_entity_fwd" in base:
        return "fm_entity_fwd"
    if "rs_onesweep" in base:
        return "rs_onesweep_kernel<T>"
    return base


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]

    def g(r, h):
        try:
            return float(r[hdr.index(h)].replace(",", ""))
        except (ValueError, IndexError):
            return float("nan")
    unit = rows[1][hdr.index("dram__bytes_read.sum")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    launches = [(short(r[hdr.index("Kernel Name")]), r) for r in rows[2:]]
    # the last epoch: from the last fm_vrows_train to the end, preceded by the entity forward that feeds it
    last = max(i for i, (n, _) in enumerate(launches) if n == "fm_vrows_train")
    epoch = launches[last:]
    out, lines = {}, ["# ncu --set full --clock-control none -k regex:fm_|rs_onesweep, python tools/ncu_two_level_probe.py "
                      "(3 M KuaiRec-shaped rows, B = 65,536, k = 64, factored rows, two-level step); last captured epoch, "
                      "one launch per line, cold cache, serialised: durations for attribution only"]
    for name, r in epoch:
        dram = (g(r, "dram__bytes_read.sum") + g(r, "dram__bytes_write.sum")) * scale
        e = out.setdefault(name, {"launches_per_step": 0, "dram_bytes": 0.0, "duration_us": 0.0})
        e["launches_per_step"] += 1
        e["dram_bytes"] += dram
        e["duration_us"] += g(r, "gpu__time_duration.sum")
        lines.append("%-22s t=%7.1f us  dram=%10.0f B  dram%%=%5.1f lts%%=%5.1f issue%%=%5.1f warps%%=%5.1f sm_active%%=%5.1f "
                     "l2hit=%5.1f regs=%3d grid=%4d  long_sb=%5.2f barrier=%5.2f" % (
                         name, g(r, "gpu__time_duration.sum"), dram,
                         g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                         g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                         100.0 * g(r, "sm__cycles_active.avg") / max(g(r, "sm__cycles_elapsed.max"), 1.0),
                         g(r, "lts__t_sector_hit_rate.pct"), g(r, "launch__registers_per_thread"), g(r, "launch__grid_size"),
                         g(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                         g(r, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")))
    for e in out.values():
        e["dram_bytes_per_step"] = e["dram_bytes"]
        e["dram_bytes"] = e["dram_bytes"] / e["launches_per_step"]
    total = sum(e["dram_bytes_per_step"] for e in out.values())
    lines.append("# step total: %.1f MB of DRAM traffic, %.1f us under ncu" % (total / 1e6, sum(e["duration_us"] for e in out.values())))
    open(os.path.join(HERE, "r2y_two_level_full.txt"), "w").write("\n".join(lines) + "\n")
    tpath = os.path.join(HERE, "ncu_traffic.json")
    tj = json.load(open(tpath))
    tj["kuairec_two_level"] = out
    tj["source_two_level"] = "profiles/r2y_two_level_full.txt (profiles/two_level_traffic.py)"
    json.dump(tj, open(tpath, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1])
