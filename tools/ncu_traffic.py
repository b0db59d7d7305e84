#!/usr/bin/env python3
"""profiles/k_physics_traffic.json and profiles/k_sense_issue.json from the per-kernel summaries of the full ncu captures
(tools/ncu_summary.py). bench.py scales these per-arena / per-robot figures to the launch it timed.

  python tools/ncu_traffic.py steady=profiles/r02_ncu_full_summary.csv early=profiles/r02_ncu_full_early_summary.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHYS = ("k_phys_pre", "k_island_big", "k_island_warp", "k_island_batch", "k_phys_post")


def main():
    by_regime, issue = {}, None
    for arg in sys.argv[1:]:
        regime, path = arg.split("=", 1)
        rows = {r["kernel"]: r for r in csv.DictReader(open(path))}
        arenas = int(rows["k_phys_pre"]["grid"].strip("()").split(",")[0])
        mb = sum(float(rows[k]["dram_read_MB"]) + float(rows[k]["dram_write_MB"]) for k in PHYS if k in rows)
        by_regime[regime] = {
            "dram_bytes_per_arena_tick": mb * 1e6 / arenas, "arenas_in_capture": arenas,
            "source": "round 2: %s (ncu --set full, PS_GROUPS=1, %d arenas, %s): dram__bytes_read.sum + dram__bytes_write.sum of %s of one tick"
                      % (os.path.relpath(path, ROOT), arenas, "steady state after 300 ticks" if regime == "steady" else "ticks 36-39 after the reset", " + ".join(PHYS)),
        }
        if regime == "steady":
            robots = int(rows["k_sense"]["grid"].strip("()").split(",")[0])
            issue = {"source": "%s (ncu --set full, %d arenas, steady state after 300 ticks, PS_GROUPS=1)" % (os.path.relpath(path, ROOT), arenas)}
            for k in ("k_sense", "k_localmap"):
                issue[k] = {"warp_instructions_per_robot_sense": float(rows[k]["warp_inst"]) / robots,
                            "inst_issued_pct_of_peak": float(rows[k]["inst_issued_pct_of_peak_active"]),
                            "threads_per_instruction": float(rows[k]["threads_per_inst"]), "warps_active_pct": float(rows[k]["warps_active_pct"]),
                            "barrier_stall_per_issue": float(rows[k]["barrier_stall_per_issue"])}
            issue["note"] = ("the camera stage's roof is its instruction count (executed warp instructions per robot-sense at the issue rate above), "
                             "not shared-memory bandwidth")
    json.dump({"note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full) of the World.step kernels of one tick per arena, by regime of the "
                       "run; bench.py scales the entry of the regime it timed (early = ticks < 300 after the reset, steady = after >= 300 warm-up ticks)",
               "by_regime": by_regime}, open(os.path.join(ROOT, "profiles", "k_physics_traffic.json"), "w"), indent=1)
    if issue:
        json.dump(issue, open(os.path.join(ROOT, "profiles", "k_sense_issue.json"), "w"), indent=1)
    print(json.dumps(by_regime, indent=1))


if __name__ == "__main__":
    main()
