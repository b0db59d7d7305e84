#!/usr/bin/env python3
"""One line per captured kernel from an ncu report (`ncu -i REP --page raw --csv`): the table kept under profiles/.

  python tools/ncu_summary.py gpurun_out/r02_prof.ncu-rep profiles/r02_ncu_full_raw.csv > profiles/r02_ncu_full_summary.csv
"""
import csv
import subprocess
import sys

COLS = [("kernel", "Kernel Name"), ("grid", "Grid Size"), ("regs", "launch__registers_per_thread"),
        ("dyn_smem_KB_per_block", "launch__shared_mem_per_block_dynamic"), ("static_smem_KB_per_block", "launch__shared_mem_per_block_static"),
        ("ms", "gpu__time_duration.sum"), ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("warp_inst", "smsp__inst_executed.sum"), ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ("inst_issued_pct_of_peak_active", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
        ("sm_throughput_pct_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("barrier_stall_per_issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("occ_limit_smem_blocks", "launch__occupancy_limit_shared_mem"), ("occ_limit_regs_blocks", "launch__occupancy_limit_registers")]


def main():
    rep, raw_out = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    if raw_out:
        open(raw_out, "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    w = csv.writer(sys.stdout)
    w.writerow([c for c, _ in COLS])
    for r in rows[2:]:
        out = []
        for c, h in COLS:
            v = r[idx[h]] if h in idx else ""
            if c == "kernel":
                v = v.split("(")[0].replace("ps::", "")
            out.append(v)
        w.writerow(out)


if __name__ == "__main__":
    main()
