#!/usr/bin/env python
"""Distil an `ncu --set full` report into the few lines the roofline discussion needs.

usage: python profiles/tools/ncu_summary.py report.ncu-rep > profiles/rNN_<kernel>_ncu.txt
(reads the report with `ncu -i ... --page raw --csv`; ncu must be on PATH)
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64", "smsp__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active",
    "dmma", "tensor", "smsp__issue_active.avg.pct", "smsp__warp_issue_stalled", "smsp__average_warp",
    "sm__sass_inst_executed_op_shared", "smsp__cycles_active.avg", "tma", "uniform",
]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    names, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(names, r)); u = dict(zip(names, units))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for n in names:
            if any(k in n for k in KEYS):
                v = d[n]
                if v in ("", "0", "n/a"): continue
                print(f"  {n:<95s} {v:>22s} {u[n]}")
        print()

if __name__ == "__main__":
    main(sys.argv[1])
