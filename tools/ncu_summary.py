#!/usr/bin/env python
"""Summary of an `ncu --set full` report for profiles/:  ncu -i X.ncu-rep --page raw --csv > raw.csv;
python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "sm__mio_inst_issued.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__warps_eligible.avg.per_cycle_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("----")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]} {units[i]}")
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
        print(f"   stall {name:40s} {v:.2f}")
