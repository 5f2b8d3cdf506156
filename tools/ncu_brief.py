#!/usr/bin/env python
"""Key metrics of one ncu report (first kernel): python tools/ncu_brief.py REPORT.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
print(vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("  %-68s %s %s" % (w, vals[i], units[i]))
st = [(h, float(vals[i])) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and vals[i] not in ("", "n/a")]
for h, v in sorted(st, key=lambda x: -x[1])[:7]:
    print("  stall %-50s %.2f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
