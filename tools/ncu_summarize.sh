#!/bin/bash
# Summarises one `ncu --set full` capture the way profiles/*_ncu_summary.txt are written:
#   tools/ncu_summarize.sh gpurun_out/X.ncu-rep "title"  >> profiles/rN_*_ncu_summary.txt
rep=$1; title=${2:-$1}
echo "== $title"
ncu -i "$rep" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "launch__block_size", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-75s %-16s %s" % (w, units[i], vals[i]))
'
ncu -i "$rep" --page source --csv 2>/dev/null | python3 "$(dirname "$0")/ncu_source_summary.py" | head -34
