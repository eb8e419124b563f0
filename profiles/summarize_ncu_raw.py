"""Short metric summary of an ncu --set full capture (the *_summary.csv files; bench.py reads dram bytes from them).

    ncu -i gpurun_out/<report>.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu_raw.py /tmp/raw.csv "<comment line>" > profiles/ncu_<name>_r02_summary.csv
"""
import csv
import sys

KEEP = ["derived__memory_l1_wavefronts_shared_excessive", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.max",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
units, vals = rows[i0 + 1], rows[i0 + 2]
print("# " + sys.argv[2])
print("metric,unit,value")
for k in sorted(KEEP):
    if k in hdr:
        j = hdr.index(k)
        print("%s,%s,%s" % (k, units[j], vals[j].replace(",", "")))
