#!/bin/bash
# One ncu --set full capture of selected kernels of one warm resident step, summarised on the GPU box so that only
# text comes back.  usage: bash profiles/ncu_kernel.sh <tag> <kernel regex> [launch-skip] [launch-count]
set -u
TAG=$1; KRE=$2; SKIP=${3:-0}; CNT=${4:-4}
OUT=gpurun_out
CMD="python bench.py --profile --batch 64 --steps 2 --warmup 3 ${BENCH_ARGS:-}"
export LFD_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_$TAG.log 2>&1
echo "capture rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page details --csv > $OUT/${TAG}_details.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --print-source cuda,sass --csv > $OUT/${TAG}_source.csv 2>/dev/null
python profiles/source_hotspots.py $OUT/${TAG}_source.csv 40 > $OUT/${TAG}_hotspots.txt 2>&1
python profiles/source_hotspots.py $OUT/${TAG}_source.csv 150 ins > $OUT/${TAG}_hot_ins.txt 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/${TAG}_raw.csv")))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_blocks", "launch__occupancy_limit_warps",
        "sm__maximum_warps_per_active_cycle_pct", "launch__shared_mem_config_size", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]
with open("$OUT/${TAG}_summary.txt", "w") as f:
    for r in data:
        f.write("== " + r[hdr.index("Kernel Name")][:90] + "\n")
        for k in keep[1:]:
            if k in hdr:
                f.write("  %-88s %s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
PY
rm -f $OUT/${TAG}_source.csv
ls -la $OUT/prof_$TAG.ncu-rep
rm -f $OUT/prof_$TAG.ncu-rep
