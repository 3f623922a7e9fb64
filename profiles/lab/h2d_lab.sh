#!/bin/bash
# Drives h2d_lab: for N in the given list (default "1 2 4 8", capped by the visible GPUs) and every variant, N processes
# start their copy loops at a common epoch; prints one table row per (N, variant): per-rank min / mean and aggregate GB/s.
# usage (GPU box): bash profiles/lab/h2d_lab.sh [seconds] [N list]     -> gpurun_out/h2d_lab.txt
set -u
SECS=${1:-2}; NLIST=${2:-"1 2 4 8"}
# optional: DEVS="0 2 4 6" runs ONE row set on exactly these devices instead of devices 0..N-1 (topology probing)
HERE=$(dirname "$0")
[ -x $HERE/h2d_lab ] || nvcc -O2 -gencode arch=compute_100a,code=sm_100a $HERE/h2d_lab.cu -o $HERE/h2d_lab || exit 1
NGPU=$(nvidia-smi -L | wc -l)
echo "# host: $(nproc) CPUs, $(grep -c '^processor' /proc/cpuinfo) processors, NUMA nodes: $(ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l), hugepages: $(cat /proc/sys/vm/nr_hugepages 2>/dev/null), GPUs: $NGPU"
echo "# N variant how | per-rank min mean | aggregate GB/s"
if [ -n "${DEVS:-}" ]; then NLIST=$(echo $DEVS | wc -w); fi
for N in $NLIST; do
  [ $N -le $NGPU ] || continue
  for V in ${VARIANTS:-a b c d e f}; do
    START=$(python3 -c "import time; print(time.time() + 4.0)")
    rm -f /tmp/h2d_lab_$$.*
    if [ -n "${DEVS:-}" ]; then DL="$DEVS"; else DL=$(seq 0 $((N - 1))); fi
    for i in $DL; do $HERE/h2d_lab $i $V $START $SECS > /tmp/h2d_lab_$$.$i 2>&1 & done
    wait
    cat /tmp/h2d_lab_$$.* | python3 -c "
import sys, json
rows = [json.loads(l) for l in sys.stdin if l.startswith('{')]
g = [r['gbs'] for r in rows]
print('%d %s %-22s | %6.1f %6.1f | %7.1f  %s' % ($N, '$V', rows[0]['how'] if rows else 'FAILED', min(g) if g else 0, sum(g) / max(len(g), 1), sum(g), 'devices ${DEVS:-}' if '${DEVS:-}' else ''))
"
  done
done
rm -f /tmp/h2d_lab_$$.*
