#!/bin/bash
# ncu launch list of the profiled bench command (run on the GPU box through gpurun): the plain run first, and only
# when it exited 0 the same command under `ncu --metrics gpu__time_duration.sum --clock-control none`; the per-kernel
# shares next to the CUDA-event stage times of the plain run go to profiles/<tag>_launch_shares.md.
# usage: bash profiles/launch_list.sh <tag> [batch]
set -u
TAG=${1:-r02x}; B=${2:-64}
OUT=gpurun_out
CMD="python bench.py --profile --batch $B --steps 2 --warmup 3"
$CMD > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
python profiles/summarize.py $TAG $OUT/${TAG}_launches.csv - $OUT/plain_$TAG.json $B && cp profiles/${TAG}_launch_shares.md $OUT/
