#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): launch list + one full ncu capture of the dominant kernel
# of `bench.py`, per /opt/skills/guides/B200_PROFILING.md.  Outputs land in gpurun_out/.
#   tools/profile_gpu.sh <tag> <kernel-regex> [bench args...]
set -u
TAG=${1:-r1}; shift
KRE=${1:-search_kernel}; shift
OUT=gpurun_out
mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --no-gather $*"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py $ARGS > $OUT/${TAG}_launches.log 2>&1
# the dominant kernel, full set, second occurrence (after the warm-up step)
timeout 1200 ncu --set full --clock-control none --import-source on -k $KRE -s 1 -c 1 \
    -f -o $OUT/${TAG}_${KRE} python bench.py $ARGS > $OUT/${TAG}_ncu.log 2>&1
ls -la $OUT | tail -20
