#!/bin/bash
# Launch list (per-kernel device time, cold & serialised: compare SHARES) of one training step,
# plus an optional full ncu capture of one kernel.  Usage: gpu_profile.sh [kernel-regex]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SKIP=${NCU_SKIP:-4600}
COUNT=${NCU_COUNT:-2400}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline ${BENCH_ARGS:-} \
  > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1
head -40 gpurun_out/launches_summary.txt
if [ -n "$1" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${FULL_SKIP:-20} -c ${FULL_COUNT:-2} \
    -o gpurun_out/prof_$2 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline ${BENCH_ARGS:-} \
    > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?"
fi
