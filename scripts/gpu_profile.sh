#!/bin/bash
# Launch list (per-kernel device time, cold & serialised: compare SHARES) of ONE eager training
# step (bench.py --profile-step brackets it with cudaProfilerStart/Stop), plus optional full ncu
# captures.  Usage: gpu_profile.sh [tag] [kernel-regex ...]   (each regex -> prof_<tag>_<i>.ncu-rep)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}; shift
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control ${NCU_CACHE:-all} --profile-from-start off --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --profile-step ${BENCH_ARGS:-} \
  > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu launches exit=$?"
python scripts/summarize_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_summary_$TAG.txt 2>&1
head -50 gpurun_out/launches_summary_$TAG.txt
i=0
for rx in "$@"; do
  timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx \
    -s ${FULL_SKIP:-2} -c ${FULL_COUNT:-2} -o gpurun_out/prof_${TAG}_$i -f python bench.py --profile-step ${BENCH_ARGS:-} \
    > gpurun_out/ncu_full_${TAG}_$i.log 2>&1
  echo "ncu full $rx exit=$?"
  i=$((i+1))
done
ls -la gpurun_out/
