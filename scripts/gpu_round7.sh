#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list of one step, ncu --set full of the
# m2g InteractionNet kernels (batch 2).  Usage: gpu_round7.sh [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01s}
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc > gpurun_out/host_cores.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  for f in test_gpu_kernels test_gpu_tc_features test_gpu_ws_forward test_gpu_ar_loss test_gpu_models test_gpu_trainer test_forcings; do
    timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
    echo "$f exit=$?" >> gpurun_out/summary.txt
    tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
  done
  timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/smoke.log >> gpurun_out/summary.txt
fi
python scripts/kernel_micro.py --iters 4 --batch ${MICRO_BATCH:-8} > gpurun_out/micro_$TAG.log 2>&1
python scripts/kernel_micro.py --iters 4 --batch ${MICRO_BATCH:-8} --which g2m >> gpurun_out/micro_$TAG.log 2>&1
cat gpurun_out/micro_$TAG.log >> gpurun_out/summary.txt
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err
  echo "bench exit=$?" >> gpurun_out/summary.txt
  tail -1 gpurun_out/bench_$TAG.log >> gpurun_out/summary.txt
  tail -3 gpurun_out/bench_$TAG.err >> gpurun_out/summary.txt
fi
if [ "${SKIP_PROFILE:-0}" != "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --profile-step ${BENCH_ARGS:-} \
    > gpurun_out/ncu_bench_$TAG.log 2>&1
  echo "ncu launches exit=$?" >> gpurun_out/summary.txt
  python scripts/summarize_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_summary_$TAG.txt 2>&1
  python scripts/launch_groups.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_groups_$TAG.txt 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-mlp2_|seg_linear}" \
    -s ${NCU_S:-6} -c ${NCU_C:-6} -o gpurun_out/prof_m2g_$TAG -f python scripts/kernel_micro.py --iters 2 --batch ${MICRO_BATCH:-8} \
    > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "ncu full exit=$?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
ls -la gpurun_out/
