#!/bin/bash
# Batch / stream sweep of the bench on one B200 (no CPU arm): which per-GPU batch amortises the
# launch-bound mesh-level kernels best.  Usage: gpu_round8.sh [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01f}
rm -f gpurun_out/sweep_$TAG.txt
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/sweep_${TAG}_$name.log 2> gpurun_out/sweep_${TAG}_$name.err
  echo "$name exit=$? args: $*" >> gpurun_out/sweep_$TAG.txt
  tail -1 gpurun_out/sweep_${TAG}_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('  value %.2f samples/s  ms/step %.2f  e2e %.2f  roofline frac %.3f (%.3f ms)  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']))
except Exception as e:
    print('  no json line', e)
" >> gpurun_out/sweep_$TAG.txt
  tail -2 gpurun_out/sweep_${TAG}_$name.err >> gpurun_out/sweep_$TAG.txt
}
run b2 --batch-per-gpu 2
run b4 --batch-per-gpu 4
run b8 --batch-per-gpu 8
run b4s2 --batch-per-gpu 4 --streams 2
run b8s2 --batch-per-gpu 8 --streams 2
run b2t12 --batch-per-gpu 2 --pred-steps 12
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/sweep_$TAG.txt
cat gpurun_out/sweep_$TAG.txt
