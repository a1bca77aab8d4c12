#!/bin/bash
# Backward-kernel thread-shape A/B (P4C_BWD_TPR = 2 | 4): parity tests on the default shape, the m2g /
# g2m InteractionNet in isolation and the bench at batch 8 for both.  Usage: gpu_round9.sh [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01g}
OUT=gpurun_out/ab_$TAG.txt
rm -f $OUT
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  for f in ${TESTS:-test_gpu_kernels test_gpu_tc_features test_gpu_models test_gpu_trainer}; do
    timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
    echo "$f exit=$?" >> $OUT
    tail -3 gpurun_out/$f.log >> $OUT
  done
fi
for tpr in ${TPRS:-2 4}; do
  echo "== P4C_BWD_TPR=$tpr" >> $OUT
  P4C_BWD_TPR=$tpr python scripts/kernel_micro.py --iters 4 --batch 2 2>&1 | tail -2 >> $OUT
  P4C_BWD_TPR=$tpr python scripts/kernel_micro.py --iters 4 --batch 2 --which g2m 2>&1 | tail -2 >> $OUT
  if [ "${SKIP_BENCH:-0}" != "1" ]; then
    P4C_BWD_TPR=$tpr timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ab_${TAG}_tpr$tpr.log 2> gpurun_out/ab_${TAG}_tpr$tpr.err
    tail -1 gpurun_out/ab_${TAG}_tpr$tpr.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline())
    print('  bench: %.2f samples/s  ms/step %.2f  e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e:
    print('  no json line', e)
" >> $OUT
    tail -2 gpurun_out/ab_${TAG}_tpr$tpr.err >> $OUT
  fi
done
cat $OUT
