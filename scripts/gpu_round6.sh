#!/bin/bash
# bwd-kernel iteration: parity tests that exercise the backward, m2g micro timings, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in test_gpu_kernels test_gpu_tc_features test_gpu_models test_gpu_trainer; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
python scripts/kernel_micro.py --iters 4 --batch 2
python scripts/kernel_micro.py --iters 4 --batch 2 --which g2m
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"
  tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','gpu_launches')})"
  tail -5 gpurun_out/bench.err
fi
