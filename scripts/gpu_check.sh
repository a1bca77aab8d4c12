#!/bin/bash
# One GPU-box visit: parity tests (one process per file so a faulting kernel cannot poison the
# rest), smoke, a short bench.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc > gpurun_out/host_cores.txt
for f in $(cd tests && grep -l "mark.gpu" test_*.py | sed 's/\.py$//'); do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log >> gpurun_out/summary.txt
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
  tail -2 gpurun_out/bench.log >> gpurun_out/summary.txt
  tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
