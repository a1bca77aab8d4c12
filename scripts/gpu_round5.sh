#!/bin/bash
# HEAD validation: parity tests, smoke, bench (with CPU arm), launch list of one eager step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
SKIP_BENCH=1 bash scripts/gpu_check.sh
timeout 300 python -m pytest tests/test_gpu_ws_forward.py -q -m gpu -x --timeout 280 > gpurun_out/test_gpu_ws_forward.log 2>&1; echo "ws exit=$?"; tail -2 gpurun_out/test_gpu_ws_forward.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"
tail -1 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
bash scripts/gpu_profile.sh r01d > gpurun_out/profile_r01d.log 2>&1
head -30 gpurun_out/launches_summary_r01d.txt
