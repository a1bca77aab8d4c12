#!/bin/bash
# pipelined forward kernel: parity tests, A/B micro timings, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ws_forward.py -q -m gpu -x --timeout 300 > gpurun_out/test_gpu_ws_forward.log 2>&1; echo "ws tests exit=$?"; tail -15 gpurun_out/test_gpu_ws_forward.log
for w in m2g g2m; do
  echo "== $w monolithic"; P4C_WS_MIN_TILES=1000000000 timeout 300 python scripts/kernel_micro.py --which $w --iters 4 2>&1 | tail -2
  echo "== $w pipelined";  timeout 300 python scripts/kernel_micro.py --which $w --iters 4 2>&1 | tail -2
done
if [ "${SKIP_REST:-0}" != "1" ]; then
for f in test_gpu_kernels test_gpu_tc_features test_gpu_models test_gpu_trainer; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit=$?"; tail -2 gpurun_out/$f.log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ws.log 2> gpurun_out/bench_ws.err; echo "bench exit=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ws.log").read().strip().splitlines()[-1])
    print("  value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "roof", d["roofline"]["kernels"])
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/bench_ws.err").read()[-2000:])
PY
fi
