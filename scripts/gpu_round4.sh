#!/bin/bash
# 16-warp pipelined forward + PDL: parity tests, bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in test_gpu_ws_forward test_gpu_kernels test_gpu_tc_features test_gpu_ar_loss test_gpu_models test_gpu_trainer; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f exit=$?"; tail -2 gpurun_out/$f.log
done
for pdl in 0 1; do
  P4C_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.log 2> gpurun_out/bench_pdl$pdl.err; echo "bench pdl=$pdl exit=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_pdl$pdl.log").read().strip().splitlines()[-1])
    print("  value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "roof", d["roofline"]["kernels"])
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/bench_pdl$pdl.err").read()[-2000:])
PY
done
