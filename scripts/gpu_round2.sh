#!/bin/bash
# bench (1 and 2 streams, batch 2 and 4) + full ncu capture of the m2g InteractionNet kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_trainer.py -q -m gpu -x --timeout 300 > gpurun_out/test_gpu_trainer.log 2>&1; echo "trainer tests exit=$?"; tail -3 gpurun_out/test_gpu_trainer.log
for cfg in "1 2" "2 2" "1 4" "2 4"; do
  set -- $cfg
  timeout 600 python bench.py --steps 10 --warmup 3 --streams $1 --batch-per-gpu $2 --no-cpu-baseline > gpurun_out/bench_s$1_b$2.log 2> gpurun_out/bench_s$1_b$2.err
  echo "bench streams=$1 batch=$2 exit=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_s$1_b$2.log").read().strip().splitlines()[-1])
    print("  value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("  parse failed", e)
PY
done
NCU_KERNELS=mlp2_tc NCU_S=4 NCU_C=4 NCU_OUT=prof_m2g_r01e bash scripts/gpu_ncu_full.sh
