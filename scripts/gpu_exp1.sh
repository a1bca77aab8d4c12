#!/bin/bash
# experiment: real (graph-replay) cost of the mesh processor (processor_layers 4 vs 0) and the
# phase traces of the small kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/trace_bwd.py > gpurun_out/trace_bwd.log 2>&1; tail -30 gpurun_out/trace_bwd.log
python scripts/trace_small.py > gpurun_out/trace_small.log 2>&1; tail -12 gpurun_out/trace_small.log
for pl in 4 0; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --processor-layers $pl > gpurun_out/bench_pl$pl.log 2>&1
  tail -1 gpurun_out/bench_pl$pl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pl=$pl', d['ms_per_step'], d['gpu_launches_per_step'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.log 2>&1
tail -1 gpurun_out/bench_nograph.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nograph', d['ms_per_step'])"
