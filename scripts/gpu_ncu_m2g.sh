#!/bin/bash
# ncu --set full of the m2g InteractionNet kernels (fwd pipelined, edge/node bwd, seg_linear) at batch 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01d}
python scripts/kernel_micro.py --iters 3 --batch 2 > gpurun_out/micro_$TAG.log 2>&1
cat gpurun_out/micro_$TAG.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-mlp2_|seg_linear}" -s ${NCU_S:-6} -c ${NCU_C:-6} \
  -o gpurun_out/prof_m2g_$TAG -f python scripts/kernel_micro.py --iters 2 --batch 2 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/*.ncu-rep
