#!/bin/bash
# Full ncu capture of the message-passing kernels of the Titan m2g InteractionNet.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/kernel_micro.py --iters 3 > gpurun_out/micro.log 2>&1
cat gpurun_out/micro.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-mlp2_tc}" -s ${NCU_S:-4} -c ${NCU_C:-4} \
  -o gpurun_out/${NCU_OUT:-prof_m2g} -f python scripts/kernel_micro.py --iters 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
