#!/bin/bash
# One line per BASELINE config that is not the bench default (profiles/<tag>_configs.txt): configs[1] GraphLAM
# fp32 scaled_ar, configs[3] HiLAM T=12 (one GPU here; the multi-GPU line comes from torchrun), configs[4]
# conv path 1024x1024, plus the HBM-bound element-wise kernels at both grids.
cd "$(dirname "$0")/.."
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_configs.txt
: > $OUT
echo "# configs[1]: GraphLAM fp32 scaled_ar T=4, batch 4" >> $OUT
timeout 900 python bench.py --model GraphLAM --precision 32-true --strategy scaled_ar --batch-per-gpu 4 --steps 5 --no-cpu-baseline --no-extras >> $OUT 2>gpurun_out/${TAG}_configs.err
echo "# configs[3] (one GPU): HiLAM bf16 diff_ar T=12, batch 5" >> $OUT
timeout 900 python bench.py --pred-steps 12 --steps 4 --no-cpu-baseline --no-extras >> $OUT 2>>gpurun_out/${TAG}_configs.err
echo "# configs[4] (one GPU): conv stand-in 1024x1024 scaled_ar T=4 bf16" >> $OUT
timeout 600 python scripts/conv_path_bench.py --grid 1024x1024 --batch 2 >> $OUT 2>>gpurun_out/${TAG}_configs.err
echo "# configs[0]-shaped: conv stand-in 64x64, fp32, batch 2, T=1" >> $OUT
timeout 600 python scripts/conv_path_bench.py --grid 64x64 --batch 2 --pred-steps 1 --precision 32-true >> $OUT 2>>gpurun_out/${TAG}_configs.err
echo "# element-wise kernels" >> $OUT
timeout 600 python scripts/elementwise_micro.py --batch 16 >> $OUT 2>>gpurun_out/${TAG}_configs.err
timeout 600 python scripts/elementwise_micro.py --batch 8 --grid 1024x1024 >> $OUT 2>>gpurun_out/${TAG}_configs.err
cat $OUT | cut -c1-400
tail -5 gpurun_out/${TAG}_configs.err
