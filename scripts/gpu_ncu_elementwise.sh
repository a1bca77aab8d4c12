#!/bin/bash
# ncu --set full of the HBM-bound kernels either side of the training path: forcing assembly,
# PSD, feature statistics, un-normalisation.  One short ncu run per kernel family -> gpurun_out/<tag>_ew_*.ncu-rep
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
for spec in "forcing_assemble:1" "psd_:3" "feature_stats:2" "scale_shift:1"; do
  k=${spec%%:*}; c=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$k" -c $c \
    -o gpurun_out/${TAG}_ew_${k%_} -f python scripts/elementwise_micro.py --batch 16 > gpurun_out/${TAG}_ew_${k%_}.log 2>&1
  echo "$k exit=$?"
done
ls -la gpurun_out/${TAG}_ew_*.ncu-rep
