#!/bin/bash
# quick loop + one full ncu capture of a full-wave analysis launch
TAG=${1:-q}
bash scripts/gpu_quick2.sh $TAG
CMD="python bench.py --steps 1 --warmup 0 --seconds 1500 --no-cpu-baseline --no-sequential-sample --no-decode --no-music"
ncu --set full --clock-control none --import-source on -k regex:'analysis_kernel' -s 3 -c 1 -f -o gpurun_out/${TAG}_analysis_fp64 $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "analysis capture rc=$?"
