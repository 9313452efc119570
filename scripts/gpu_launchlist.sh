#!/bin/bash
# ncu launch list (durations + DRAM bytes) of the first 120 launches of a 1500 s encode
TAG=${1:-ll}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --seconds 1500 --no-cpu-baseline --no-sequential-sample --no-decode --no-music"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv, sys, os
tag = os.environ.get("TAG_", "")
PY
