#!/bin/bash
# ncu full capture (with source) of one analysis_kernel launch on a 120 s stream; usage: scripts/prof_analysis.sh <tag>
TAG=${1:-p}
CMD="python bench.py --steps 1 --warmup 1 --seconds 120 --no-cpu-baseline --no-sequential-sample"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:analysis_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_an $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
python bench.py --steps 3 --warmup 2 --seconds 600 --no-cpu-baseline --no-sequential-sample | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['stage_ms_per_step'], d['executed_work'])"
