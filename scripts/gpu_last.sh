#!/bin/bash
# Evidence of the build at HEAD inside what is left of the GPU budget, most valuable first: the driver's bench line, one
# --set full capture of a full-wave analysis launch (after the plain command exited 0), the whole GPU test suite, fp32.
TAG=${1:-r04final}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 100 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
el "bench rc=$?"; head -c 160 gpurun_out/${TAG}_bench.json; echo
CMD="python bench.py --steps 1 --warmup 0 --seconds 1500 --precision fp64 --no-cpu-baseline --no-sequential-sample --no-decode --no-music"
timeout 40 $CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1
el "plain rc=$?"
timeout 70 ncu --set full --clock-control none --import-source on -k regex:'analysis_kernel' -s 3 -c 1 -f -o gpurun_out/${TAG}_analysis_fp64 $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
el "analysis capture rc=$?"
timeout 170 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
el "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 40 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline --no-sequential-sample --no-music > gpurun_out/${TAG}_bench_fp32.json 2>> gpurun_out/${TAG}_bench.err
el "fp32 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('fp64', d['value'], d['e2e']['value'], d['decode']['e2e_value'], d['music']['value'], d['roofline']['frac'])
d=json.load(open('gpurun_out/${TAG}_bench_fp32.json')); print('fp32', d['value'], d['e2e']['value'])"
