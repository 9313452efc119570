#!/bin/bash
# bench lines of the build at HEAD with the refreshed analysis counters (scripts/gpu_last.sh made the capture)
TAG=${1:-r04final}
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 60 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
el "bench rc=$?"
timeout 30 python bench.py --steps 3 --warmup 3 --block-switching --no-cpu-baseline --no-sequential-sample --no-music > gpurun_out/${TAG}_bench_switching.json 2>> gpurun_out/${TAG}_bench.err
el "switching rc=$?"
timeout 30 python bench.py --steps 3 --warmup 3 --workload batch --no-cpu-baseline --no-sequential-sample --no-music > gpurun_out/${TAG}_bench_batch.json 2>> gpurun_out/${TAG}_bench.err
el "batch rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('fp64', d['value'], d['e2e']['value'], d['decode']['e2e_value'], d['music']['value'], d['roofline']['frac'], d['roofline']['frac_pipe_slots'])
for f in ('switching','batch'):
    d=json.load(open('gpurun_out/${TAG}_bench_%s.json' % f)); print(f, d['value'], d['e2e']['value'], d.get('decode',{}).get('e2e_value'))"
