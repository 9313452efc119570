#!/bin/bash
TAG=${1:-r02b}
mkdir -p gpurun_out
MRC_FULLSIZE_MINUTES=${MRC_FULLSIZE_MINUTES:-10} timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-sequential-sample --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['executed_work'])"
MRC_TIMELINE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-sequential-sample --no-cpu-baseline 2> gpurun_out/${TAG}_timeline.txt > /dev/null
timeout 600 python bench.py --steps 3 --warmup 3 --seconds 600 --no-sequential-sample --no-cpu-baseline > gpurun_out/${TAG}_bench600.json 2>> gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench600.json')); print('600s', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
