#!/bin/bash
# quick loop: parity tests, a short bench line, phase clocks
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_switching.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-sequential-sample --no-cpu-baseline --no-decode --no-music > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('fp64', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
timeout 600 python bench.py --steps 3 --warmup 2 --precision fp32 --no-sequential-sample --no-cpu-baseline --no-decode --no-music > gpurun_out/${TAG}_bench_fp32.json 2>> gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_fp32.json')); print('fp32', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
timeout 300 python scripts/phase_clocks.py 120 > gpurun_out/${TAG}_phase_clocks.log 2>&1
head -30 gpurun_out/${TAG}_phase_clocks.log
