#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_switching.py tests/test_huffman_train.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-sequential-sample --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['executed_work'], d.get('decode'), d.get('music'), d['roofline'])"
timeout 300 python scripts/phase_clocks.py 120 > gpurun_out/${TAG}_phase_clocks.log 2>&1
cat gpurun_out/${TAG}_phase_clocks.log
