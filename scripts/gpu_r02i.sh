#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
for P in fp32 fp64; do
MRC_TIMELINE=1 timeout 600 python bench.py --precision $P --steps 3 --warmup 2 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2> gpurun_out/${TAG}_timeline_$P.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$P', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
done
