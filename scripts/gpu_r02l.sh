#!/bin/bash
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for HI in 1024 2048; do
MRC_CHAIN_TABLE_HI=$HI MRC_TIMELINE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2> gpurun_out/${TAG}_timeline_$HI.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('hi=$HI', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
grep "serial pass" gpurun_out/${TAG}_timeline_$HI.txt | tail -1
done
timeout 300 python scripts/phase_clocks.py 120 > gpurun_out/${TAG}_phase_clocks.log 2>&1
cat gpurun_out/${TAG}_phase_clocks.log
timeout 300 python scripts/debug_asserts.py > gpurun_out/${TAG}_debug_asserts.log 2>&1; echo "debug asserts rc=$?"; tail -3 gpurun_out/${TAG}_debug_asserts.log
