#!/bin/bash
# full evidence run for one build: whole GPU test suite (with the 10-minute census), the driver's bench line with default
# flags (timed), fp32 / batch / switching lines, the debug-assert build, phase clocks, ncu launch lists and captures
TAG=${1:-r02z}
mkdir -p gpurun_out
MRC_FULLSIZE_MINUTES=10 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$? in $(( $(date +%s) - T0 )) s"; head -c 300 gpurun_out/${TAG}_bench.json; echo
T0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
echo "reference rc=$? in $(( $(date +%s) - T0 )) s"; head -c 200 gpurun_out/${TAG}_bench_reference.json; echo
timeout 600 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload batch --no-cpu-baseline --no-sequential-sample --no-music > gpurun_out/${TAG}_bench_batch.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --block-switching --no-cpu-baseline > gpurun_out/${TAG}_bench_switching.json 2>> gpurun_out/${TAG}_bench.err
for f in fp32 batch switching; do python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_$f.json')); print('$f', d['value'], d['e2e']['value'], d.get('decode',{}).get('e2e_value'))"; done
timeout 300 python scripts/debug_asserts.py > gpurun_out/${TAG}_debug_asserts.log 2>&1; echo "debug asserts rc=$?"; tail -3 gpurun_out/${TAG}_debug_asserts.log
timeout 300 python scripts/phase_clocks.py 120 > gpurun_out/${TAG}_phase_clocks.log 2>&1
bash scripts/gpu_ncu.sh ${TAG} fp64
bash scripts/gpu_ncu.sh ${TAG}f fp32
