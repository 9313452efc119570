#!/bin/bash
# last pass of a round: whole GPU test suite at its defaults, then the bench lines
TAG=${1:-r03f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 200 gpurun_out/${TAG}_bench.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload batch --no-cpu-baseline --no-sequential-sample --no-music > gpurun_out/${TAG}_bench_batch.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --block-switching --no-cpu-baseline > gpurun_out/${TAG}_bench_switching.json 2>> gpurun_out/${TAG}_bench.err
for f in fp32 batch switching; do python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_$f.json')); print('$f', d['value'], d['e2e']['value'], d.get('decode',{}).get('e2e_value'))"; done
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('fp64', d['value'], d['e2e']['value'], d['decode']['e2e_value'], d['music']['value'], d['roofline']['frac'], d['roofline']['frac_pipe_slots'])"
