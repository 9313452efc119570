#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_switching.py tests/test_huffman_train.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for P in fp64 fp32; do
timeout 600 python bench.py --steps 3 --warmup 2 --precision $P --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2>gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_$P.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$P', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
done
