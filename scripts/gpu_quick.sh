#!/bin/bash
# quick GPU iteration: parity diagnostics, the GPU test suite, a short bench line
TAG=${1:-q}
mkdir -p gpurun_out
python scripts/gpu_check.py > gpurun_out/${TAG}_check.log 2>&1
grep -E "^==|mismatching|identical|rel err|Error|error" gpurun_out/${TAG}_check.log | head -60
MRC_FULLSIZE_MINUTES=${MRC_FULLSIZE_MINUTES:-10} timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | tee gpurun_out/${TAG}_bench.json
