#!/bin/bash
# full GPU test suite, bench line, phase clocks, then the ncu evidence (launch list + full captures) of this build
TAG=${1:-r02p}
mkdir -p gpurun_out
MRC_FULLSIZE_MINUTES=10 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench.json; echo
timeout 300 python scripts/phase_clocks.py 120 > gpurun_out/${TAG}_phase_clocks.log 2>&1
cat gpurun_out/${TAG}_phase_clocks.log
bash scripts/gpu_ncu.sh ${TAG} fp64
