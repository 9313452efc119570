#!/bin/bash
# round 2, first GPU call: parity tests incl. the new census tests, the full census, a baseline bench
TAG=${1:-r02a}
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_nproc.txt
MRC_FULLSIZE_MINUTES=${MRC_FULLSIZE_MINUTES:-10} timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python scripts/parity_census.py --material stream --seconds 3600 --random-windows 800 --out gpurun_out/${TAG}_census.json 2> gpurun_out/${TAG}_census.err
timeout 600 python scripts/parity_census.py --material music --seconds 600 --random-windows 300 --window-blocks 16 --out gpurun_out/${TAG}_census.json 2>> gpurun_out/${TAG}_census.err
timeout 600 python scripts/parity_census.py --material batch --seconds 1800 --random-windows 600 --window-blocks 10 --out gpurun_out/${TAG}_census.json 2>> gpurun_out/${TAG}_census.err
timeout 600 python scripts/parity_census.py --material stream --seconds 600 --random-windows 200 --independent --out gpurun_out/${TAG}_census.json 2>> gpurun_out/${TAG}_census.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 1500 gpurun_out/${TAG}_bench.json
