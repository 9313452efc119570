#!/bin/bash
TAG=${1:-r03q}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
MRC_FULLSIZE_MINUTES=10 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_switching.py tests/test_gpu_fullsize.py -m gpu -x -q -k "decode or roundtrip or malformed or golden or cli or tiny or one_hour or switching or seam or buffer or scouts" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-sequential-sample --no-cpu-baseline --no-music > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['e2e']['value'], d['decode'])"
