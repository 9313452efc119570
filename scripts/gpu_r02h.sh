#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "chain or shard or one_hour or batch_of or tiny or switching_at" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
MRC_TIMELINE=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2> gpurun_out/${TAG}_timeline.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
grep "serial pass" gpurun_out/${TAG}_timeline.txt | tail -1
tail -17 gpurun_out/${TAG}_timeline.txt | head -8
timeout 600 python bench.py --steps 3 --warmup 2 --seconds 600 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('600 s:', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
