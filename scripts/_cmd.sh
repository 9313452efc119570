timeout 600 python -m pytest tests/test_gpu_switching.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "one_hour" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sequential-sample > gpurun_out/r01z16_bench.json 2> gpurun_out/r01z16_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r01z16_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms_per_step'])
"
