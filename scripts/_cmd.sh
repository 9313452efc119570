timeout 500 python -m pytest tests/test_gpu_switching.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sequential-sample > gpurun_out/r01z14_bench.json 2> gpurun_out/r01z14_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r01z14_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['executed_work']['chain_iters'])
"
python scripts/phase_clocks.py 120 > gpurun_out/r01z_phase_clocks11.log 2>&1; head -14 gpurun_out/r01z_phase_clocks11.log | tail -9; tail -7 gpurun_out/r01z_phase_clocks11.log
