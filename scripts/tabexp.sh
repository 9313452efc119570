# chain-table range experiment: value / e2e / stage times on the 1 h stream and on a 600 s stream (where the serial walk binds)
for cfg in "128 640" "128 832" "128 1024" "64 960"; do
set -- $cfg
for secs in 3600 600; do
MRC_CHAIN_TABLE_LO=$1 MRC_CHAIN_TABLE_HI=$2 python bench.py --seconds $secs --steps 3 --warmup 2 --no-cpu-baseline --no-sequential-sample 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lo -$1 hi $2 seconds $secs:', round(d['value']), round(d['e2e']['value']), {k:round(v,1) for k,v in d['stage_ms_per_step'].items() if k!='note'}, d['executed_work']['chain_iters'])
"
done
done
