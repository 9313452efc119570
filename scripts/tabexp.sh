for cfg in "384 640" "384 384" "128 384" "128 256" "64 192" "384 1024"; do
set -- $cfg
MRC_CHAIN_TABLE_LO=$1 MRC_CHAIN_TABLE_HI=$2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-sequential-sample 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 $2', round(d['value']), round(d['e2e']['value']), {k:round(v,1) for k,v in d['stage_ms_per_step'].items() if k!='note'}, d['executed_work']['chain_iters'])
"
done
