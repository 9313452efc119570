#!/bin/bash
for B in 1 2 3 4 16; do
MRC_BLOCKS_PER_CTA=$B timeout 600 python bench.py --steps 3 --warmup 2 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B=$B fp64', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
done
for B in 1 2 4; do
MRC_BLOCKS_PER_CTA=$B timeout 600 python bench.py --steps 3 --warmup 2 --precision fp32 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B=$B fp32', d['value'], d['e2e']['value'], d['stage_ms_per_step'])"
done
