#!/bin/bash
# multi-GPU: byte check of the sharded stream (also without walking ahead), then the strong-scaling line with the timeline
N=${1:-2}
TAG=${2:-r03m}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN scripts/shard_check.py 600 > gpurun_out/${TAG}_shard_check_n${N}.json 2> gpurun_out/${TAG}_shard_check_n${N}.err
echo "shard check rc=$?"; cat gpurun_out/${TAG}_shard_check_n${N}.json; tail -3 gpurun_out/${TAG}_shard_check_n${N}.err
MRC_TIMELINE=1 timeout 600 $RUN bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n${N}.json 2> gpurun_out/${TAG}_bench_strong_n${N}.err
echo "strong rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_strong_n${N}.json').read().strip().splitlines()[-1]); print('strong', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"
grep "walked ahead" gpurun_out/${TAG}_bench_strong_n${N}.err | sort | uniq -c | head
MRC_SHARD_NO_SPECULATION=1 timeout 600 $RUN bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_nospec_n${N}.json 2> gpurun_out/${TAG}_bench_strong_nospec_n${N}.err
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_strong_nospec_n${N}.json').read().strip().splitlines()[-1]); print('strong, not walking ahead', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-decode --no-music --no-cpu-baseline --no-sequential-sample > gpurun_out/${TAG}_bench_weak_n${N}.json 2> gpurun_out/${TAG}_bench_weak_n${N}.err
echo "weak rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_weak_n${N}.json').read().strip().splitlines()[-1]); print('weak', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
