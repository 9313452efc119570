#!/bin/bash
# multi-GPU: byte check of the sharded stream, strong-scaling line (with the timeline)
N=${1:-2}
TAG=${2:-r03n}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN scripts/shard_check.py 3000 > gpurun_out/${TAG}_shard_check_n${N}.json 2> gpurun_out/${TAG}_shard_check_n${N}.err
echo "shard check rc=$?"; tail -1 gpurun_out/${TAG}_shard_check_n${N}.json | cut -c1-300; tail -3 gpurun_out/${TAG}_shard_check_n${N}.err
MRC_TIMELINE=1 timeout 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n${N}.json 2> gpurun_out/${TAG}_bench_strong_n${N}.err
echo "strong rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_strong_n${N}.json').read().strip().splitlines()[-1]); print('strong', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"
grep "walked ahead" gpurun_out/${TAG}_bench_strong_n${N}.err | sort | uniq -c | head
grep -E "^wave" gpurun_out/${TAG}_bench_strong_n${N}.err | tail -8
