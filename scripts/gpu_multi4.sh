#!/bin/bash
# strong-scaling line at N GPUs with the reservoir relayed over gloo (default) and over NCCL point-to-point
N=${1:-8}
TAG=${2:-r03o}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for R in gloo nccl; do
MRC_SHARD_RELAY=$R MRC_TIMELINE=1 timeout 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_${R}_n${N}.json 2> gpurun_out/${TAG}_bench_strong_${R}_n${N}.err
echo "strong $R rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_strong_${R}_n${N}.json').read().strip().splitlines()[-1]); print('strong $R', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"
grep -E "^wave" gpurun_out/${TAG}_bench_strong_${R}_n${N}.err | tail -8
done
