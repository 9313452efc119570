#!/bin/bash
# strong-scaling line at N GPUs (reservoir relayed over gloo)
N=${1:-2}
TAG=${2:-r03s}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n${N}.json 2> gpurun_out/${TAG}_bench_strong_n${N}.err
echo "strong rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_strong_n${N}.json').read().strip().splitlines()[-1]); print('strong', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
