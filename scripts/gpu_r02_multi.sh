#!/bin/bash
# multi-GPU: byte check of the sharded stream, then strong and weak scaling bench lines at N GPUs
N=${1:-2}
TAG=${2:-r02m}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN scripts/shard_check.py 600 > gpurun_out/${TAG}_shard_check_n${N}.json 2> gpurun_out/${TAG}_shard_check_n${N}.err
echo "shard check rc=$?"; cat gpurun_out/${TAG}_shard_check_n${N}.json; tail -3 gpurun_out/${TAG}_shard_check_n${N}.err
timeout 600 $RUN bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n${N}.json 2> gpurun_out/${TAG}_bench_strong_n${N}.err
echo "strong rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_strong_n${N}.json')); print('strong', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms_per_step'])"
tail -3 gpurun_out/${TAG}_bench_strong_n${N}.err
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-decode --no-music --no-cpu-baseline --no-sequential-sample > gpurun_out/${TAG}_bench_weak_n${N}.json 2> gpurun_out/${TAG}_bench_weak_n${N}.err
echo "weak rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_weak_n${N}.json')); print('weak', d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'])"
