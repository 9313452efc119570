#!/bin/bash
# ncu evidence for the decode path: launch durations and one full capture of decode_kernel / ola_kernel
TAG=${1:-r02dec}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --seconds 600 --no-cpu-baseline --no-sequential-sample --no-music"
$CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'parse_kernel|decode_kernel|ola_kernel' -c 60 --csv \
    --log-file gpurun_out/${TAG}_launches_decode.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'parse_kernel|decode_kernel|ola_kernel' -s 3 -c 3 -f -o gpurun_out/${TAG}_decode $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "decode capture rc=$?"
