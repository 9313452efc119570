#!/bin/bash
# ncu evidence for one build (B200_PROFILING.md recipe): (1) the plain command exits 0, (2) launch list with durations and
# DRAM bytes of every kernel across the first full wave, (3) one --set full capture of the kernels of a full wave.
# usage: scripts/gpu_ncu.sh <tag> [precision]
TAG=${1:-r02}
PREC=${2:-fp64}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --seconds 1500 --precision $PREC --no-cpu-baseline --no-sequential-sample --no-decode --no-music"
$CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_ncu_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches_${PREC}.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# waves of a 1500 s call: 2048, 4096, 8192, 16384 (full), ...: the 4th analysis launch is a full wave
ncu --set full --clock-control none --import-source on -k regex:'analysis_kernel' -s 3 -c 1 -f -o gpurun_out/${TAG}_analysis_${PREC} $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "analysis capture rc=$?"
if [ "$PREC" = fp64 ]; then
ncu --set full --clock-control none --import-source on -k regex:'cost_kernel|table_kernel|segment_kernel|extras_kernel|chain_seg_kernel|expand_kernel|finish_kernel|offsets_kernel|pack_kernel' -s 27 -c 9 -f -o gpurun_out/${TAG}_others $CMD > gpurun_out/${TAG}_ncu_full2.log 2>&1
echo "others capture rc=$?"
fi
ls -la gpurun_out/ | grep ${TAG}
