#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, then the ncu launch list and one full capture of the top
# kernels (B200_PROFILING.md recipe).  Everything lands in gpurun_out/.  Usage: scripts/gpu_round.sh <tag> [what]
#   what = all | tests | bench | ncu   (default all)
TAG=${1:-r01}
WHAT=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
RC=0
if [ "$WHAT" = all ] || [ "$WHAT" = tests ]; then
    MRC_FULLSIZE_MINUTES=${MRC_FULLSIZE_MINUTES:-10} timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
    RC=$?
    tail -15 gpurun_out/${TAG}_pytest.log
fi
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
    timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
    echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
    timeout 600 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline --no-sequential-sample > gpurun_out/${TAG}_bench_fp32.json 2>> gpurun_out/${TAG}_bench.err
    timeout 600 python bench.py --steps 3 --warmup 3 --workload batch --decode --no-cpu-baseline --no-sequential-sample > gpurun_out/${TAG}_bench_batch.json 2>> gpurun_out/${TAG}_bench.err
    timeout 600 python bench.py --steps 3 --warmup 3 --block-switching --decode --no-cpu-baseline > gpurun_out/${TAG}_bench_switching.json 2>> gpurun_out/${TAG}_bench.err
    timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
    tail -c 600 gpurun_out/${TAG}_bench_reference.json
fi
if [ "$WHAT" = all ] || [ "$WHAT" = ncu ]; then
    CMD="python bench.py --steps 2 --warmup 1 --seconds 600 --no-cpu-baseline --no-sequential-sample"
    $CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
    echo "launch list rc=$?"
    CMD2="python bench.py --steps 1 --warmup 1 --seconds 120 --no-cpu-baseline --no-sequential-sample"
    $CMD2 > gpurun_out/${TAG}_ncu_plain2.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'analysis_kernel|cost_kernel|table_kernel|chain_kernel|chain_table_kernel|finish_kernel|offsets_kernel|clip_scan_kernel|pack_kernel' -s 8 -c 8 -f -o gpurun_out/${TAG}_prof $CMD2 > gpurun_out/${TAG}_ncu_full.log 2>&1
    echo "full capture rc=$?"
    # block switching: launch list of one switched encode (transient detector + per-geometry launches)
    CMD3="python bench.py --steps 1 --warmup 1 --seconds 300 --block-switching --no-cpu-baseline"
    $CMD3 > gpurun_out/${TAG}_ncu_plain3.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_switching.csv $CMD3 > gpurun_out/${TAG}_ncu_launches3.log 2>&1
    echo "switching launch list rc=$?"
    ls -la gpurun_out/
fi
exit $RC
