#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
MRC_TIMELINE=1 timeout 600 python bench.py --steps 1 --warmup 0 --no-sequential-sample --no-cpu-baseline --no-decode --no-music 2> gpurun_out/${TAG}_timeline.txt > /dev/null
grep "serial pass" gpurun_out/${TAG}_timeline.txt | tail -2
