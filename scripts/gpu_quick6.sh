#!/bin/bash
# quick loop: parity tests, bench lines fp64/fp32, then the launch list for per-kernel durations
TAG=${1:-q}
bash scripts/gpu_quick4.sh $TAG
bash scripts/gpu_launchlist.sh $TAG
