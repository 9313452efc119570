#!/bin/bash
TAG=${1:-q}
bash scripts/gpu_quick4.sh $TAG
bash scripts/gpu_launchlist.sh $TAG
