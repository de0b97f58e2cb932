#!/bin/bash
# usage: gpu_quick.sh <name> <pytest args...>   (one stage of gpu_check.sh)
mkdir -p gpurun_out
name=$1; shift
timeout -k 10 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider -s > "gpurun_out/$name.log" 2>&1
echo "== $name: exit $? =="; tail -n 15 "gpurun_out/$name.log"
