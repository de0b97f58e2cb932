#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python bench.py --workload smoke --steps 2 --warmup 1 > gpurun_out/bench_smoke.json 2> gpurun_out/bench_smoke.err
echo "smoke exit $?"; tail -c 2500 gpurun_out/bench_smoke.json; tail -n 5 gpurun_out/bench_smoke.err
timeout -k 10 1200 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err
echo "config2 exit $?"; tail -c 4000 gpurun_out/bench_config2.json; tail -n 8 gpurun_out/bench_config2.err
