#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "hift or sola" -x > gpurun_out/k_hift.log 2>&1
echo "hift kernels exit $?"; tail -n 15 gpurun_out/k_hift.log
timeout -k 10 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -s -k "hift or graphed" > gpurun_out/e2e_hift.log 2>&1
echo "hift e2e exit $?"; grep -E "rel-L2|passed|failed|Error|error" gpurun_out/e2e_hift.log | tail -30
timeout -k 10 900 python bench.py --workload config5_hift --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_config5_hift.json 2> gpurun_out/bench_config5_hift.err
echo "bench config5_hift exit $?"; head -c 600 gpurun_out/bench_config5_hift.json; echo; tail -n 5 gpurun_out/bench_config5_hift.err
timeout -k 10 900 python bench.py --workload config5_hift --euler-steps 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_config5_hift_n4.json 2> gpurun_out/bench_config5_hift_n4.err
echo "bench config5_hift n4 exit $?"; head -c 600 gpurun_out/bench_config5_hift_n4.json; echo; tail -n 5 gpurun_out/bench_config5_hift_n4.err
