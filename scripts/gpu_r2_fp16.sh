#!/bin/bash
# fp16 operand mode: kernel parity, e2e parity, full-size parity, bench in both 16-bit modes
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x > gpurun_out/k_all.log 2>&1
echo "kernels exit $?"; tail -n 4 gpurun_out/k_all.log
timeout -k 10 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -s > gpurun_out/e2e_all.log 2>&1
echo "e2e exit $?"; grep -E "rel-L2|passed|failed" gpurun_out/e2e_all.log | grep -E "fp16|passed|failed" | tail -40
timeout -k 10 900 python -m pytest tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider -s > gpurun_out/full_size.log 2>&1
echo "full-size exit $?"; grep -E "rel-L2|passed|failed" gpurun_out/full_size.log | tail -40
timeout -k 10 900 python bench.py --no-cpu-baseline --mode fp16 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench fp16 exit $?"; head -c 2500 gpurun_out/bench_fp16.json; echo
