#!/bin/bash
# Staged GPU validation: each stage runs in its own process (a trapped kernel poisons the CUDA
# context) under `timeout`.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
    local name=$1 tmo=$2; shift 2
    timeout -k 10 "$tmo" python -m pytest "$@" -q -m gpu -p no:cacheprovider -s > "gpurun_out/$name.log" 2>&1
    echo "== $name: exit $? ==" | tee -a gpurun_out/summary.txt
    tail -n 4 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run k_misc 600 tests/test_gpu_kernels.py -k "not gemm and not attention"
run k_gemm_fp32 600 tests/test_gpu_kernels.py -k "gemm and not bf16-False and not share_weight"
run k_gemm_tc 600 tests/test_gpu_kernels.py -k "gemm and bf16-False or share_weight"
run k_attn_simt 600 tests/test_gpu_kernels.py -k "attention and not bf16-False"
run k_attn_tc 600 tests/test_gpu_kernels.py -k "attention and bf16-False"
run e2e_fp32 900 tests/test_gpu_e2e.py -k "fp32"
run e2e_bf16 900 tests/test_gpu_e2e.py -k "not fp32"
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "== smoke: exit $? ==" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
