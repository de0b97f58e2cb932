#!/bin/bash
# quick iteration: tc kernel tests, bf16 e2e, micro-benchmarks
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "bf16-False or share_weight" > gpurun_out/it_kernels.log 2>&1
echo "kernels exit $?"; tail -n 3 gpurun_out/it_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -s -k "not fp32" > gpurun_out/it_e2e.log 2>&1
echo "e2e exit $?"; tail -n 3 gpurun_out/it_e2e.log; grep -h "rel-L2" gpurun_out/it_e2e.log | sed 's/^\.*//' | sort | awk '{print}' | head -40
timeout -k 10 600 python scripts/kbench.py $KB > gpurun_out/kbench.txt 2>&1; cat gpurun_out/kbench.txt
