#!/bin/bash
# round 2, first GPU call: full-size parity (new), whole GPU suite, 1-GPU bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider -s > gpurun_out/full_size.log 2>&1
echo "full-size parity exit $?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/full_size.log | tail -40
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_full_size.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -n 5 gpurun_out/pytest_gpu.log
timeout -k 10 900 python bench.py --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"; head -c 3000 gpurun_out/bench_ours.json; echo
