#!/bin/bash
# round-end style validation: full GPU test suite, smoke, 1-GPU bench, reference arm
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout -k 10 1200 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"; head -c 1500 gpurun_out/bench_ours.json; echo
timeout -k 10 1200 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; head -c 600 gpurun_out/bench_ref.json; echo
