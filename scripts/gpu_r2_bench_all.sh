#!/bin/bash
# parity suites + every bench workload on one GPU
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -x > gpurun_out/k_e2e.log 2>&1
echo "kernels+e2e exit $?"; tail -n 3 gpurun_out/k_e2e.log
timeout -k 10 900 python -m pytest tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider -s > gpurun_out/full_size.log 2>&1
echo "full-size exit $?"; grep -E "rel-L2|passed|failed" gpurun_out/full_size.log | grep -E "bf16|passed|failed" | tail -20
for wl in config2 config1 config5 config3 config4; do
  timeout -k 10 900 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl exit $?"; head -c 700 gpurun_out/bench_$wl.json; echo; tail -n 3 gpurun_out/bench_$wl.err
done
timeout -k 10 600 python bench.py --workload config5 --euler-steps 4 --steps 3 --warmup 2 --no-cpu-baseline --no-profile > gpurun_out/bench_config5_n4.json 2> gpurun_out/bench_config5_n4.err
echo "bench config5 n4 exit $?"; head -c 500 gpurun_out/bench_config5_n4.json; echo
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; head -c 400 gpurun_out/bench_ref.json; echo
