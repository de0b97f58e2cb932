#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "gemm or conv or guard" > gpurun_out/k_gemm.log 2>&1
echo "gemm tests exit $?"; tail -n 3 gpurun_out/k_gemm.log
timeout -k 10 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider > gpurun_out/e2e_all.log 2>&1
echo "e2e + full-size exit $?"; tail -n 3 gpurun_out/e2e_all.log
KB_FILTER="" python scripts/kbench.py gemm 2>&1 | head -12
for wl in config2 config5_hift; do
  timeout -k 10 900 python bench.py --workload $wl --no-cpu-baseline --steps 3 --warmup 2 --latency-ticks 50 > gpurun_out/bench_${wl}_f.json 2> gpurun_out/bench_${wl}_f.err; echo "bench $wl exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${wl}_f.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d.get('roofline',{}).get('achieved'))
for k,v in d['kernel_breakdown'].items():
    if v['ms']>10: print('  ',k,v)
PY
done
