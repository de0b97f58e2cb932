#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "gemm or conv" > gpurun_out/k_gemm.log 2>&1
echo "gemm tests exit $?"; tail -n 3 gpurun_out/k_gemm.log
SEEDVC_B200_LIB=$PWD/seed-vc_b200/libseedvc_b200_vdb.so timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "gemm or conv" > gpurun_out/k_gemm_vdb.log 2>&1
echo "gemm tests (vdb) exit $?"; tail -n 3 gpurun_out/k_gemm_vdb.log
( KB=gemm LINES_PER=30 bash scripts/gpu_variants.sh ) > gpurun_out/gemm_variants.txt 2>&1; cat gpurun_out/gemm_variants.txt
for lib in "" _vdb; do
  SEEDVC_B200_LIB=$PWD/seed-vc_b200/libseedvc_b200$lib.so timeout -k 10 900 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/bench_config2_e$lib.json 2> gpurun_out/bench_config2_e$lib.err; echo "bench$lib exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_config2_e$lib.json'))
print(d['value'], d['ms_per_step'], d['clocks'])
for k,v in d['kernel_breakdown'].items():
    if v['ms']>20: print('  ',k,v)
PY
done
