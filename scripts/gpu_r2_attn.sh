#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "attention or hift" > gpurun_out/k_attn.log 2>&1
echo "attention+hift kernel tests exit $?"; tail -n 6 gpurun_out/k_attn.log
timeout -k 10 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -s -k "hift" > gpurun_out/e2e_hift.log 2>&1
echo "hift e2e exit $?"; grep -E "rel-L2|passed|failed" gpurun_out/e2e_hift.log | tail -12
( KB=attn LINES_PER=4 bash scripts/gpu_variants.sh ) > gpurun_out/attn_variants.txt 2>&1; cat gpurun_out/attn_variants.txt
timeout -k 10 900 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/bench_config2_b.json 2> gpurun_out/bench_config2_b.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_config2_b.json'))
print(d['value'], d['ms_per_step'], d['clocks'])
for k,v in d['kernel_breakdown'].items():
    if v['ms']>0.5: print('  ',k,v)
PY
timeout -k 10 900 python bench.py --workload config5_hift --euler-steps 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_config5_hift_n4.json 2> gpurun_out/bench_config5_hift_n4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_config5_hift_n4.json'))
print(d['value'], d['ms_per_step'], d['latency'] and {k:v for k,v in d['latency'].items() if k!='what'})
for k,v in d['kernel_breakdown'].items():
    if v['ms']>0.5: print('  ',k,v)
PY
