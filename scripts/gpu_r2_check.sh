#!/bin/bash
# full GPU suite + smoke + default bench + per-shape DiT / vocoder profiles
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
timeout -k 10 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_config2.json'))
print(d['value'], d['e2e'], d['ms_per_step'], d['clocks'], d['cpu_baseline'], d['roofline'])
for k,v in d['kernel_breakdown'].items():
    if v['ms']>0.5: print('  ',k,v)
PY
timeout -k 10 600 python scripts/dit_profile.py > gpurun_out/dit_profile.txt 2>&1; tail -n 40 gpurun_out/dit_profile.txt
