#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
for wl in config2 config1; do
timeout -k 10 900 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_${wl}_d.json 2> gpurun_out/bench_${wl}_d.err; echo "bench $wl exit $?"; tail -n 3 gpurun_out/bench_${wl}_d.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${wl}_d.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])
PY
done
timeout -k 10 600 python bench.py --workload config1 --no-graph --steps 5 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_config1_nograph.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_config1_nograph.json')); print('config1 eager (C step):', d['value'], d['ms_per_step'])"
