#!/bin/bash
# compute-sanitizer pass over the kernel parity tests: ONE tool per gpurun call (B200_PROFILING.md).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh memcheck'     (or racecheck / synccheck / initcheck)
# The hand-rolled mbarrier rings, TMEM double buffering and cp.reduce.async.bulk in-place adds of gemm.cu /
# attention.cu are what the racecheck / memcheck runs are for; summaries go to profiles/rN_sanitizer_<tool>.txt.
TOOL=${1:-memcheck}
SEL=${2:-"gemm_epilogues or gemm_store_paths or gemm_plain or conv or attention or norm_mod or snake or cfg_euler or reflect or hift or sola or crossfade"}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -n 1 gpurun_out/sanitize_plain.log
timeout -k 10 ${SAN_TIMEOUT:-1300} compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 99 \
    python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard|Invalid|error" gpurun_out/sanitize_$TOOL.log | head -20
