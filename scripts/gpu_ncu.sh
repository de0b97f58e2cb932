#!/bin/bash
# ncu evidence: launch list (time per launch) + full-set captures of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit $?"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 10 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attention capture exit $?"
$CMD > gpurun_out/ncu_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:snake_aa_kernel -s 30 -c 2 -o gpurun_out/prof_snake $CMD > gpurun_out/ncu_snake.log 2>&1
echo "snake capture exit $?"
ls -la gpurun_out | tail -20
