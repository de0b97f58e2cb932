mkdir -p gpurun_out
CMD="python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
C2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
$C2 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:svc:: -c 3519 --csv --log-file gpurun_out/launches.csv $C2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:svc:: -c 3519 --csv $C2" > gpurun_out/launches.cmd
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 6 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit $?"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 10 -c 2 -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attention capture exit $?"
