#!/bin/bash
# ncu evidence: launch list (time per launch) + full-set captures of the top kernels.
# gpurun copies back at most 64 MiB: every .ncu-rep is exported to its raw-page CSV (what scripts/summarize_ncu.py
# reads) + a per-instruction source page, and then deleted on the box.
mkdir -p gpurun_out
export_rep() {   # name
  local f=gpurun_out/$1.ncu-rep
  [ -e "$f" ] || return
  ncu -i $f --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i $f --page source --csv 2>/dev/null | head -c 6000000 > gpurun_out/$1_source.csv
  rm -f $f
}
CMD="python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
# launch list of ONE conversion pass of the bench's default command (config 2): our kernels only
C2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
$C2 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:svc:: -c ${PASS_LAUNCHES:-3520} --csv --log-file gpurun_out/launches.csv $C2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:svc:: -c ${PASS_LAUNCHES:-3520} --csv $C2" > gpurun_out/launches.cmd
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit $?"; export_rep prof_gemm
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 10 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attention capture exit $?"; export_rep prof_attn
$CMD > gpurun_out/ncu_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:snake_mma -s 30 -c 1 -o gpurun_out/prof_snake $CMD > gpurun_out/ncu_snake.log 2>&1
echo "snake capture exit $?"; export_rep prof_snake
# DRAM traffic of every gemm_tc_kernel launch of ONE config-2 conversion pass (bench default workload):
# feeds roofline.traffic (profiles/<round>_gemm_traffic.json)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel -c ${GEMM_LAUNCHES:-2142} --csv --log-file gpurun_out/gemm_traffic.csv $C2 > gpurun_out/ncu_traffic.log 2>&1
echo "gemm traffic exit $?"
du -sh gpurun_out; ls -la gpurun_out | tail -20
