#!/bin/bash
mkdir -p gpurun_out
for w in qkvN w13; do python scripts/gemm_trace.py $w > gpurun_out/gemm_trace_$w.txt 2>&1; echo "== $w"; sed -n 16,44p gpurun_out/gemm_trace_$w.txt; done
