#!/bin/bash
# ncu full capture of one kbench case: KB=<kbench arg> KB_FILTER=<substr> KREGEX=<kernel regex>
mkdir -p gpurun_out
KB=${KB:-gemm}
python scripts/kbench.py $KB > gpurun_out/ncuk_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -f -o gpurun_out/prof_k python scripts/kbench.py $KB > gpurun_out/ncuk.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncuk.log
