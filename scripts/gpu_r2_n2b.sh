#!/bin/bash
# 2-GPU sanity of the final build: config 2 (weak) and config 3 (strong split) under torchrun, reference arm
mkdir -p gpurun_out
run2() {  # name, args...
  local name=$1; shift
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 "$@" > gpurun_out/n2_$name.out 2> gpurun_out/n2_$name.err
  echo "== $name exit $? stdout lines: $(wc -l < gpurun_out/n2_$name.out)"; head -c 400 gpurun_out/n2_$name.out; echo
}
run2 config2 --steps 2 --warmup 3 --no-cpu-baseline
run2 config3 --workload config3 --steps 1 --warmup 3 --no-profile --no-cpu-baseline
run2 ref --impl reference --steps 1 --warmup 0
