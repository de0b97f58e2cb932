#!/bin/bash
# 2-GPU run: torchrun plumbing (stdout discipline, strong split) + BigVGAN C path test on one GPU first
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider -k "bigvgan" > gpurun_out/voc_c.log 2>&1
echo "bigvgan tests exit $?"; tail -n 4 gpurun_out/voc_c.log
run2() {  # name, args...
  local name=$1; shift
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 "$@" > gpurun_out/n2_$name.out 2> gpurun_out/n2_$name.err
  echo "== $name exit $? stdout lines: $(wc -l < gpurun_out/n2_$name.out)"; head -c 500 gpurun_out/n2_$name.out; echo
}
run2 config2 --steps 2 --warmup 1
run2 config3 --workload config3 --steps 2 --warmup 1 --no-profile
run2 config4 --workload config4 --steps 2 --warmup 1 --no-profile
run2 ref --impl reference --steps 1 --warmup 0
NCCL_DEBUG=INFO run2 config2_nccl --steps 1 --warmup 1 --no-profile
grep -c "NCCL INFO" gpurun_out/n2_config2_nccl.err; grep -m2 "nranks" gpurun_out/n2_config2_nccl.err | cut -c1-200
