#!/bin/bash
# tensor-core Snake: kernel tests, BigVGAN parity, micro-benchmark (default = split taps, variant = single tap)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "snake" > gpurun_out/sn_kernels.log 2>&1
echo "kernels exit $?"; tail -n 15 gpurun_out/sn_kernels.log
timeout -k 10 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider -s -k "bigvgan" > gpurun_out/sn_e2e.log 2>&1
echo "e2e exit $?"; tail -n 5 gpurun_out/sn_e2e.log; grep -h "rel-L2\|rel_l2" gpurun_out/sn_e2e.log | head -40
timeout -k 10 300 python scripts/kbench.py snake > gpurun_out/sn_kbench.txt 2>&1; cat gpurun_out/sn_kbench.txt
for lib in seed-vc_b200/libseedvc_b200_nosplit.so; do
  [ -e "$lib" ] || continue
  echo "== $lib"; SEEDVC_B200_LIB=$PWD/$lib timeout -k 10 300 python scripts/kbench.py snake 2>&1 | tee gpurun_out/sn_kbench_nosplit.txt
  SEEDVC_B200_LIB=$PWD/$lib timeout -k 10 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_full_size.py -q -m gpu -p no:cacheprovider -s -k "bigvgan" > gpurun_out/sn_e2e_nosplit.log 2>&1
  echo "nosplit e2e exit $?"; grep -h "rel-L2\|rel_l2" gpurun_out/sn_e2e_nosplit.log | head -40
done
timeout -k 10 600 python scripts/voc_profile.py > gpurun_out/sn_voc_profile.txt 2>&1; tail -n 30 gpurun_out/sn_voc_profile.txt
