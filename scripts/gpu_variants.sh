#!/bin/bash
# run one kbench case against every experiment build: KB=<kbench arg> bash scripts/gpu_variants.sh
KB=${KB:-attn}
echo "== default"; python scripts/kbench.py $KB 2>&1 | head -${LINES_PER:-2}
for lib in seed-vc_b200/libseedvc_b200_v*.so; do
  [ -e "$lib" ] || continue
  echo "== $lib"; SEEDVC_B200_LIB=$PWD/$lib python scripts/kbench.py $KB 2>&1 | head -${LINES_PER:-2}
done
