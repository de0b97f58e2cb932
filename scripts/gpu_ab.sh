#!/bin/bash
# A/B of two library builds in alternating order (box-to-box and process-to-process noise is +-4 %):
#   build the variant as seed-vc_b200/libseedvc_b200_v_old.so (python seed-vc_b200/_build.py --variant v_old --sources <file>.cu <DEFINE>)
for i in 1 2; do
for lib in seed-vc_b200/libseedvc_b200_v_old.so seed-vc_b200/libseedvc_b200.so; do
 echo "== $lib"
 [ -n "$AB_VOC" ] && SEEDVC_B200_LIB=$PWD/$lib timeout 200 python scripts/voc_profile.py 2>&1 | tail -1
 [ -n "$AB_KB" ] && SEEDVC_B200_LIB=$PWD/$lib KB_FILTER="$AB_KB" timeout 200 python scripts/kbench.py gemm 2>&1 | head -4
 SEEDVC_B200_LIB=$PWD/$lib timeout 300 python scripts/dit_profile.py 2>&1 | grep -E "total per step|norm_mod|'rope,out_op'"
done; done
