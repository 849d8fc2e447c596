#!/bin/bash
# round-1 queued kernel variants, finally measured: parity (test_gpu_tc) then bench A/B (default first and last)
for n in pdl1 t2s5 spf bal; do
  echo "== parity $n"; FCN_LIB_PATH=$PWD/frustum_convnet_b200/variants/libfrustum_b200_$n.so \
    timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_mega.py -x -q -m gpu 2>&1 | tail -1
done
bash scripts/gpu_ab.sh --steps 200 -- pdl1 t2s5 spf bal
for e in "FCN_PRIO_PN=-1" "FCN_PRIO_CONV=-1"; do
  echo "== $e"; env $e bash scripts/gpu_ab.sh --steps 200 -- | head -1
done
