#!/bin/bash
# A/B candidates queued for the next GPU session (build here on the CPU box, then one gpurun call):
#   pdl1  : release PDL dependents after the last MMA issue instead of right after griddepcontrol.wait
#   t2s5  : 5 weight stages (instead of 3) in the 2-CTA PointNet kernel (smem freed by the slab removal)
#   spf   : hand-pipelined shared-memory loads in the PointNet SIMT phases (layer 1 weights, epilogue-2 biases)
#   bal   : persistent PointNet kernels use only the CTAs their round count needs (frees SMs for other streams)
#   gm2   : conv GEMM with 2 instead of 3 stages (129 KB: can share an SM with a 64-channel PointNet CTA)
#   both  : pdl1 + t2s5 + spf + bal together
# Usage:  scripts/next_round_ab.sh build      (CPU container)
#         gpurun --timeout 600 -- 'bash scripts/next_round_ab.sh run'
set -e
cd "$(dirname "$0")/.."
case "$1" in
  build)
    scripts/build_variant.sh pdl1 -DFCN_PDL_MODE=1
    scripts/build_variant.sh t2s5 -DFCN_T2_NSTAGE=5
    scripts/build_variant.sh spf -DFCN_SIMT_PREFETCH=1
    scripts/build_variant.sh bal -DFCN_BALANCE_ROUNDS=1
    scripts/build_variant.sh gm2 -DFCN_GM_NSTAGE=2
    scripts/build_variant.sh both -DFCN_PDL_MODE=1 -DFCN_T2_NSTAGE=5 -DFCN_SIMT_PREFETCH=1 -DFCN_BALANCE_ROUNDS=1 ;;
  run)
    for n in pdl1 t2s5 spf bal gm2 both; do
      FCN_LIB_PATH=$PWD/frustum_convnet_b200/variants/libfrustum_b200_$n.so \
        timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -1
    done
    bash scripts/gpu_ab.sh -- pdl1 t2s5 spf bal gm2 both
    # runtime knobs (default library): launch priority of the PointNet / conv kernels
    for e in "FCN_PRIO_PN=-1" "FCN_PRIO_CONV=-1" "FCN_PRIO_PN=-2 FCN_PRIO_CONV=-1"; do
      echo "== $e"; env $e bash scripts/gpu_ab.sh -- | head -1
    done ;;
  *) echo "usage: $0 build|run"; exit 1 ;;
esac
