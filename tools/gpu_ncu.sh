#!/bin/bash
# ncu captures of every kernel class on the bench-shaped problem (one GPU; run after the plain command exits 0)
set -u
mkdir -p gpurun_out
CMD="python tools/ncu_targets.py"
$CMD > gpurun_out/ncu_targets_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'embed_fwd_kernel|embed_bwd_kernel|gram_kernel|gram_grad_rows|gemvT_partial|gemv_lower|kst_kernel|score_kernel' \
    -c 14 -o gpurun_out/prof_small -f $CMD > gpurun_out/ncu_small.log 2>&1
echo "ncu small rc=$?"
$CMD > gpurun_out/ncu_targets_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_tiles_kernel|panel_kernel' -s 200 -c 6 \
    -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
tail -3 gpurun_out/ncu_small.log gpurun_out/ncu_gemm.log
