#!/bin/bash
# ncu --set full of the top kernels of the batch-64 inference forward (one gpurun call; the plain run goes first).
set -u
P=${1:-gpurun_out/r2_full}
python tools/ncu_forward.py 64 3 > ${P}_plain.log 2>&1 || exit 1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd_tc3 -s 20 -c 2 -o ${P}_attn -f python tools/ncu_forward.py 64 2 > ${P}_attn.log 2>&1
$NCU -k regex:pool_ln_march -s 50 -c 5 -o ${P}_pool -f python tools/ncu_forward.py 64 2 > ${P}_pool.log 2>&1
$NCU -k regex:gemm_tc_tma -s 80 -c 5 -o ${P}_gemm -f python tools/ncu_forward.py 64 2 > ${P}_gemm.log 2>&1
$NCU -k regex:patch_embed_tc -s 1 -c 1 -o ${P}_pe -f python tools/ncu_forward.py 64 2 > ${P}_pe.log 2>&1
$NCU -k regex:mlp_ -s 3 -c 3 -o ${P}_mlp -f python tools/ncu_forward.py 64 2 > ${P}_mlp.log 2>&1
ls -la ${P}_*.ncu-rep
