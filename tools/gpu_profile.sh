#!/bin/bash
# Round profiling pass (run under gpurun): launch lists + ncu --set full captures of the hot kernels.
# usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_step_c2.csv python tools/prof_stage.py step c2 > $O/${TAG}_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_factor_h512.csv python tools/prof_stage.py factor h512 >> $O/${TAG}_prof.log 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -c 1 -o $O/${TAG}_pgemm_h512 -f python tools/prof_stage.py predict h512 >> $O/${TAG}_prof.log 2>&1
$NCU --set full --import-source on -k regex:kcross_mean -c 1 -o $O/${TAG}_kcross_h512 -f python tools/prof_stage.py predict h512 >> $O/${TAG}_prof.log 2>&1
$NCU --set full --import-source on -k "regex:gemm_tc_kernel|diag_block_kernel|gemm_simt_kernel" -s 60 -c 3 -o $O/${TAG}_chol_h512 -f python tools/prof_stage.py chol h512 >> $O/${TAG}_prof.log 2>&1
$NCU --set full --import-source on -k regex:kmat_kernel -c 2 -o $O/${TAG}_kmat_h512 -f python tools/prof_stage.py kmat h512 >> $O/${TAG}_prof.log 2>&1
tail -3 $O/${TAG}_prof.log
ls -la $O
