#!/bin/bash
# Round profiling pass (run under gpurun): launch lists + ncu --set full captures of the hot kernels.
# usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
LOG=$O/${TAG}_prof.log
: > $LOG
# every launch of one step (c2) and of one factorisation (h512) with its device time
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_step_c2.csv python tools/prof_stage.py step c2 >> $LOG 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_factor_h512.csv python tools/prof_stage.py factor h512 >> $LOG 2>&1
# variance GEMM (the dominant kernel) at both sizes
$NCU --set full --import-source on -k regex:gemm_tc_kernel -c 1 -o $O/${TAG}_pgemm_c2 -f python tools/prof_stage.py predict c2 >> $LOG 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -c 1 -o $O/${TAG}_pgemm_h512 -f python tools/prof_stage.py predict h512 >> $LOG 2>&1
# K* tile assembly
$NCU --set full --import-source on -k regex:kcross_mean -c 1 -o $O/${TAG}_kcross_h512 -f python tools/prof_stage.py predict h512 >> $LOG 2>&1
# Cholesky: first outer (K = 512) trailing update = 4th tcgen05 launch; an inner (K = 128) one; diag block + panel
# (standalone gpg_cholesky: SIMT panel, so per outer panel the tcgen05 launches are 3 inner updates + 1 outer)
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/${TAG}_chol_outer_h512 -f python tools/prof_stage.py chol h512 >> $LOG 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 4 -c 1 -o $O/${TAG}_chol_inner_h512 -f python tools/prof_stage.py chol h512 >> $LOG 2>&1
# the tensor-core panel of the factor path: 2nd tcgen05 launch of gpg_factorize (after the first panel's inner update)
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o $O/${TAG}_chol_panel_h512 -f python tools/prof_stage.py factor h512 >> $LOG 2>&1
$NCU --set full --import-source on -k "regex:diag_block_kernel" -s 8 -c 1 -o $O/${TAG}_chol_diag_h512 -f python tools/prof_stage.py chol h512 >> $LOG 2>&1
# marginal-likelihood gradient reduction (one Adam iteration at c2) and the acquisition sweep over 2^20 points
$NCU --set full --import-source on -k regex:grad_partial -c 1 -o $O/${TAG}_grad_c2 -f python tools/prof_stage.py fit c2 >> $LOG 2>&1
$NCU --set full --import-source on -k "regex:acq_eval_kernel|topk_round_kernel" -c 2 -o $O/${TAG}_acq_1m -f python tools/prof_stage.py acq c2 >> $LOG 2>&1
# inducing-point path (fp32, C2 size): launch list of two Adam iterations, the fused kernel-derivative reduction over the
# m x N sensitivity matrix (2nd launch of an iteration) and the split-K S = B B^T product (2nd tcgen05 launch)
$NCU --profile-from-start on --metrics gpu__time_duration.sum -c 300 --csv --log-file $O/${TAG}_launches_sparse_c2_f32.csv python tools/sparse_stages.py c2 2 f32 >> $LOG 2>&1
$NCU --profile-from-start on --set full --import-source on -k regex:sgp_kgrad_kernel -s 1 -c 1 -o $O/${TAG}_sparse_kgrad_c2 -f python tools/sparse_stages.py c2 2 f32 >> $LOG 2>&1
$NCU --profile-from-start on --set full --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -o $O/${TAG}_sparse_syrk_c2 -f python tools/sparse_stages.py c2 2 f32 >> $LOG 2>&1
# kernel-matrix assembly (full and lower-only)
$NCU --set full --import-source on -k regex:kmat_kernel -c 2 -o $O/${TAG}_kmat_h512 -f python tools/prof_stage.py kmat h512 >> $LOG 2>&1
# summarise on the box: the reports themselves are too big to travel back (64 MiB cap), keep only the GEMM one
for f in $O/${TAG}_*.ncu-rep; do
    python tools/ncu_summary.py $f > ${f%.ncu-rep}.md 2>> $LOG
    ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>> $LOG
    case $f in *pgemm_h512*) ;; *) rm -f $f ;; esac
done
grep -E "^ok|Error|error" $LOG | tail -12
ls $O | grep ${TAG}_
