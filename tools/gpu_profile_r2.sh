#!/bin/bash
# Round-2 profiling pass (run under gpurun): launch lists + ncu --set full captures of the hot kernels.
# usage: bash tools/gpu_profile_r2.sh <tag>
TAG=${1:-r2_08}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
LOG=$O/${TAG}_prof.log
: > $LOG
# every launch of one step (c2), of one factorisation (h512) and of one sweep + top-k over 2^20 points, with device times
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_step_c2.csv python tools/prof_stage.py step c2 >> $LOG 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_factor_h512.csv python tools/prof_stage.py factor h512 >> $LOG 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches_acq_1m.csv python tools/prof_stage.py acq c2 >> $LOG 2>&1
# variance GEMM (the dominant kernel) at both sizes
$NCU --set full --import-source on -k regex:gemm_tc_kernel -c 1 -o $O/${TAG}_ncu_pgemm_c2 -f python tools/prof_stage.py predict c2 >> $LOG 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_kernel -c 1 -o $O/${TAG}_ncu_pgemm_h512 -f python tools/prof_stage.py predict h512 >> $LOG 2>&1
# the Cholesky panel kernel (2nd panel of the h512 factorisation) and its outer trailing update
$NCU --set full --import-source on -k regex:chol_panel_kernel -s 1 -c 1 -o $O/${TAG}_ncu_chol_panel_h512 -f python tools/prof_stage.py factor h512 >> $LOG 2>&1
$NCU --set full --import-source on -k regex:chol_panel_kernel -s 1 -c 1 -o $O/${TAG}_ncu_chol_panel_c2 -f python tools/prof_stage.py factor c2 >> $LOG 2>&1
# K* tile assembly
$NCU --set full --import-source on -k regex:kcross_mean -c 1 -o $O/${TAG}_ncu_kcross_h512 -f python tools/prof_stage.py predict h512 >> $LOG 2>&1
for f in $O/${TAG}_ncu_*.ncu-rep; do
    python tools/ncu_summary.py $f > ${f%.ncu-rep}.md 2>> $LOG
    case $f in *chol_panel_c2*) ;; *) rm -f $f ;; esac
done
for f in $O/${TAG}_launches_*.csv; do python tools/summarize_launches.py $f > ${f%.csv}.md 2>> $LOG; done
grep -E "^ok|Error|error" $LOG | tail -12
ls $O | grep ${TAG}_
