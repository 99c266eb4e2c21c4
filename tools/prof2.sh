O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/r1c_launches_factor_h512.csv python tools/prof_stage.py factor h512 > $O/r1c_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/r1c_launches_factor_c2.csv python tools/prof_stage.py factor c2 >> $O/r1c_prof.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/r1c_launches_chol_h512.csv python tools/prof_stage.py chol h512 >> $O/r1c_prof.log 2>&1
tail -2 $O/r1c_prof.log
