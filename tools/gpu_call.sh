P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 120 python tools/umma2_check.py stamps > $P/u2_stamps.log 2>&1; echo "stamps rc=$?" >> $P/rc.log
timeout 120 python tools/mt_profile.py > $P/mt_profile.log 2>&1; echo "mt rc=$?" >> $P/rc.log
timeout 330 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/stage_launches_v7.csv python tools/ncu_targets.py > $P/ncu_targets.log 2>&1; echo "ncu rc=$?" >> $P/rc.log
cat $P/rc.log; tail -8 $P/u2_stamps.log; tail -50 $P/mt_profile.log; wc -l $P/stage_launches_v7.csv
