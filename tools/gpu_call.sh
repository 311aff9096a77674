#!/bin/bash
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 500 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --ncu-window > $P/ncu_bench.log 2>&1; echo "ncu bench window rc=$?" >> $P/rc.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/launches_stages.csv python tools/ncu_targets.py > $P/ncu_stages.log 2>&1; echo "ncu stages rc=$?" >> $P/rc.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"encoder_layers_persistent|mt_decode_persistent|mt_prefix_persistent" -c 4 -f -o $P/r2_persist_full python tools/ncu_targets.py > $P/ncu_full1.log 2>&1; echo "ncu full persist rc=$?" >> $P/rc.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/launches_large.csv python tools/ncu_large.py > $P/ncu_large.log 2>&1; echo "ncu large list rc=$?" >> $P/rc.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"umma2_kernel" -c 8 -f -o $P/r2_umma2_full python tools/ncu_large.py > $P/ncu_full2.log 2>&1; echo "ncu full umma2 rc=$?" >> $P/rc.log
ncu -i $P/r2_persist_full.ncu-rep --page raw --csv > $P/r2_persist_full_raw.csv 2>/dev/null
ncu -i $P/r2_umma2_full.ncu-rep --page raw --csv > $P/r2_umma2_full_raw.csv 2>/dev/null
rm -f $P/r2_umma2_full.ncu-rep
cat $P/rc.log; cut -c1-600 $P/bench.json; ls -la $P | head -40
