#!/bin/bash
P=gpurun_out; mkdir -p $P
export SS_CLUSTER_COOPERATIVE=0
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/stage_launches.csv python tools/ncu_targets.py > $P/ncu_stage.log 2>&1; echo "ncu stage rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'encoder_layers_cluster' -o $P/cluster_full -f python tools/ncu_targets.py > $P/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -5 $P/ncu_stage.log; tail -5 $P/ncu_full.log; grep -c gpu__time $P/stage_launches.csv
