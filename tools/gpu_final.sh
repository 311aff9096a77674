#!/bin/bash
# Final validation of a round in ONE GPU call: the driver's own commands (single-process pytest, smoke, bench), then the ncu evidence
# (launch list of one invocation of every stage, --set full of the dominant kernel, launch list of the bench window).
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log $P/parity_report.jsonl
timeout 900 python -m pytest tests/ -x -q -m gpu > $P/t_all.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $P/smoke.log 2>&1; echo "smoke rc=$?" >> $P/rc.log
timeout 600 python bench.py --steps 10 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $P/bench_ref.json 2> $P/bench_ref.err; echo "bench ref rc=$?" >> $P/rc.log
SS_CLUSTER_COOPERATIVE=0 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/stage_launches.csv python tools/ncu_targets.py > $P/ncu_stage.log 2>&1; echo "ncu stage rc=$?" >> $P/rc.log
SS_CLUSTER_COOPERATIVE=0 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'encoder_layers_cluster' -o $P/cluster_full -f python tools/ncu_targets.py > $P/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $P/rc.log
timeout 200 python tools/stage_profile.py --out $P/stage.json > $P/stage.log 2>&1
timeout 200 python tools/cluster_ab.py > $P/cluster_ab.log 2>&1
timeout 200 python tools/mt_phases.py > $P/mt_phases.log 2>&1
cat $P/rc.log; tail -3 $P/t_all.log; tail -2 $P/smoke.log; cut -c1-600 $P/bench.json; cut -c1-300 $P/bench_ref.json; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage.log | cut -c1-110
