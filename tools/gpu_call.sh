P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 600 python -m pytest tests -m gpu -x -q > $P/t.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $P/smoke.log 2>&1; echo "smoke rc=$?" >> $P/rc.log
timeout 300 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $P/bench_ref.json 2> $P/bench_ref.err; echo "ref rc=$?" >> $P/rc.log
timeout 120 python tools/stage_profile.py --out $P/stage_final.json > $P/stage_final.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"encoder_layers_persistent|umma2_kernel" -c 16 -f -o $P/r1_full_v9 python tools/ncu_targets.py > $P/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $P/rc.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/stage_launches_v9.csv python tools/ncu_targets.py > $P/ncu_targets.log 2>&1; echo "ncu list rc=$?" >> $P/rc.log
cat $P/rc.log; tail -3 $P/t.log; tail -2 $P/smoke.log; cut -c1-1400 $P/bench.json; echo; cut -c1-500 $P/bench_ref.json; echo; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage_final.log | cut -c1-110
