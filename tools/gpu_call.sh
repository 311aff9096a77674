P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 600 python -m pytest tests -m gpu -x -q > $P/t.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
timeout 300 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
timeout 120 python tools/stage_profile.py --out $P/stage_default.json > $P/stage_default.log 2>&1
SS_VOCODER_GRAPH=0 SS_PERSISTENT_PREFETCH=0 timeout 120 python tools/stage_profile.py --out $P/stage_nograph_noprefetch.json > $P/stage_nograph.log 2>&1
SS_UMMA2_SPLIT_BELOW=100 timeout 120 python tools/stage_profile.py --out $P/stage_split100.json > $P/stage_split100.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"encoder_layers_persistent|umma2_kernel" -c 24 -f -o $P/r1_full python tools/ncu_targets.py > $P/ncu_full.log 2>&1; echo "ncu rc=$?" >> $P/rc.log
cat $P/rc.log; tail -5 $P/t.log; cut -c1-1800 $P/bench.json; tail -3 $P/bench.err
for f in default nograph split100; do echo "== $f"; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|_total|_host)" $P/stage_$f.log | cut -c1-120; done
tail -4 $P/ncu_full.log; ls -la $P/*.ncu-rep
