P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 600 python -m pytest tests -m gpu -x -q > $P/t.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
timeout 120 python tools/umma2_check.py stamps > $P/u2_stamps.log 2>&1; echo "stamps rc=$?" >> $P/rc.log
run() { name=$1; shift; env "$@" timeout 120 python tools/stage_profile.py --out $P/stage_$name.json > $P/stage_$name.log 2>&1; echo "== $name $@"; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|_total|_host)" $P/stage_$name.log | cut -c1-110; }
run default X=1
run graph SS_VOCODER_GRAPH=1
timeout 300 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
cat $P/rc.log; tail -4 $P/t.log; tail -7 $P/u2_stamps.log | cut -c1-420; cut -c1-1000 $P/bench.json
