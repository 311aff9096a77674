P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 300 python -m pytest tests -m gpu -q > $P/t_default.log 2>&1; echo "pytest default rc=$?" >> $P/rc.log
SS_UNIT_GROUPED=1 timeout 300 python -m pytest tests -m gpu -q > $P/t_grouped.log 2>&1; echo "pytest grouped rc=$?" >> $P/rc.log
run() { name=$1; shift; env "$@" timeout 100 python tools/stage_profile.py --out $P/stage_$name.json > $P/stage_$name.log 2>&1; echo "== $name $@"; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage_$name.log | cut -c1-100; }
run f_default X=1
run f_grouped SS_UNIT_GROUPED=1
run f_grouped_split60 SS_UNIT_GROUPED=1 SS_UMMA2_SPLIT_BELOW=60
run f_grouped_split60_minch32 SS_UNIT_GROUPED=1 SS_UMMA2_SPLIT_BELOW=60 SS_UMMA_MIN_CHANNELS=32
cat $P/rc.log; tail -6 $P/t_default.log; tail -6 $P/t_grouped.log
