P=gpurun_out; mkdir -p $P
timeout 240 python tools/umma2_check.py conv > $P/u2_conv.log 2>&1; echo "conv rc=$?" > $P/rc.log
timeout 240 python tools/umma2_check.py time > $P/u2_time.log 2>&1; echo "time rc=$?" >> $P/rc.log
timeout 240 python tools/umma2_check.py e2e > $P/u2_e2e.log 2>&1; echo "e2e rc=$?" >> $P/rc.log
if grep -q "conv rc=0" $P/rc.log; then
SS_UMMA_VOCODER=12 SS_UMMA_LINEAR=13 timeout 500 python -m pytest tests -m gpu -x -q > $P/t_umma2.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
SS_UMMA_VOCODER=12 SS_UMMA_LINEAR=13 timeout 200 python tools/stage_profile.py --out $P/stage_umma2.json > $P/stage_umma2.log 2>&1
SS_UMMA_VOCODER=12 SS_UMMA_LINEAR=13 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $P/bench_umma2.json 2> $P/bench_umma2.err
fi
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $P/launches_base.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --ncu-window > $P/ncu_base.log 2>&1; echo "ncu rc=$?" >> $P/rc.log
cat $P/rc.log; tail -30 $P/u2_conv.log; tail -14 $P/u2_time.log; tail -8 $P/u2_e2e.log; tail -5 $P/t_umma2.log; tail -12 $P/stage_umma2.log; cat $P/bench_umma2.json | cut -c1-700; wc -l $P/launches_base.csv
