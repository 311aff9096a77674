#!/bin/bash
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log $P/t_*.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "480 or front_door" > $P/t_fix.log 2>&1; echo "pytest fix rc=$?" >> $P/rc.log
timeout 200 python tools/persist_phases.py > $P/phases_fused.log 2>&1
SS_PERSISTENT_FFN_FUSED=0 timeout 200 python tools/persist_phases.py > $P/phases_unfused.log 2>&1
timeout 200 python tools/mt_profile.py > $P/mt_profile.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras > $P/bench_noextras.json 2> $P/bench_noextras.err; echo "bench rc=$?" >> $P/rc.log
timeout 500 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench extras rc=$?" >> $P/rc.log
cat $P/rc.log; tail -3 $P/t_fix.log; head -30 $P/phases_fused.log; tail -12 $P/phases_fused.log; head -25 $P/phases_unfused.log | tail -20; cat $P/mt_profile.log; cut -c1-700 $P/bench_noextras.json; tail -4 $P/bench.err
