#!/bin/bash
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log $P/t_*.log $P/parity_report.jsonl
for f in test_gpu_parity test_multistream test_offline_batch; do
  timeout 900 python -m pytest tests/$f.py -m "gpu" -q --durations=4 > $P/t_$f.log 2>&1; echo "pytest $f rc=$?" >> $P/rc.log
done
timeout 200 python tools/mt_profile.py > $P/mt_profile.log 2>&1
timeout 120 python tools/stage_profile.py --out $P/stage.json > $P/stage.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
cat $P/rc.log; grep -hE 'FAILED|ERROR|passed|failed' $P/t_*.log | tail -12; grep "prefix_kernel" $P/mt_profile.log; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage.log | cut -c1-110; cut -c1-500 $P/bench.json
