#!/bin/bash
# Template of one gpurun call (run as: gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'); edit per experiment.
# Test files run in separate processes: a device fault in one file must not poison the others.
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log $P/parity_report.jsonl $P/t_*.log
for f in test_gpu_parity test_multistream test_offline_batch; do
  timeout 900 python -m pytest tests/$f.py -m "gpu" -q --durations=6 > $P/t_$f.log 2>&1; echo "pytest $f rc=$?" >> $P/rc.log
done
timeout 120 python tools/stage_profile.py --out $P/stage.json > $P/stage.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
cat $P/rc.log; grep -hE 'FAILED|ERROR|passed|failed' $P/t_*.log | tail -40; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage.log | cut -c1-110; cut -c1-1800 $P/bench.json; tail -3 $P/bench.err
