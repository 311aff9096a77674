#!/bin/bash
# Template of one gpurun call (run as: gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'); edit per experiment.
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log $P/parity_report.jsonl
timeout 900 python -m pytest tests -m "gpu" -q -x --durations=8 > $P/t.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
timeout 120 python tools/stage_profile.py --out $P/stage.json > $P/stage.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
cat $P/rc.log; tail -25 $P/t.log; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage.log | cut -c1-110; cut -c1-1500 $P/bench.json; tail -3 $P/bench.err
