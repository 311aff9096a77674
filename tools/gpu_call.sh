#!/bin/bash
# Template of one gpurun call (run as: gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'); edit per experiment.
# Round-2 first call: the regression tests, the STAGED tests that have not run on a B200 yet, a stage profile and the bench.
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 600 python -m pytest tests -m "gpu or gpu_staged" -q > $P/t.log 2>&1; echo "pytest rc=$?" >> $P/rc.log
timeout 120 python tools/stage_profile.py --out $P/stage.json > $P/stage.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > $P/bench.json 2> $P/bench.err; echo "bench rc=$?" >> $P/rc.log
cat $P/rc.log; tail -6 $P/t.log; grep -E "^(mt_greedy|t2u|vocoder_generate|encoder_stream|ctc|_total|_host)" $P/stage.log | cut -c1-110; cut -c1-900 $P/bench.json
