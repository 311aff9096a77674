#!/bin/bash
# Two-GPU validation of both bench arms exactly as the driver launches them (run as: gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_call_n2.sh')
P=gpurun_out; mkdir -p $P; rm -f $P/rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $P/bench_n2.json 2> $P/bench_n2.err; echo "bench n2 rc=$?" >> $P/rc.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $P/ref_n2.json 2> $P/ref_n2.err; echo "ref n2 rc=$?" >> $P/rc.log
cat $P/rc.log; cut -c1-400 $P/bench_n2.json; echo; cut -c1-600 $P/ref_n2.json
