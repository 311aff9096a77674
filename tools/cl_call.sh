#!/bin/bash
P=gpurun_out; mkdir -p $P
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cluster_encoder" > $P/t_cluster.log 2>&1; echo "pytest cluster rc=$?"
tail -3 $P/t_cluster.log
timeout 200 python tools/cluster_ab.py > $P/cluster_ab.log 2>&1; echo "ab rc=$?"; tail -3 $P/cluster_ab.log | cut -c1-1500
