#!/bin/bash
P=gpurun_out; mkdir -p $P
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mt or persistent or agent or fixture or s2tt" > $P/t_mt.log 2>&1; echo "pytest rc=$?"; tail -4 $P/t_mt.log
timeout 200 python tools/mt_profile.py > $P/mt_profile.log 2>&1; grep "prefix_kernel" $P/mt_profile.log
