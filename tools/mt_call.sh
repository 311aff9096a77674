#!/bin/bash
P=gpurun_out; mkdir -p $P
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mt or persistent_kernel or agent or fixture or s2tt" > $P/t_mt.log 2>&1; echo "pytest rc=$?"; tail -3 $P/t_mt.log
timeout 200 python tools/mt_profile.py > $P/mt_profile.log 2>&1; grep "prefix_kernel 1" $P/mt_profile.log
timeout 100 python tools/mt_phases.py 2>&1 | tail -1 | cut -c1-700
