#!/bin/bash
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_api.py -q > $O/r2k_api.log 2>&1; tail -n 12 $O/r2k_api.log
timeout 600 python tools/kbench_com.py > $O/r2k_com.log 2>&1; cat $O/r2k_com.log
