#!/bin/bash
# GPU call r2c: ncu of the lane-split kernels
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
NCU_SOURCE=1 timeout 600 bash tools/ncu_capture.sh split_threebody18_dop853_sparse threebody18_dp8 262144 $O/r2c_ncu.json > $O/r2c_cap1.log 2>&1
timeout 600 bash tools/ncu_capture.sh split_threebody18_dop853_cont8 threebody18_dp8_com 262144 $O/r2c_ncu.json > $O/r2c_cap2.log 2>&1
cat $O/r2c_cap1.log $O/r2c_cap2.log | tail -n 4; ls -la $O
