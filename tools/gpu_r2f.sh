#!/bin/bash
# GPU call r2f: tuning variants of the lane-split kernels + f32 tests
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for v in "" variants/libode_sp256.so variants/libode_sp512.so variants/libode_dn384.so variants/libode_sync2.so variants/libode_sync8.so; do
  echo "== lib=$v" >> $O/r2f_kb.log
  ODE_B200_LIB=$v timeout 300 python tools/kbench.py threebody18_dp8 threebody18_dp8_com 2>&1 | grep -v "output bytes" >> $O/r2f_kb.log
done
cat $O/r2f_kb.log
timeout 900 python -m pytest tests/test_gpu_f32.py -q > $O/r2f_f32.log 2>&1; tail -n 3 $O/r2f_f32.log
