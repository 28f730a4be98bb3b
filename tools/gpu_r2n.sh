#!/bin/bash
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2n_kb.log
for v in "" variants/libode_rk4g4.so; do
  echo "== lib=$v" >> $O/r2n_kb.log
  ODE_B200_LIB=$v timeout 300 python tools/kbench.py lorenz_rk4_all_steps lorenz_rk4 >> $O/r2n_kb.log 2>&1
  KBENCH_N=131072 ODE_B200_LIB=$v timeout 300 python tools/kbench.py lorenz_rk4_all_steps >> $O/r2n_kb.log 2>&1
done
cat $O/r2n_kb.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_host_api.py tests/test_gpu_examples.py -q > $O/r2n_py.log 2>&1; tail -n 3 $O/r2n_py.log
