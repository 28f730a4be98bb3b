#!/bin/bash
# GPU call r2d: lane-split v2 (select-free RHS, staged outputs): tests, kbench, ncu
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane_split.py -q -x > $O/r2d_split.log 2>&1; echo "split rc=$?" >> $O/r2d_split.log
tail -n 15 $O/r2d_split.log
timeout 300 python tools/kbench.py threebody18_dp8 threebody18_dp8_com > $O/r2d_kb.log 2>&1
KBENCH_N=65536 timeout 300 python tools/kbench.py threebody18_dp8_com >> $O/r2d_kb.log 2>&1
cat $O/r2d_kb.log
NCU_SOURCE=1 timeout 600 bash tools/ncu_capture.sh split2_threebody18_dop853_sparse threebody18_dp8 262144 $O/r2d_ncu.json > $O/r2d_cap1.log 2>&1
NCU_SOURCE=1 timeout 600 bash tools/ncu_capture.sh split2_threebody18_dop853_cont8 threebody18_dp8_com 262144 $O/r2d_ncu.json > $O/r2d_cap2.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py -q > $O/r2d_pytest.log 2>&1; tail -n 5 $O/r2d_pytest.log
