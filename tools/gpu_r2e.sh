#!/bin/bash
# GPU call r2e: f32 adaptive steppers, lane-split v3 (streamed Continuous8 rows), whole suite
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_f32.py -q -x > $O/r2e_f32.log 2>&1; echo "f32 rc=$?" >> $O/r2e_f32.log
tail -n 25 $O/r2e_f32.log
timeout 900 python -m pytest tests/test_gpu_lane_split.py -q -x > $O/r2e_split.log 2>&1; echo "split rc=$?" >> $O/r2e_split.log
tail -n 8 $O/r2e_split.log
timeout 300 python tools/kbench.py threebody18_dp8 threebody18_dp8_com > $O/r2e_kb.log 2>&1
KBENCH_N=65536 timeout 300 python tools/kbench.py threebody18_dp8_com >> $O/r2e_kb.log 2>&1
cat $O/r2e_kb.log
timeout 600 bash tools/ncu_capture.sh split3_threebody18_dop853_cont8 threebody18_dp8_com 262144 $O/r2e_ncu.json > $O/r2e_cap2.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log; tail -n 8 $O/r2e_pytest.log
