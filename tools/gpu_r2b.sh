#!/bin/bash
# GPU call r2b: lane-split tests first, whole GPU suite, kbench split vs per-thread for 3-body-18
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane_split.py -q -x > $O/r2b_split.log 2>&1; echo "split rc=$?" >> $O/r2b_split.log
tail -n 30 $O/r2b_split.log
for e in "" 1; do
  echo "== KBENCH_NOSPLIT=$e" >> $O/r2b_kb.log
  KBENCH_NOSPLIT=$e timeout 300 python tools/kbench.py threebody18_dp8 threebody18_dp8_com >> $O/r2b_kb.log 2>&1
done
cat $O/r2b_kb.log
timeout 1800 python -m pytest tests -m gpu -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
tail -n 15 $O/r2b_pytest.log
