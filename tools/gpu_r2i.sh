#!/bin/bash
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2i_kb.log
timeout 300 python tools/kbench.py lorenz_dp5 lorenz_dp8 lorenz_rk4 kepler_dp8 kepler_dp5 threebody18_dp8 robertson_dp8 > $O/r2i_kb.log 2>&1
cat $O/r2i_kb.log
timeout 600 python tools/e2e_modes.py > $O/r2i_e2e.log 2>&1; cat $O/r2i_e2e.log
timeout 1800 python -m pytest tests -m gpu -q > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest.log; tail -n 6 $O/r2i_pytest.log
