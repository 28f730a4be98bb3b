#!/bin/bash
# GPU call r2h: pipeline with deferred signalling, scheduler exit fix, bounds-check build, sanitized host layer, suite
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_edge_cases.py -q -x > $O/r2h_pipe.log 2>&1; echo "pipe rc=$?" >> $O/r2h_pipe.log
tail -n 12 $O/r2h_pipe.log
timeout 600 python tools/e2e_modes.py > $O/r2h_e2e.log 2>&1; cat $O/r2h_e2e.log
timeout 1500 bash tools/gpu_boundscheck.sh
timeout 900 python -m pytest tests/test_sanitizers.py -q > $O/r2h_san.log 2>&1; tail -n 4 $O/r2h_san.log
timeout 1800 python -m pytest tests -m gpu -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log; tail -n 6 $O/r2h_pytest.log
timeout 300 python tools/kbench.py lorenz_dp5 kepler_dp8 kepler_dp5 threebody18_dp8 > $O/r2h_kb.log 2>&1; cat $O/r2h_kb.log
