#!/bin/bash
# GPU call r2g: host pipeline (tests + e2e modes), bounds-check build under the fuzz, sanitized host layer
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -x > $O/r2g_pipe.log 2>&1; echo "pipe rc=$?" >> $O/r2g_pipe.log
tail -n 25 $O/r2g_pipe.log
timeout 600 python tools/e2e_modes.py > $O/r2g_e2e.log 2>&1; cat $O/r2g_e2e.log
timeout 1500 bash tools/gpu_boundscheck.sh
timeout 900 python -m pytest tests/test_sanitizers.py -q > $O/r2g_san.log 2>&1; tail -n 8 $O/r2g_san.log
timeout 1800 python -m pytest tests -m gpu -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log; tail -n 8 $O/r2g_pytest.log
