#!/bin/bash
# Run the staging / fuzz / lane-split GPU tests against a -DODE_B200_BOUNDS_CHECK build of the kernels (built
# here beforehand with:  tools/variant.sh boundscheck "-DODE_B200_BOUNDS_CHECK" inst_lorenz inst_kepler inst_cr3bp
# inst_robertson inst_bouncing_ball inst_threebody18 inst_tests ). A trap inside a kernel surfaces as a CUDA
# error of the launch, i.e. as failing tests.
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ODE_B200_LIB=variants/libode_boundscheck.so timeout 1500 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py \
    tests/test_gpu_lane_split.py tests/test_gpu_edge_cases.py tests/test_gpu_host_api.py -q > gpurun_out/boundscheck.log 2>&1
echo "boundscheck rc=$?" >> gpurun_out/boundscheck.log; tail -n 5 gpurun_out/boundscheck.log
