#!/bin/bash
# Evidence pass (run on the GPU box: gpurun -- bash tools/gpu_evidence.sh): full GPU suite, ncu captures of the final kernels, bench (native + reference arm)
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
J=$O/r2_ncu_summaries.json
cp profiles/r2_ncu_summaries.json $J 2>/dev/null
timeout 1800 python -m pytest tests -m gpu -q > $O/ev_pytest.log 2>&1; echo "pytest rc=$?" >> $O/ev_pytest.log; tail -n 4 $O/ev_pytest.log
timeout 600 bash tools/ncu_capture.sh bench_lorenz_dopri5_2p20 lorenz_dp5 1048576 $J > $O/ev_c1.log 2>&1
timeout 900 bash tools/ncu_capture.sh bench_kepler_dop853_2p22 kepler_dp8 4194304 $J > $O/ev_c2.log 2>&1
timeout 600 bash tools/ncu_capture.sh split_threebody18_dop853_sparse_2p18 threebody18_dp8 262144 $J > $O/ev_c3.log 2>&1
timeout 600 bash tools/ncu_capture.sh split_threebody18_dop853_cont8_2p18 threebody18_dp8_com 262144 $J > $O/ev_c4.log 2>&1
timeout 600 bash tools/ncu_capture.sh split_threebody18_dopri5_sparse_2p18 threebody18_dp5 262144 $J > $O/ev_c4a.log 2>&1
timeout 600 bash tools/ncu_capture.sh split_threebody18_dopri5_cont_2p18 threebody18_dp5_com 262144 $J > $O/ev_c4b.log 2>&1
timeout 600 bash tools/ncu_capture.sh kepler_dopri5_2p20 kepler_dp5 1048576 $J > $O/ev_c5.log 2>&1
timeout 600 bash tools/ncu_capture.sh lorenz_rk4_all_steps_2p18 lorenz_rk4_all_steps 262144 $J > $O/ev_c6.log 2>&1
timeout 600 bash tools/ncu_capture.sh lorenz_dopri5_dense_2001_2p18 lorenz_dp5_dense_hbm 262144 $J > $O/ev_c7.log 2>&1
# the ContinuousOutputModel::evaluate kernel (second launch = the 2001-query uniform grid, first timed repetition)
KBENCH_N=131072 python tools/kbench_com.py > $O/ev_com.log 2>&1
KBENCH_N=131072 ncu --set full --clock-control none -k regex:com_evaluate -c 1 -f -o $O/com_eval python tools/kbench_com.py > $O/ev_ncu_com.log 2>&1
ncu -i $O/com_eval.ncu-rep --page raw --csv > $O/com_evaluate_2001q_2p17_raw.csv; rm -f $O/com_eval.ncu-rep
python tools/ncu_summary.py $J com_evaluate_2001q_2p17=$O/com_evaluate_2001q_2p17_raw.csv
cp $J profiles/r2_ncu_summaries.json
timeout 900 python bench.py > $O/bench_ev_n1.json 2> $O/bench_ev_n1.err; echo "bench rc=$?" >> $O/bench_ev_n1.err; tail -n 3 $O/bench_ev_n1.err
timeout 600 python bench.py --impl reference > $O/bench_ev_ref.json 2> $O/bench_ev_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-pipeline > $O/ev_ncu_bench.log 2>&1
cat $O/ev_com.log; ls -la $O | head -40
