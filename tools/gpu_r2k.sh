#!/bin/bash
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_api.py -q > $O/r2k_api.log 2>&1; tail -n 4 $O/r2k_api.log
timeout 600 python tools/kbench_com.py > $O/r2k_com.log 2>&1; cat $O/r2k_com.log
KBENCH_N=131072 ncu --set full --clock-control none -k regex:com_evaluate -c 1 -f -o $O/com_eval python tools/kbench_com.py > $O/r2k_ncu_com.log 2>&1
ncu -i $O/com_eval.ncu-rep --page raw --csv > $O/com_evaluate_2001q_2p17_raw.csv; rm -f $O/com_eval.ncu-rep
cp profiles/r2_ncu_summaries.json $O/r2_ncu_summaries.json 2>/dev/null
python tools/ncu_summary.py $O/r2_ncu_summaries.json com_evaluate_2001q_2p17=$O/com_evaluate_2001q_2p17_raw.csv
