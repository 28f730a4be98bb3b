#!/bin/bash
# GPU call r2a: tests, ncu captures at the bench's sizes, bench (native + reference arm), Kepler/Dopri5 lockstep variants
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
rm -f $O/r2_ncu_summaries.json
timeout 600 bash tools/ncu_capture.sh bench_lorenz_dopri5_2p20 lorenz_dp5 1048576 $O/r2_ncu_summaries.json > $O/r2a_cap1.log 2>&1
timeout 900 bash tools/ncu_capture.sh bench_kepler_dop853_2p22 kepler_dp8 4194304 $O/r2_ncu_summaries.json > $O/r2a_cap2.log 2>&1
cp $O/r2_ncu_summaries.json profiles/r2_ncu_summaries.json
timeout 900 python bench.py > $O/bench_r2a_n1.json 2> $O/bench_r2a_n1.err; echo "bench rc=$?" >> $O/bench_r2a_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2a_ref.json 2> $O/bench_r2a_ref.err
for v in "" variants/libode_dp5lock384.so variants/libode_dp5lock256.so; do
  echo "== lib=$v" >> $O/r2a_kb_dp5lock.log
  ODE_B200_LIB=$v timeout 300 python tools/kbench.py kepler_dp5 kepler_dp8 >> $O/r2a_kb_dp5lock.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r2a_ncu_bench.log 2>&1
tail -n 5 $O/r2a_pytest.log; cat $O/bench_r2a_n1.err | tail -n 5; cat $O/r2a_kb_dp5lock.log
