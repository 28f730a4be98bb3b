#!/bin/bash
# 2-GPU call: bench under torchrun (+ the single-process multi-GPU leg), reference arm under torchrun, multi-GPU tests
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2j_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_r2j_n2.json 2> $O/bench_r2j_n2.err; echo "bench rc=$?" >> $O/bench_r2j_n2.err
tail -n 5 $O/bench_r2j_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_r2j_ref2.json 2> $O/bench_r2j_ref2.err
timeout 600 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_pipeline.py -q -k "multi or pipeline" > $O/r2j_multi.log 2>&1; tail -n 3 $O/r2j_multi.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2j_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
c=d['co_headline']; print('kepler value',c['value'],c['ms_per_step'],'e2e',c['e2e']['value'],c['e2e']['ms_per_step'],'multi',c['e2e_multi'])
r=json.load(open('gpurun_out/bench_r2j_ref2.json')); print('ref',r['value'],r['cpu_baseline']['cores'])
PY
