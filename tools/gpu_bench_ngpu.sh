#!/bin/bash
# 8-GPU call: bench under torchrun at N = 8 (headline weak scaling, Kepler 4M strong scaling, single-process multi-GPU leg)
cd $GRAFT_REPO_ROOT 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
nvidia-smi -L > $O/ngpu_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_ngpu_n$N.json 2> $O/bench_ngpu_n$N.err; echo "bench rc=$?" >> $O/bench_ngpu_n$N.err
tail -n 3 $O/bench_ngpu_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.load(open('gpurun_out/bench_ngpu_n%s.json'%N))
print('N',N,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
c=d['co_headline']; print('kepler value',c['value'],c['ms_per_step'],'e2e',c['e2e']['value'],c['e2e']['ms_per_step'],'multi',c['e2e_multi'].get('ms_per_step'), c['e2e_multi'].get('error'))
PY
