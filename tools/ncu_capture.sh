#!/bin/bash
# One `ncu --set full` capture of the step-loop kernel of a tools/kbench.py workload (run on the GPU box
# under gpurun, one GPU, only after the same command has exited 0 without ncu), exported as raw CSV into
# gpurun_out/ and merged into a summaries JSON.
# usage: tools/ncu_capture.sh <label> <kbench workload> <n trajectories> [summaries.json]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LABEL=$1; WL=$2; N=$3; OUT=${4:-$ROOT/gpurun_out/r2_ncu_summaries.json}
mkdir -p $ROOT/gpurun_out
cd $ROOT
KBENCH_N=$N KBENCH_REPS=1 python tools/kbench.py $WL > gpurun_out/plain_$LABEL.log 2>&1
# -k regex: the step loop only; -s 0 -c 1: the first launch (kbench runs one with KBENCH_REPS=1)
KBENCH_N=$N KBENCH_REPS=1 ncu --set full --clock-control none --import-source on -k regex:step_loop -c 1 \
    -f -o gpurun_out/$LABEL python tools/kbench.py $WL > gpurun_out/ncu_$LABEL.log 2>&1
ncu -i gpurun_out/$LABEL.ncu-rep --page raw --csv > gpurun_out/${LABEL}_raw.csv
if [ -n "$NCU_SOURCE" ]; then  # per-instruction view (SASS + source lines), compressed
  ncu -i gpurun_out/$LABEL.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${LABEL}_src.csv.gz
fi
# gpurun copies back at most 64 MiB per call: the report itself stays on the box unless asked for
[ -n "$NCU_KEEP_REP" ] || rm -f gpurun_out/$LABEL.ncu-rep
python tools/ncu_summary.py $OUT $LABEL=gpurun_out/${LABEL}_raw.csv
echo "captured $LABEL -> $OUT"
