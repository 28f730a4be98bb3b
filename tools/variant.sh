#!/bin/bash
# Build an experimental variant of libode_b200.so: recompile only the named translation units with extra
# flags and link them with the default build's other objects.
# usage: tools/variant.sh <name> "<extra nvcc flags>" <unit> [<unit> ...]     (unit: inst_kepler, ode_b200, ...)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/ode_solvers_b200/csrc
NAME=$1; FLAGS=$2; shift 2
B=$ROOT/variants/build_$NAME
mkdir -p $B
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -prec-div=true -prec-sqrt=true -lineinfo --std=c++17 -I$ROOT/include -I$SRC -Xcompiler -fPIC -Xcompiler -ffp-contract=off -diag-suppress 222 $FLAGS"
pids=()
for u in "$@"; do
  if [ -f $SRC/$u.cu ]; then $NV -c $SRC/$u.cu -o $B/$u.o & else $NV -x cu -c $SRC/$u.cpp -o $B/$u.o & fi
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for o in $SRC/build/*.o; do
  b=$(basename $o)
  if [ -f $B/$b ]; then OBJS="$OBJS $B/$b"; else OBJS="$OBJS $o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/libode_$NAME.so $OBJS -cudart static -ldl -lpthread $VARIANT_LINK_FLAGS
echo built $ROOT/variants/libode_$NAME.so
