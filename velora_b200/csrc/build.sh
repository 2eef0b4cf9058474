#!/bin/bash
# Builds libvelora_b200.so in-tree for sm_100a (cross-compiles without a GPU).  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
OUT=${VB_OUT:-../lib}       # VB_OUT=../lib_trace VB_TRACE_BUILD=1 builds the instrumented copy next to the product library
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
# VB_TRACE_BUILD=1 compiles the clock64 phase instrumentation of the tcgen05 backward kernel in (profiles/microbench/trace_run.py)
# VB_EXTRA_FLAGS: experiment switches (e.g. -DVB_FWD_PIPELINED=0) for a side build under VB_OUT
FLAGS="$FLAGS $VB_EXTRA_FLAGS"
if [ -n "$VB_TRACE_BUILD" ]; then FLAGS="$FLAGS -DVB_ENABLE_TRACE"; touch liquid_tc.cu; fi
SRCS="api.cu pack.cu liquid_small.cu liquid_tc.cu generic.cu gather.cu optim.cu ncp_tc.cu liquid_big.cu liquid_big_bwd.cu collective.cu sac_loss.cu"
OBJS=""
for s in $SRCS; do
  o="$OUT/${s%.cu}.o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tile.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ big_common.cuh -nt "$o" ] || [ ../../include/velora_b200.h -nt "$o" ]; then
    echo "nvcc $s"
    $NVCC $FLAGS -c "$s" -o "$o" 2> "$OUT/${s%.cu}.ptxas.log" || { cat "$OUT/${s%.cu}.ptxas.log"; exit 1; }
  fi
  OBJS="$OBJS $o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libvelora_b200.so" $OBJS -lcudart -lcublas -Xlinker -rpath=/usr/local/cuda/lib64
echo "built $OUT/libvelora_b200.so"
