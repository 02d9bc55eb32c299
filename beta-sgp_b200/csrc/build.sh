#!/bin/bash
# Build libbsgp.so in-tree for B200 (sm_100a).  Used by __graft_entry__.build().
# Five translation units (C ABI + small kernels; fp64 / fp32 solver, each for the circular and for the
# zero-padded operator) compile in parallel.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# the library reports the hash of the sources it was built from (bsgp_version()); __graft_entry__.build() checks it,
# so a stale binary cannot run unnoticed
HASH=$(cat $(ls *.cu *.cuh *.h ../../include/bsgp.h | LC_ALL=C sort) | sha256sum | cut -c1-16)
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DBSGP_SRC_HASH=\"$HASH\" $BSGP_NVCC_FLAGS"
OUT=${BSGP_OUT:-libbsgp.so}
B=build${BSGP_TAG:+_$BSGP_TAG}
mkdir -p $B
pids=()
for tu in bsgp_kernels bsgp_solve_f64 bsgp_solve_f32 bsgp_solve_f64_padded bsgp_solve_f32_padded; do
    $NVCC $FLAGS -c -o $B/$tu.o $tu.cu > $B/$tu.log 2>&1 &
    pids+=($!)
done
rc=0
for pid in "${pids[@]}"; do wait $pid || rc=1; done
cat $B/*.log
[ $rc -eq 0 ] || exit 1
$NVCC -shared -o $OUT $B/bsgp_kernels.o $B/bsgp_solve_f64.o $B/bsgp_solve_f32.o $B/bsgp_solve_f64_padded.o $B/bsgp_solve_f32_padded.o
