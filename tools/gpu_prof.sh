#!/bin/bash
# one ncu --set full capture of the solve kernel + per-function summary.  usage: gpurun -- bash tools/gpu_prof.sh <tag> [workload]
tag=${1:-prof}; wl=${2:-tiles256}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --workload $wl > gpurun_out/plain_$tag.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-bsgp_solve} -s ${KSKIP:-3} -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --workload $wl > gpurun_out/ncu_$tag.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep beta-sgp_b200/csrc/libbsgp.so ${3:-bsgp_solve_kernelIdLi128ELi4ELb0} "ncu --set full, $wl, round 2 ($tag)" > gpurun_out/prof_${tag}_summary.txt 2>&1
cat gpurun_out/prof_${tag}_summary.txt | cut -c1-180
rm -f gpurun_out/prof_$tag.ncu-rep.tmp
