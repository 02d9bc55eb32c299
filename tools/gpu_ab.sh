#!/bin/bash
# A/B of library builds on the three bench workloads.  usage: gpurun -- bash tools/gpu_ab.sh lib1.so lib2.so ...   (paths relative to beta-sgp_b200/csrc)
mkdir -p gpurun_out
for lib in "$@"; do
  for wl in tiles256 stamps32 frame; do
    st=3; [ $wl = frame ] && st=2
    BSGP_LIB=$PWD/beta-sgp_b200/csrc/$lib python bench.py --no-extra --no-cpu-baseline --no-clocks --steps $st --workload $wl 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '$wl', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value'],2), 'frac', round(d['roofline']['frac'],3))
"
  done
done
