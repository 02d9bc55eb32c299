#!/bin/bash
# A/B of library builds on the same box: csrc/libbsgp_<tag>.so, alternating runs.  usage: gpu_ab.sh <reps> <workloads> tag...
reps=$1; wl=$2; shift 2
mkdir -p gpurun_out
for rep in $(seq 1 $reps); do
  for tag in "$@"; do
    for w in $wl; do
      BSGP_LIB=$PWD/beta-sgp_b200/csrc/libbsgp_$tag.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/ab_${w}_${tag}_$rep.json 2>/dev/null
    done
  done
done
python - <<PY
import json, glob, collections
acc = collections.defaultdict(list)
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); acc[f.split("/")[-1].rsplit("_", 1)[0]].append(round(d["ms_per_step"], 2))
    except Exception as e: print(f, "failed", e)
for k, v in acc.items(): print(k, v, "min", min(v))
PY
