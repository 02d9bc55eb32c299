#!/bin/bash
# compare kernel build variants (csrc/libbsgp_<tag>.so) on the tiles256 workload with the default launch shape
mkdir -p gpurun_out; rm -f gpurun_out/v_*
for tag in "$@"; do
  BSGP_LIB=$PWD/beta-sgp_b200/csrc/libbsgp_$tag.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v_${tag}.json 2> gpurun_out/v_${tag}.err || tail -3 gpurun_out/v_${tag}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/v_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3), "util", round(d["config"]["cluster_slot_utilisation"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
