#!/bin/bash
# compare kernel build variants (csrc/libbsgp_<tag>.so) on the tiles256 workload
mkdir -p gpurun_out
for tag in "$@"; do
  for th in 512 256; do
    BSGP_LIB=$PWD/beta-sgp_b200/csrc/libbsgp_$tag.so python bench.py --steps 2 --warmup 2 --no-cpu-baseline --threads $th > gpurun_out/v_${tag}_$th.json 2> gpurun_out/v_${tag}_$th.err || tail -3 gpurun_out/v_${tag}_$th.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/v_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
