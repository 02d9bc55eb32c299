#!/bin/bash
# quick perf check: tiles256 at several thread configurations + stamps
mkdir -p gpurun_out
for th in 512 256; do
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline --threads $th > gpurun_out/q_tiles_$th.json 2> gpurun_out/q_tiles_$th.err || tail -3 gpurun_out/q_tiles_$th.err
done
for th in 256 128; do
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline --workload stamps32 --threads $th > gpurun_out/q_stamps_$th.json 2> gpurun_out/q_stamps_$th.err || tail -3 gpurun_out/q_stamps_$th.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/q_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "img/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "us/img-iter", round(1e3 * d["ms_per_image_iteration"], 3), "frac", round(d["roofline"]["frac"], 3), "clusters", d["config"]["clusters_in_flight"], "thr", d["config"]["threads"], "smem", d["config"]["smem_bytes"])
    except Exception as e:
        print(f, "failed", e)
PY
