#!/bin/bash
# sweep: cluster size, threads, CTAs per SM, FFT workspace limit (KB)
mkdir -p gpurun_out; rm -f gpurun_out/c_*
for cfg in "8 128 4 38" "16 128 4 38" "8 128 4 20" "4 128 4 20"; do
  set -- $cfg
  BSGP_MINB=$3 BSGP_WS_KB=$4 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --cluster $1 --threads $2 > gpurun_out/c_$1_$2_$3_$4.json 2> gpurun_out/c_$1_$2_$3_$4.err || tail -2 gpurun_out/c_$1_$2_$3_$4.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 3), "clusters", d["config"]["clusters_in_flight"], "smem", d["config"]["smem_bytes"], "util", round(d["config"]["cluster_slot_utilisation"], 3), "maxit", d["config"]["max_iterations"])
    except Exception as e:
        print(f, "failed", e)
PY
