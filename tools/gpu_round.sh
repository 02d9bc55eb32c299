#!/bin/bash
# One GPU session: parity tests, then benches (no profiler).  Usage: gpurun -- bash tools/gpu_round.sh [tag]
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -5 gpurun_out/pytest_gpu_$tag.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tiles_$tag.json 2> gpurun_out/bench_tiles_$tag.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --workload stamps32 --no-cpu-baseline > gpurun_out/bench_stamps_$tag.json 2> gpurun_out/bench_stamps_$tag.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_tiles_$tag.json", "gpurun_out/bench_stamps_$tag.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "img/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "us/img-iter", round(1e3 * d["ms_per_image_iteration"], 3), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1), d["config"]["clusters_in_flight"], d["config"]["threads"], d["config"]["smem_bytes"])
    except Exception as e:
        print(f, "failed", e)
PY
