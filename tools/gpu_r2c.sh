#!/bin/bash
# round 2, session c: parity suite + benches after the active-set projection
tag=${1:-r2c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -8 gpurun_out/pytest_gpu_$tag.log
python bench.py --no-extra --no-cpu-baseline --steps 3 > gpurun_out/bench_tiles_$tag.json 2> gpurun_out/bench_tiles_$tag.err; echo "bench rc=$?"
python bench.py --no-extra --no-cpu-baseline --steps 3 --workload stamps32 > gpurun_out/bench_stamps_$tag.json 2> gpurun_out/bench_stamps_$tag.err; echo "bench rc=$?"
python bench.py --no-extra --no-cpu-baseline --steps 2 --workload frame > gpurun_out/bench_frame_$tag.json 2> gpurun_out/bench_frame_$tag.err; echo "bench rc=$?"
python - <<PY
import json
for w in ("tiles", "stamps", "frame"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{w}_$tag.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],2), "slot_util", round(d["config"]["cluster_slot_utilisation"],3))
    except Exception as e:
        print(w, "failed", e)
PY
timeout 300 python tools/latency_probe.py tiles 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['B'] in (1, 40, 80, 320): print('probe', d['G'], d['threads'], d['slots'], 'B', d['B'], 'ms', round(d['ms'], 2), 'us/it longest', round(d['us_per_it_longest'], 1))
"
