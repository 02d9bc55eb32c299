#!/bin/bash
# Round-end GPU session: parity suite, the three bench workloads, ncu launch list and one full capture of the solve kernel
# (each profiler pass only after its plain command exited 0).  usage: gpurun --timeout 1500 -- bash tools/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tiles_$tag.json 2> gpurun_out/bench_tiles_$tag.err; echo "bench tiles rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "bench ref rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --workload stamps32 --no-cpu-baseline > gpurun_out/bench_stamps_$tag.json 2> gpurun_out/bench_stamps_$tag.err; echo "bench stamps rc=$?"
timeout 400 python bench.py --steps 2 --warmup 3 --workload frame --no-cpu-baseline > gpurun_out/bench_frame_$tag.json 2> gpurun_out/bench_frame_$tag.err; echo "bench frame rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsgp_solve -s 3 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for f in ("tiles", "stamps", "frame", "ref"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{f}_$tag.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "ms/step", round(d.get("ms_per_step", 0), 2), "frac", d.get("roofline", {}).get("frac"), "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
