#!/bin/bash
# round-2 final single-GPU session: parity suite, smoke, default bench (driver's command), reference arm, ncu launch list
tag=${1:-r2final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_default_$tag.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_default_$tag.json").read().strip().splitlines()[-1])
print("main", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "launches", d["gpu_launches"], "cpu", d.get("cpu_baseline",{}).get("value"), "clocks", d["clocks"])
for k, v in d.get("workloads", {}).items():
    if "error" in v: print(k, v); continue
    print(k, "value", round(v["value"],2), "ms/step", round(v["ms_per_step"],3), "ms/it(longest)", round(v["ms_per_iteration_longest_solve"],4), "frac", round(v["roofline"]["frac"],3), "e2e", round(v["e2e"]["value"],2), "cpu", v.get("cpu_baseline",{}).get("value"), "cfg", v["cluster_size"], v["threads"])
r = json.loads(open("gpurun_out/bench_ref_$tag.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
