#!/bin/bash
# round 2, first GPU session: parity suite, baseline benches, latency probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_r2a.log
tail -15 gpurun_out/pytest_gpu_r2a.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiles_r2a.json 2> gpurun_out/bench_tiles_r2a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_tiles_r2a.json
timeout 600 python tools/latency_probe.py tiles > gpurun_out/latency_tiles_r2a.jsonl 2> gpurun_out/latency_tiles_r2a.err; echo "probe rc=$?"
cat gpurun_out/latency_tiles_r2a.jsonl
timeout 300 python tools/latency_probe.py stamps > gpurun_out/latency_stamps_r2a.jsonl 2> gpurun_out/latency_stamps_r2a.err; echo "probe rc=$?"
cat gpurun_out/latency_stamps_r2a.jsonl
