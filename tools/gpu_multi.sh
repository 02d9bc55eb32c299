#!/bin/bash
# multi-GPU session: sharded front end check + strong-scaling bench at the box's GPU count.  Usage: gpurun --gpus N -- bash tools/gpu_multi.sh N tag
N=${1:-2}; tag=${2:-run}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/two_gpu_check.py > gpurun_out/multi_check_${N}_$tag.log 2>&1; echo "check rc=$?"
tail -3 gpurun_out/multi_check_${N}_$tag.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 $BENCH_FLAGS > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${N}gpu_$tag.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_${N}gpu_$tag.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N", d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "cfg", d["config"]["cluster_size"], d["config"]["threads"], d["config"]["clusters_in_flight"], "replicas", d.get("replicas",{}).get("value"))
print("kernel ms per rank", d.get("longest_solve_bound",{}).get("kernel_ms_per_rank"))
for k, v in d.get("workloads", {}).items():
    print(k, "value", round(v["value"],1), "ms/step", round(v["ms_per_step"],3), "e2e", round(v["e2e"]["value"],1), v.get("longest_solve_bound",{}).get("kernel_ms_per_rank"))
PY
