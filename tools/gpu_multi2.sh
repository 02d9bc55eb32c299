#!/bin/bash
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --no-extra --no-cpu-baseline "${@:2}" 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'kernel', [round(v,2) for v in d['longest_solve_bound']['kernel_ms_per_rank']], 'e2e', round(d['e2e']['value'],1))
"; }
echo "as is:"; run 29701 --steps 5 --warmup 3
echo "no clocks:"; run 29702 --steps 5 --warmup 3 --no-clocks
echo "10 steps, 6 warmup, no clocks:"; run 29703 --steps 10 --warmup 6 --no-clocks
