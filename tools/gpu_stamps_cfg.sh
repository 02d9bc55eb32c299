#!/bin/bash
run() { python bench.py --no-extra --no-cpu-baseline --no-clocks --steps 3 --workload stamps32 "$@" 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value'],1), 'cfg', d['config']['cluster_size'], d['config']['threads'], d['config']['clusters_in_flight'], d['config']['smem_bytes'])
"; }
echo "default:"; run
echo "128 thr (minb 3):"; run --threads 128
echo "128 thr minb 4:"; BSGP_MINB=4 run --threads 128
echo "256 thr minb 1:"; BSGP_MINB=1 run --threads 256
