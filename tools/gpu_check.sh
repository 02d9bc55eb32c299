#!/bin/bash
# parity suite + short runs of the workloads named in $WL (default: tiles256 stamps32 cutouts31); everything is logged to gpurun_out/check_$1.log
tag=${1:-x}
mkdir -p gpurun_out
exec > >(tee gpurun_out/check_$tag.log) 2>&1
run() { timeout 600 python bench.py --no-extra --no-cpu-baseline --no-clocks --steps ${STEPS:-3} --workload "$@" 2>gpurun_out/check_err_$tag.log | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cfg', d['config']['cluster_size'], d['config']['threads'], d['config']['clusters_in_flight'], d['config']['smem_bytes'])
" || grep -v "^frame" gpurun_out/check_err_$tag.log | tail -5; }
if [ -z "$NOTEST" ]; then timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log; fi
for w in ${WL:-tiles256 stamps32 cutouts31}; do echo "== $w"; run $w; done
