#!/bin/bash
# stamp / cut-out workloads under other CTA configurations; logged to gpurun_out/stamps_cfg.log
mkdir -p gpurun_out
exec > >(tee gpurun_out/stamps_cfg.log) 2>&1
run() { python bench.py --no-extra --no-cpu-baseline --no-clocks --steps 3 "$@" 2>gpurun_out/stamps_cfg_err.log | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value'],1), 'cfg', d['config']['cluster_size'], d['config']['threads'], d['config']['clusters_in_flight'], d['config']['smem_bytes'])
" || tail -3 gpurun_out/stamps_cfg_err.log; }
for w in stamps32 cutouts31; do
echo "$w default:"; run --workload $w
echo "$w 128 thr (3 CTAs/SM):"; run --workload $w --threads 128
done
