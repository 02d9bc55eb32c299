#!/bin/bash
for v in 5 4 3 2 1; do
  BSGP_FRAME_LGCP=$v python bench.py --no-extra --no-cpu-baseline --no-clocks --steps 2 --workload frame 2>/dev/null | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lg_cp', $v, 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))
"
done
python -m pytest tests -m gpu -q -x -k "frame" 2>&1 | tail -2
BSGP_FRAME_LGCP=2 python -m pytest tests -m gpu -q -x -k "frame" 2>&1 | tail -2
