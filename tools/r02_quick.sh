#!/bin/bash
# quick A/B of the headline kernels: tools/r02_quick.sh  -> prints ms per step and kernel times (3 runs)
for i in 1 2 3; do python bench.py --no-extra-configs --e2e-steps 0 --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['roofline']['kernel_ms'].items()}, round(d['roofline']['frac'], 4))"; done
