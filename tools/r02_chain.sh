#!/bin/bash
out=gpurun_out/${1:-r02chain}; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_checkpoints.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
timeout 300 python bench_configs.py --configs 4,5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k not in ('workload', 'note', 'checkpoints_every_1024', 'checkpoints_every_128')})
"
