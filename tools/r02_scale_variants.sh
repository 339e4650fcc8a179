#!/bin/bash
# N GPUs: bench.py variants of the gather (push streams, lag).  tools/r02_scale_variants.sh <tag> <N>
tag=${1:-r02v}; n=${2:-4}; out=gpurun_out/$tag; mkdir -p $out
run() { name=$1; shift; env "$@" CTR_BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --no-extra-configs --e2e-steps 0 --no-cpu-baseline --steps 20 > $out/$name.json 2> $out/$name.err; python - <<PY
import json
try:
    d=json.load(open("$out/$name.json")); print("$name", d["ms_per_step"], d["host_issue_ms_per_step"], d["roofline"]["kernel_ms"])
except Exception as e:
    print("$name failed", e); print(open("$out/$name.err").read()[-1500:])
PY
grep "steps(ms)" $out/$name.err | cut -c1-230
}
run p1_lag1_a CTR_PUSH_STREAMS=1 CTR_GATHER_LAG=1
run p4_lag1_a CTR_PUSH_STREAMS=4 CTR_GATHER_LAG=1
run p4_lag2_a CTR_PUSH_STREAMS=4 CTR_GATHER_LAG=2
