#!/bin/bash
# GPU box, N GPUs: multi-GPU parity test + bench.  tools/r02_run2.sh <tag> <N>
tag=${1:-r02m}; n=${2:-2}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $out/pytest_multi.log 2>&1; tail -15 $out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err; cat $out/bench_${n}gpu.json; tail -20 $out/bench_${n}gpu.err
