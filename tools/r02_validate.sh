#!/bin/bash
# GPU box, 1 GPU: the round's evidence run.  tools/r02_validate.sh <tag>  -> gpurun_out/<tag>/
tag=${1:-r02_final}; out=gpurun_out/$tag; mkdir -p $out
timeout 120 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 240 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
timeout 120 python tools/readme_example.py > $out/readme_example.log 2>&1; tail -1 $out/readme_example.log
timeout 300 python bench.py > $out/bench_1gpu.json 2> $out/bench.err; cut -c1-300 $out/bench_1gpu.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2>&1
timeout 240 python bench_configs.py --configs 2,3,4,5,6,7 > $out/configs.jsonl 2> $out/configs.err; wc -l $out/configs.jsonl
timeout 120 python tools/mirror_latency.py > $out/mirror_latency.txt 2>&1; cat $out/mirror_latency.txt
timeout 100 python tools/time_chain.py > $out/chain_kernels.txt 2>&1; K=8192 timeout 100 python tools/time_chain.py >> $out/chain_kernels.txt 2>&1; CTR_CHAIN=0 timeout 100 python tools/time_chain.py >> $out/chain_kernels.txt 2>&1; cat $out/chain_kernels.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extra-configs > $out/b_ncu.log 2>&1
COUNT=2 SKIP=4 bash tools/r02_ncu.sh $tag/ncu_main "ans_(en|de)code_kernel" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --no-extra-configs > $out/ncu_main.log 2>&1
rm -f gpurun_out/$tag/ncu_main/rep.ncu-rep
ls -la $out
