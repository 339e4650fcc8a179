#!/bin/bash
# GPU box: the round's evidence run.  tools/validate.sh <tag>  -> gpurun_out/<tag>/
tag=${1:-val}; out=gpurun_out/$tag; mkdir -p $out
timeout 120 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -3 $out/pytest.log
timeout 300 python tools/readme_example.py > $out/readme_example.log 2>&1; tail -1 $out/readme_example.log
timeout 300 python bench.py > $out/bench.json 2> $out/bench.err; cat $out/bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2>&1
timeout 400 python bench_configs.py --configs 2,3,4,5,6 > $out/configs.jsonl 2> $out/configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $out/b_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ans_(en|de)code_kernel" -s 2 -c 2 -o $out/ans_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > $out/ncu_full.log 2>&1
ls -la $out
