#!/bin/bash
# GPU box, 1 GPU: tests + bench of the round-2 state.  tools/r02_run1.sh <tag>  -> gpurun_out/<tag>/
tag=${1:-r02a}; out=gpurun_out/$tag; mkdir -p $out
timeout 120 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -15 $out/pytest.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; cat $out/bench.json; tail -5 $out/bench.err
