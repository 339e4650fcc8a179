#!/bin/bash
# Runs on the GPU box: builds the library once per line of nvcc flags in $1, runs the GPU parity tests and a short
# bench, and appends one JSON line per variant to gpurun_out/variants/<tag>.jsonl.   tools/run_variants.sh flags.txt tag
set -u
flags_file=$1; tag=$2
out=gpurun_out/variants; mkdir -p $out
: > $out/$tag.jsonl
while IFS= read -r flags || [ -n "$flags" ]; do
  [ "${flags:0:1}" = "#" ] && continue
  CTR_EXTRA_NVCC_FLAGS="$flags" python constriction_b200/build.py --force > $out/build.log 2>&1 || { echo "{\"flags\": \"$flags\", \"error\": \"build\"}" >> $out/$tag.jsonl; continue; }
  if [ "${QUICK_TESTS:-1}" = "1" ]; then
    timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $out/pytest.log 2>&1; t=$?
  else t=-1; fi
  line=$(timeout 200 python bench.py --no-cpu-baseline --e2e-steps 0 --steps 20 2>$out/bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'], 4), 'kernel_ms': {k: round(v, 4) for k, v in d['roofline']['kernel_ms'].items()}, 'frac': round(d['roofline']['frac'], 4), 'sm_mhz': d['clocks']['sm_mhz']}))")
  echo "{\"flags\": \"$flags\", \"tests_rc\": $t, \"bench\": ${line:-null}}" >> $out/$tag.jsonl
done < "$flags_file"
cat $out/$tag.jsonl
