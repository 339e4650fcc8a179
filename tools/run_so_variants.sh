#!/bin/bash
# GPU box: times the headline kernels with every prebuilt library in _variants/ (tools/build_variants.sh).
#   tools/run_so_variants.sh <tag> [names...]   -> gpurun_out/<tag>/variants.jsonl
set -u
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
cp constriction_b200/libconstriction_b200.so /tmp/default.so
names=${@:-$(ls _variants | sed 's/\.so$//')}
for v in $names; do
  cp _variants/$v.so constriction_b200/libconstriction_b200.so
  if [ "${TESTS:-1}" = "1" ]; then timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $out/pytest_$v.log 2>&1; t=$?; else t=-1; fi
  r=$(REPS=${REPS:-20} timeout 200 python tools/time_kernels.py 2>$out/err_$v.log)
  echo "{\"variant\": \"$v\", \"tests_rc\": $t, \"res\": ${r:-null}}" | tee -a $out/variants.jsonl
done
cp /tmp/default.so constriction_b200/libconstriction_b200.so
