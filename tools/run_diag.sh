#!/bin/bash
# GPU box: kernel-only timings of diagnostic builds (one line of nvcc flags per line of $1); tools/run_diag.sh flags.txt tag [args]
set -u
flags_file=$1; tag=$2; shift 2
out=gpurun_out/variants; mkdir -p $out
: > $out/$tag.jsonl
while IFS= read -r flags || [ -n "$flags" ]; do
  [ "${flags:0:1}" = "#" ] && continue
  export CTR_EXTRA_NVCC_FLAGS="$flags"
  python constriction_b200/build.py --force > $out/build.log 2>&1 || { echo "{\"flags\": \"$flags\", \"error\": \"build\"}" >> $out/$tag.jsonl; continue; }
  timeout 200 python tools/time_kernels.py "$@" >> $out/$tag.jsonl 2>$out/time.err || echo "{\"flags\": \"$flags\", \"error\": \"run\"}" >> $out/$tag.jsonl
done < "$flags_file"
cat $out/$tag.jsonl
