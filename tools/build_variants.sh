#!/bin/bash
# Builds one library per line of "name|nvcc flags" in $1 into _variants/<name>.so (git-ignored, travels with gpurun),
# then rebuilds the default library.   tools/build_variants.sh variants.txt
set -eu
mkdir -p _variants
while IFS='|' read -r name flags || [ -n "$name" ]; do
  [ -z "$name" ] && continue; [ "${name:0:1}" = "#" ] && continue
  CTR_EXTRA_NVCC_FLAGS="$flags" python constriction_b200/build.py --force > /dev/null 2>&1
  cp constriction_b200/libconstriction_b200.so _variants/$name.so
  echo "built $name ($flags)"
done < "$1"
python constriction_b200/build.py --force > /dev/null 2>&1
