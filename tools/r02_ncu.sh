#!/bin/bash
# tools/r02_ncu.sh <tag> <kernel regex> <python script...>: one ncu --set full capture with source, summarised
tag=$1; regex=$2; shift 2; out=gpurun_out/$tag; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s ${SKIP:-2} -c ${COUNT:-2} -f -o $out/rep "$@" > $out/ncu.log 2>&1
tail -3 $out/ncu.log
ncu -i $out/rep.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
python tools/ncu_summary.py $out/raw.csv > $out/summary.txt; cat $out/summary.txt
ncu -i $out/rep.ncu-rep --page source --csv --print-source cuda,sass > $out/src.csv 2>/dev/null
python tools/ncu_source_hot.py $out/src.csv 22 > $out/hot.txt; cat $out/hot.txt
rm -f $out/src.csv
ncu -i $out/rep.ncu-rep --page source --csv --print-source sass > $out/sass.csv 2>/dev/null
python tools/ncu_sass_hot.py $out/sass.csv ${TOP:-60} > $out/sass_hot.txt; head -8 $out/sass_hot.txt
rm -f $out/sass.csv
