#!/bin/bash
# ncu captures of the forest kernel only (all launches of one pass + one launch with source).
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
B="python bench.py --forests random --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_forest_predict_ranked' -s 38 -c 1 -f \
    -o $out/prof_forest1_$tag $B > $out/ncu_forest1_$tag.log 2>&1
ncu -i $out/prof_forest1_$tag.ncu-rep --page raw --csv > $out/prof_forest1_$tag.csv 2>/dev/null
ncu -i $out/prof_forest1_$tag.ncu-rep --page source --csv > $out/prof_forest1_${tag}_src.csv 2>/dev/null
timeout 1200 ncu --set full --clock-control none -k regex:'k_forest_predict_ranked' -s 32 -c 32 -f \
    -o $out/prof_forest_$tag $B > $out/ncu_forest_$tag.log 2>&1
ncu -i $out/prof_forest_$tag.ncu-rep --page raw --csv > $out/prof_forest_$tag.csv 2>/dev/null
rm -f $out/prof_forest_$tag.ncu-rep
find $out -size +20M -delete
ls -la $out | tail
