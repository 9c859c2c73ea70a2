#!/bin/bash
# round 2, call A: new regime / golden tests, bench line with forest stats, forest ncu (trained forests, with source)
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_forest_regime.py tests/test_gpu_goldens.py -q -m gpu > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -25 $out/pytest_$tag.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_forest_predict_ranked' -s 38 -c 2 -f \
    -o $out/prof_forest1_$tag $B > $out/ncu_forest1_$tag.log 2>&1
ncu -i $out/prof_forest1_$tag.ncu-rep --page raw --csv > $out/prof_forest1_$tag.csv 2>/dev/null
ncu -i $out/prof_forest1_$tag.ncu-rep --page source --csv > $out/prof_forest1_${tag}_src.csv 2>/dev/null
find $out -size +30M -delete
ls -la $out | tail -8
