#!/bin/bash
# One GPU box: parity suite, bench line, ncu launch list + full captures of the main kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_profile.sh <tag> [quick]
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
B="python bench.py --forests random --steps 3 --warmup 1 --no-e2e --no-cpu-baseline"
for lay in wide16 bytes; do
  DR_RANKED_LAYOUT=$lay $B > $out/bench_${tag}_$lay.json 2> $out/bench_${tag}_$lay.err; echo "bench $lay exit $?"
done
python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
[ "$2" = quick ] && exit 0
B="python bench.py --forests random --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 3000 --csv \
    --log-file $out/launches_$tag.csv $B > $out/ncu_launch_$tag.log 2>&1
# (kernel regex, launches of the untimed set-up to skip, launches to capture)
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f \
      -o $out/prof_$1_$tag $B > $out/ncu_$1_$tag.log 2>&1
  ncu -i $out/prof_$1_$tag.ncu-rep --page raw --csv > $out/prof_$1_$tag.csv 2>/dev/null
}
cap scan 'k_scan_hist' 1 1
cap pairs 'k_pairs' 2 2
cap gather 'k_gather_rows_masked' 32 1
cap domain 'k_domain_score' 16 2
cap forest1 'k_forest_predict_ranked' 38 1
timeout 1200 ncu --set full --clock-control none -k regex:'k_forest_predict_ranked' -s 32 -c 32 -f \
    -o $out/prof_forest_$tag $B > $out/ncu_forest_$tag.log 2>&1
ncu -i $out/prof_forest_$tag.ncu-rep --page raw --csv > $out/prof_forest_$tag.csv 2>/dev/null
ncu -i $out/prof_forest1_$tag.ncu-rep --page source --csv > $out/prof_forest1_${tag}_src.csv 2>/dev/null
rm -f $out/prof_forest_$tag.ncu-rep
find $out -size +20M -delete
du -sh $out; ls -la $out | tail -30
