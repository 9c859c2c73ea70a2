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
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_scan_hist|k_gather_rows_masked|k_pairs|k_tile_null_bitmaps' -s 3 -c 6 -f -o $out/prof_stream_$tag \
    $B > $out/ncu_stream_$tag.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:'k_forest_predict_ranked' -s 32 -c 32 -f \
    -o $out/prof_forest_$tag $B > $out/ncu_forest_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_forest_predict_ranked' -s 40 -c 1 -f \
    -o $out/prof_forest1_$tag $B > $out/ncu_forest1_$tag.log 2>&1
for r in stream forest; do
  ncu -i $out/prof_${r}_$tag.ncu-rep --page raw --csv > $out/prof_${r}_$tag.csv 2>/dev/null
done
rm -f $out/prof_forest_$tag.ncu-rep
find $out -size +20M -delete
du -sh $out; ls -la $out | tail -20
