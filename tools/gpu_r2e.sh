#!/bin/bash
# round 2, call E: Arrow ingest / domain-prune tests, full default bench line (API e2e, other configs, cpu baseline),
# reference arm, launch list of the step kernels, ncu --set full of the forest kernel (trained forests, with source)
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_arrow.py tests/test_gpu_pipeline.py tests/test_gpu_goldens.py -q -m gpu --durations=25 \
    > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -45 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --profile-calls --trace > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
grep -v "^trace chain" $out/bench_$tag.err | tail -60
cat $out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref exit $?"
tail -5 $out/bench_ref_$tag.err; cat $out/bench_ref_$tag.json
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-verify --configs none"
STEP='k_forest|k_scan|k_pairs|k_domain|k_dc_|k_gather|k_write_rows|k_block_popc|k_scan_counts|k_popc|k_bitmap|k_lut|k_tile|k_lookup|k_changed|k_key|k_combine|k_index|k_ids|k_valid|k_widen|k_range|k_discretize'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$STEP" -c 3000 --csv \
    --log-file $out/launches_$tag.csv $B > $out/ncu_launch_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_forest_predict_ranked' -s 5 -c 2 -f \
    -o $out/prof_forest_$tag $B > $out/ncu_forest_$tag.log 2>&1
ncu -i $out/prof_forest_$tag.ncu-rep --page raw --csv > $out/prof_forest_$tag.csv 2>/dev/null
ncu -i $out/prof_forest_$tag.ncu-rep --page source --csv > $out/prof_forest_${tag}_src.csv 2>/dev/null
find $out -size +30M -delete
ls -la $out | tail -12
