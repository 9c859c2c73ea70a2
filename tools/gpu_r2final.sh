#!/bin/bash
# round 2, final call: the whole GPU suite, smoke(), the default bench line, the reference arm, the launch list of the
# step's kernels and ncu --set full captures of the forest kernel (trained forests) and the streaming kernels
tag=${1:-r2final}
out=gpurun_out
mkdir -p $out
python -m pytest tests -q -m gpu --durations=10 > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -18 $out/pytest_$tag.log
python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke exit $?"; tail -2 $out/smoke_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
tail -5 $out/bench_$tag.err; cat $out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref exit $?"
tail -3 $out/bench_ref_$tag.err; cat $out/bench_ref_$tag.json
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-verify --configs none"
STEP='k_forest|k_scan|k_pairs|k_domain|k_dc_|k_gather|k_write_rows|k_block_popc|k_scan_counts|k_popc|k_bitmap|k_lut|k_tile|k_lookup|k_changed|k_key|k_combine|k_index|k_ids|k_valid|k_widen|k_range|k_discretize'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$STEP" -c 3000 --csv \
    --log-file $out/launches_$tag.csv $B > $out/ncu_launch_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_forest_predict_ranked' -s 5 -c 2 -f \
    -o $out/prof_forest_$tag $B > $out/ncu_forest_$tag.log 2>&1
ncu -i $out/prof_forest_$tag.ncu-rep --page raw --csv > $out/prof_forest_$tag.csv 2>/dev/null
ncu -i $out/prof_forest_$tag.ncu-rep --page source --csv > $out/prof_forest_${tag}_src.csv 2>/dev/null
# streaming kernels of the timed step (the set-up pass launches each of them once before): scan, pair counts,
# dirty-row gather, domain pruning
timeout 900 ncu --set full --clock-control none -k regex:'k_scan_hist|k_pairs|k_gather_rows_masked|k_domain_prune_bits|k_dc_fd' \
    -c 40 -f -o $out/prof_stream_$tag $B > $out/ncu_stream_$tag.log 2>&1
ncu -i $out/prof_stream_$tag.ncu-rep --page raw --csv > $out/prof_stream_$tag.csv 2>/dev/null
find $out -size +30M -delete
ls -la $out | tail -12
