#!/bin/bash
# round 2, call D (re-entry): whole GPU suite, default bench line with the per-call table, launch list
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
python -m pytest tests -q -m gpu > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -30 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --profile-calls --trace > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
tail -90 $out/bench_$tag.err
cat $out/bench_$tag.json
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-verify"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 4000 --csv \
    --log-file $out/launches_$tag.csv $B > $out/ncu_launch_$tag.log 2>&1
ls -la $out | tail -8
