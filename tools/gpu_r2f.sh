#!/bin/bash
# round 2, call F: tests after the trainer binning / batched compaction / goldens budget changes, default bench line,
# a 12.5M-row pass (the per-GPU share of the 8-GPU strong-scaling run) with its phase trace
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_goldens.py tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_arrow.py \
    tests/test_gpu_forest_regime.py -q -m gpu --durations=15 > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -40 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --profile-calls > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
tail -30 $out/bench_$tag.err
cat $out/bench_$tag.json
python bench.py --rows 12500000 --steps 5 --warmup 3 --no-cpu-baseline --configs none --profile-calls --trace \
    > $out/bench_12M_$tag.json 2> $out/bench_12M_$tag.err; echo "bench 12.5M exit $?"
grep -v "^trace chain" $out/bench_12M_$tag.err | tail -45
cat $out/bench_12M_$tag.json
ls -la $out | tail -6
