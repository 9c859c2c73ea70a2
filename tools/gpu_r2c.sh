#!/bin/bash
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
python -m pytest tests -q -m gpu --deselect tests/test_gpu_goldens.py > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -30 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-calls --trace > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
tail -70 $out/bench_$tag.err
