#!/bin/bash
# round 2, call B: forest kernel v2 (rank slots, per-node leaf values): parity tests + bench per layout
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_forest_regime.py tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -x -m gpu > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -5 $out/pytest_$tag.log
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
for lay in "" wide8 wide16 bytes; do
  DR_RANKED_LAYOUT=$lay $B > $out/bench_${tag}_${lay:-auto}.json 2> $out/bench_${tag}_${lay:-auto}.err; echo "bench $lay exit $?"
done
ls -la $out | tail -6
