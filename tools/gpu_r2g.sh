#!/bin/bash
# round 2, call G: balanced forest partition + pipelined ingest: tests, short bench at 100M and 12.5M rows
tag=${1:-r2g}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_forest_regime.py tests/test_gpu_kernels.py tests/test_gpu_arrow.py tests/test_gpu_dist.py \
    "tests/test_gpu_pipeline.py::test_synthetic_parity" -q -m gpu -x > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -12 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs none --profile-calls > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
head -4 $out/bench_$tag.err
python -c "
import json
d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','ms_detect_phase','ms_repair_phase','verify')}); print(d['e2e']); print(d['kernels'])
"
python bench.py --rows 12500000 --steps 5 --warmup 3 --no-cpu-baseline --configs none --profile-calls --trace \
    > $out/bench_12M_$tag.json 2> $out/bench_12M_$tag.err; echo "bench 12.5M exit $?"
grep -v "^trace chain" $out/bench_12M_$tag.err | head -40
cat $out/bench_12M_$tag.json | cut -c1-1500
