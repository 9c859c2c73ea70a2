#!/bin/bash
# round 2, call H: leaf planes split into two 32-bit loads: forest tests + short bench
tag=${1:-r2h}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_forest_regime.py tests/test_gpu_kernels.py "tests/test_gpu_pipeline.py::test_synthetic_parity" \
    "tests/test_gpu_pipeline.py::test_adult_pmf_modes_parity" -q -m gpu -x > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -6 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs none --no-e2e --profile-calls > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
head -3 $out/bench_$tag.err
python -c "
import json
d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','ms_detect_phase','ms_repair_phase','verify','roofline_forest')})
"
