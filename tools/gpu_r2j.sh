#!/bin/bash
# round 2, call J: forest launch grid in units of warps (small cell lists spread over all SMs): tests + bench at 100M / 12.5M
tag=${1:-r2j}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_forest_regime.py tests/test_gpu_kernels.py tests/test_gpu_arrow.py "tests/test_gpu_pipeline.py::test_synthetic_parity" \
    "tests/test_gpu_pipeline.py::test_adult_pmf_modes_parity" "tests/test_gpu_pipeline.py::test_hospital_repair_parity" -q -m gpu -x > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -5 $out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --configs none --profile-calls --trace > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
grep "chain:predict c0[0-7]" $out/bench_$tag.err
python -c "
import json
d=json.loads(open('$out/bench_$tag.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','ms_detect_phase','ms_repair_phase','verify')}); print(d['e2e']); print(d['kernels']['forest_predict_ranked'])
"
python bench.py --rows 12500000 --steps 5 --warmup 3 --no-cpu-baseline --configs none --no-e2e --profile-calls --trace \
    > $out/bench_12M_$tag.json 2> $out/bench_12M_$tag.err; echo "bench 12.5M exit $?"
grep "chain:predict c0[0-7]" $out/bench_12M_$tag.err
python -c "
import json
d=json.loads(open('$out/bench_12M_$tag.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','ms_detect_phase','ms_repair_phase','verify')}); print(d['kernels']['forest_predict_ranked'])
"
