#!/bin/bash
# round 2, last call: the whole GPU suite on the final tree, smoke(), the default bench line, the reference arm
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
python -m pytest tests -q -m gpu --durations=5 > $out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_$tag.log
tail -12 $out/pytest_$tag.log
python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke exit $?"; tail -2 $out/smoke_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?"
tail -5 $out/bench_$tag.err; cat $out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref exit $?"
tail -3 $out/bench_ref_$tag.err; cat $out/bench_ref_$tag.json
