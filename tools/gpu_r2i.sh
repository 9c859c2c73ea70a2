#!/bin/bash
# round 2, call I (2 GPUs): the 2-rank device parity test over NCCL, the strong-scaling bench line at N=2,
# the reference arm under torchrun (all host cores)
tag=${1:-r2i}
out=gpurun_out
mkdir -p $out
nvidia-smi -L
python -m pytest tests/test_gpu_dist.py -q -m gpu > $out/pytest_dist_$tag.log 2>&1; echo "pytest exit $?" >> $out/pytest_dist_$tag.log
tail -5 $out/pytest_dist_$tag.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 5 --warmup 3 --configs none --profile-calls > $out/bench_2gpu_$tag.json 2> $out/bench_2gpu_$tag.err; echo "bench 2gpu exit $?"
tail -25 $out/bench_2gpu_$tag.err | cut -c1-300
cat $out/bench_2gpu_$tag.json
$T bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-rows 20000 > $out/bench_ref_2gpu_$tag.json 2> $out/bench_ref_2gpu_$tag.err; echo "ref 2gpu exit $?"
tail -3 $out/bench_ref_2gpu_$tag.err | cut -c1-300; cat $out/bench_ref_2gpu_$tag.json | cut -c1-1200
