#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N): the sharded parity test, then bench.py under torchrun for each exchange.
# usage: scripts/gpu_multi.sh <tag> <N> [configs] [exchanges]
tag=$1; N=$2; configs=${3:-c2}; exchanges=${4:-"peer nccl"}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$tag.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rf -s > gpurun_out/tests_multi_$tag.log 2>&1; echo "multi-gpu tests rc=$?"
tail -15 gpurun_out/tests_multi_$tag.log
for c in $configs; do
  for ex in $exchanges; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 30 --warmup 5 --config $c --exchange $ex > gpurun_out/bench_${tag}_${c}_${ex}.json 2> gpurun_out/bench_${tag}_${c}_${ex}.err
    echo "bench $c $ex rc=$?"; tail -c 600 gpurun_out/bench_${tag}_${c}_${ex}.err; cat gpurun_out/bench_${tag}_${c}_${ex}.json | cut -c1-1800
  done
done
