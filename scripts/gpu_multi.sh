#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N): the sharded parity test, then bench.py under torchrun.
# usage: scripts/gpu_multi.sh <tag> <N> [configs] [variants: "peer:pixel peer:time nccl:time"] [tests|notests]
tag=$1; N=$2; configs=${3:-c2}; variants=${4:-"peer:pixel peer:time"}; tests=${5:-tests}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$tag.txt 2>&1
if [ "$tests" = tests ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rf -s > gpurun_out/tests_multi_$tag.log 2>&1; echo "multi-gpu tests rc=$?"
  tail -4 gpurun_out/tests_multi_$tag.log
fi
for c in $configs; do
  for v in $variants; do
    ex=${v%%:*}; sh=${v##*:}
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 30 --warmup 5 --config $c --exchange $ex --shard $sh > gpurun_out/bench_${tag}_${c}_${ex}_${sh}.json 2> gpurun_out/bench_${tag}_${c}_${ex}_${sh}.err
    echo "bench $c $ex $sh rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/bench_${tag}_${c}_${ex}_${sh}.err | tail -c 800
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_${c}_${ex}_${sh}.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "ms_per_step", "value", "plan_ms", "reshard_ms", "events_this_rank", "sharded_vs_single")}, d["e2e"]["value"])
except Exception as e:
    print("no bench line:", e)
PY
  done
done
