#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, bench lines, launch list, optional full ncu capture.
# usage: scripts/gpu_run.sh <tag> [tests|notests] [ncu|noncu] [configs, e.g. "c2 c3"] [extra bench flags]
tag=$1; tests=${2:-tests}; ncu=${3:-noncu}; configs=${4:-c2}; shift 4
mkdir -p gpurun_out
if [ "$tests" = tests ]; then
  timeout 2400 python -m pytest tests -m gpu -q -rf -s > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests_$tag.log
  grep -E "^FAILED|passed|failed|config [0-9]" gpurun_out/tests_$tag.log | tail -40
fi
for c in $configs; do
  timeout 900 python bench.py --steps 20 --warmup 5 --config $c "$@" > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err; echo "bench $c rc=$?"
  tail -c 1200 gpurun_out/bench_${tag}_$c.err
  cat gpurun_out/bench_${tag}_$c.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_${tag}_$c.csv \
      python bench.py --steps 5 --warmup 3 --skip-cpu --config $c "$@" > gpurun_out/bench_under_ncu_${tag}_$c.log 2>&1
  if [ "$ncu" = ncu ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vote_|grad_|image_kernel|flow_voxel|tile_' -s 12 -c 8 -f -o gpurun_out/prof_${tag}_$c \
        python bench.py --steps 5 --warmup 3 --skip-cpu --no-graph --config $c "$@" > gpurun_out/ncu_full_${tag}_$c.log 2>&1
  fi
done
ls -la gpurun_out | tail -12
