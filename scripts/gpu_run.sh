#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, one bench line, launch list, optional full ncu capture.
# usage: scripts/gpu_run.sh <tag> [tests|notests] [ncu|noncu] [extra bench flags]
tag=$1; tests=${2:-tests}; ncu=${3:-noncu}; shift 3
mkdir -p gpurun_out
if [ "$tests" = tests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests_$tag.log
  tail -5 gpurun_out/tests_$tag.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_$tag.err
cat gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 --skip-cpu "$@" > gpurun_out/bench_under_ncu_$tag.log 2>&1
if [ "$ncu" = ncu ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vote_|grad_|image_kernel' -s 9 -c 3 -f -o gpurun_out/prof_$tag \
      python bench.py --steps 5 --warmup 3 --skip-cpu --no-graph "$@" > gpurun_out/ncu_full_$tag.log 2>&1
fi
ls -la gpurun_out | tail -20
