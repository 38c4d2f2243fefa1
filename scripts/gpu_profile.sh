#!/bin/bash
# Runs on the GPU box (via gpurun): launch list of one bench run + full ncu capture of the two event kernels.
# usage: scripts/gpu_profile.sh <tag> [extra bench flags]
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 --skip-cpu "$@" > gpurun_out/bench_under_ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'vote_|grad_' -s 6 -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 5 --warmup 3 --skip-cpu --no-graph "$@" > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out
