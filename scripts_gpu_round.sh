#!/bin/bash
# usage: scripts_gpu_round.sh TAG  -- bench + launch list + full ncu capture of the top kernel
TAG=$1
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-collisions --no-c4 --no-mass-matrix > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_advance_cc1 -s 4 -c 2 -f -o gpurun_out/prof_$TAG \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-collisions --no-c4 --no-mass-matrix > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail
