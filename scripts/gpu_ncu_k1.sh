#!/bin/bash
# ncu evidence for K1 on the current build: launch list of a short bench run, then one --set full capture with source
# correlation.  Each ncu pass only after the same command exited 0 without it.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --eager --settle-s 0.05"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_rne.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rne_batch_kernel -s 30 -c 2 -f -o gpurun_out/prof_rne $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_rne.ncu-rep
