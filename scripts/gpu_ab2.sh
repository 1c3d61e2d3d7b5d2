#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/bench_variants.py > gpurun_out/variants2.log 2>&1; tail -8 gpurun_out/variants2.log
python scripts/run_edge_ik.py > gpurun_out/plain_edge_ik.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:edge_kernel -s 2 -c 1 -f -o gpurun_out/prof_edge python scripts/run_edge_ik.py > gpurun_out/ncu_edge.log 2>&1
python scripts/run_edge_ik.py > gpurun_out/plain_edge_ik2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ik_kernel -s 2 -c 1 -f -o gpurun_out/prof_ik python scripts/run_edge_ik.py > gpurun_out/ncu_ik.log 2>&1
tail -2 gpurun_out/ncu_ik.log
