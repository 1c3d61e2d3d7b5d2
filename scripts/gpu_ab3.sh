#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/bench_variants.py > gpurun_out/variants3.log 2>&1; tail -6 gpurun_out/variants3.log
N=16000000 python scripts/bench_variants.py > gpurun_out/variants3_16m.log 2>&1; tail -6 gpurun_out/variants3_16m.log
