#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/bench_variants.py > gpurun_out/variants.log 2>&1; tail -8 gpurun_out/variants.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
