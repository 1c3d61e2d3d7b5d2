#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
