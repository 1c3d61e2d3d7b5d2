#!/bin/bash
set -x
mkdir -p gpurun_out
python bench_sharded.py > gpurun_out/sharded_n1.log 2>&1; grep '^{' gpurun_out/sharded_n1.log
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench_sharded.py > gpurun_out/sharded_n$N.log 2>&1; echo "exit $?" >> gpurun_out/sharded_n$N.log
grep -E '^\{|exit|Error|error' gpurun_out/sharded_n$N.log | cut -c1-300
done
python -m pytest tests/test_gpu_guards.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -4
