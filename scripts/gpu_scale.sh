#!/bin/bash
# The driver's scaling run: N = 1, 2, 4, 8 back to back on one box.
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_n1.log 2>&1; tail -1 gpurun_out/scale_n1.log | cut -c1-400
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$N.log 2>&1; echo "exit $?" >> gpurun_out/scale_n$N.log
grep '^{' gpurun_out/scale_n$N.log | cut -c1-260
tail -3 gpurun_out/scale_n$N.log | cut -c1-300
done
