#!/bin/bash
# The driver's scaling run on one 8-GPU box: N = 2, 4, 8 back to back (N = 1 comes from scripts/gpu_final.sh on a
# 1-GPU box -- an 8-GPU box is charged 8x).  SCALE_N1=1 adds the N = 1 line here too.
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ "${SCALE_N1:-0}" = "1" ]; then
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_n1.log 2>&1; tail -1 gpurun_out/scale_n1.log | cut -c1-400
fi
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$N.log 2>&1; echo "exit $?" >> gpurun_out/scale_n$N.log
grep '^{' gpurun_out/scale_n$N.log | cut -c1-260
tail -3 gpurun_out/scale_n$N.log | cut -c1-300
done
