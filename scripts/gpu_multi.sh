#!/bin/bash
# N-GPU bench through torchrun exactly as the driver launches it.
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; echo "exit $?" >> gpurun_out/bench_n$N.log
tail -4 gpurun_out/bench_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; tail -2 gpurun_out/bench_ref_n$N.log
