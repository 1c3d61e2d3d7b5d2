#!/bin/bash
# Runs on the GPU box via gpurun: parity tests, smoke, bench, ncu launch list + one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/bench_ref.log
if [ "${NCU:-1}" = "1" ]; then
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rne_batch_kernel -s 3 -c 2 -f -o gpurun_out/prof_rne \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
