#!/bin/bash
# ncu evidence for (a) the fused compute + gather kernel with NVLink counters (2 GPUs, one process) and (b) the IK kernel.
mkdir -p gpurun_out
python scripts/ncu_scatter_single_process.py > gpurun_out/scatter_plain.log 2>&1 || { tail -5 gpurun_out/scatter_plain.log; exit 1; }
M=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_op_write.sum,dram__bytes_write.sum,dram__bytes_read.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:rne_batch_kernel --csv --log-file gpurun_out/ncu_scatter_n2.csv python scripts/ncu_scatter_single_process.py > gpurun_out/ncu_scatter.log 2>&1
tail -3 gpurun_out/ncu_scatter.log
python scripts/time_ik_once.py > gpurun_out/ik_plain.log 2>&1 || { tail -5 gpurun_out/ik_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ik_kernel_compact -s 2 -c 2 -f -o gpurun_out/prof_ik python scripts/time_ik_once.py > gpurun_out/ncu_ik.log 2>&1
tail -2 gpurun_out/ncu_ik.log
