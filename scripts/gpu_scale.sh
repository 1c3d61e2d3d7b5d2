#!/bin/bash
# 1 -> N weak-scaling run of bench.py on one box: gpurun --gpus 8 -- bash scripts/gpu_scale.sh [1 2 4 8]
mkdir -p gpurun_out
NS=${@:-1 2 4 8}
for n in $NS; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "== N=$n rc=$?"; tail -2 gpurun_out/scale_n$n.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
print(d["n_gpus"], "value %.3e ms/step %.5f sustained %.3e kernel_ms %.5f e2e %.3e" % (d["value"], d["ms_per_step"], d["sustained"]["value"], d["roofline"]["kernel_ms"], d["e2e"]["value"]), d["launch"], d.get("gather_check"))
print("   link", {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["e2e"]["link"].items() if k!="note"})
print("   clocks", d["clocks"])
PY
done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu > gpurun_out/lscpu.txt 2>&1; numactl -H >> gpurun_out/lscpu.txt 2>&1; ls /sys/devices/system/node/ >> gpurun_out/lscpu.txt
