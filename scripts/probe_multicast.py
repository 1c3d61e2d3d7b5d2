#!/usr/bin/env python
"""torchrun probe: does this box offer NVSwitch multicast (NVLS) through torch's symmetric memory?"""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
    h = symm.rendezvous(t, group=dist.group.WORLD.group_name)
    if rank == 0:
        print("symm ok: world", h.world_size, "multicast_ptr", hex(h.multicast_ptr), "buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:4],
              "signal_pad_ptrs", len(h.signal_pad_ptrs))
except Exception as e:
    if rank == 0:
        print("symm failed:", repr(e)[:500])
import ctypes
try:
    cu = ctypes.CDLL("libcuda.so.1")
    v = ctypes.c_int(0)
    cu.cuDeviceGetAttribute(ctypes.byref(v), 132, local)   # CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED
    if rank == 0:
        print("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", v.value)
except Exception as e:
    print("cuda attr failed", e)
dist.barrier()
dist.destroy_process_group()
