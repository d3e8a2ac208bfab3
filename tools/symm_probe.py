"""Probe: can two ranks on one node map each other's device buffers (torch symmetric memory) and copy over NVLink?
   torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm

    n = 1 << 28  # 256 Mi complex128 would be 4 GiB; use float32: 1 GiB
    buf = symm.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "multicast", getattr(hdl, "multicast_ptr", None), flush=True)
    buf.fill_(float(rank + 1))
    torch.cuda.synchronize(); dist.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    mine = torch.empty(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mine.copy_(peer)
    e0.record()
    for _ in range(5):
        mine.copy_(peer)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, "peer read GB/s", 4 * n / ms / 1e6, "value", float(mine[12345]), flush=True)
    dist.barrier()
    e0.record()
    for _ in range(5):
        peer.copy_(mine)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, "peer write GB/s", 4 * n / ms / 1e6, flush=True)
    # big allocation probe: 2 x 8 GiB
    big = symm.empty(1 << 31, dtype=torch.float32, device=torch.device("cuda", local))
    symm.rendezvous(big, dist.group.WORLD.group_name)
    print(rank, "8 GiB symmetric allocation ok", flush=True)
except Exception as exc:  # noqa: BLE001
    import traceback

    traceback.print_exc()
    print(rank, "symmetric memory unavailable:", repr(exc)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
