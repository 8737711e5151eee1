"""Probe: torch symmetric memory across ranks (peer buffers through NVLink), copy-engine pull bandwidth, device barrier."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
n = 1 << 28   # 1 GiB of int32
t = symm.empty(n, dtype=torch.int32, device=dev)
t.fill_(rank + 1)
hdl = symm.rendezvous(t, group=dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, flush=True)
peer = (rank + 1) % world
pbuf = hdl.get_buffer(peer, (n,), torch.int32)
dst = torch.empty(n, dtype=torch.int32, device=dev)
hdl.barrier()
torch.cuda.synchronize()
for chunk in (1 << 23, 1 << 25, n):          # 32 MiB, 128 MiB, 1 GiB pieces
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for o in range(0, n, chunk):
        dst[o:o + chunk].copy_(pbuf[o:o + chunk], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(rank, f"pull 1 GiB in pieces of {chunk * 4 >> 20} MiB: {ms:.3f} ms = {4 * n / ms / 1e6:.1f} GB/s, ok={bool((dst == peer + 1).all())}", flush=True)
hdl.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    hdl.barrier()
b.record()
torch.cuda.synchronize()
print(rank, f"device barrier: {a.elapsed_time(b) / 20 * 1e3:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
