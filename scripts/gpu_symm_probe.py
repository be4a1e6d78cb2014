"""Probe: torch symmetric memory (CUDA VMM peer mappings over NVLink) between the ranks of one box.
torchrun --nproc-per-node 2 scripts/gpu_symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64 * 1024 * 1024
t = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", hdl.world_size, hdl.rank, [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr), flush=True)
t.fill_(float(rank + 1))
hdl.barrier()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (n,), torch.float32)
print(rank, "peer value", float(pt[0]), float(pt[-1]), flush=True)
# P2P write bandwidth: copy a local buffer into the peer's buffer
src = torch.full((n,), 7.0, device=dev)
torch.cuda.synchronize()
hdl.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pt.copy_(src)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(rank, f"p2p write {n * 4 / ms / 1e6:.1f} GB/s", flush=True)
hdl.barrier()
torch.cuda.synchronize()
print(rank, "own buffer after peer write", float(t[0]), flush=True)
t0 = time.time()
for _ in range(20):
    hdl.barrier()
torch.cuda.synchronize()
print(rank, f"barrier {(time.time() - t0) / 20 * 1e6:.1f} us", flush=True)
dist.destroy_process_group()
