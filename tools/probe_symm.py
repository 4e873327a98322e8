"""Probe: torch symmetric memory (CUDA VMM / IPC) between the ranks of one box + P2P read bandwidth (dev tool)."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
n = 64 << 20   # floats: 256 MB
t = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
t.fill_(float(rank + 1))
torch.cuda.synchronize(); dist.barrier()
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (n,), torch.float32)
print(rank, "peer", peer, "first", float(pb[0]), "ptr", hex(pb.data_ptr()), flush=True)
dst = torch.empty(n, device=dev)
for _ in range(2): dst.copy_(pb)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dst.copy_(pb); e1.record(); torch.cuda.synchronize()
print(rank, f"peer copy {n*4/e0.elapsed_time(e1)/1e6:.1f} GB/s", flush=True)
# random 256 B row gather from the peer
idx = torch.randint(0, n // 64, (227000,), device=dev)
rows = pb.view(-1, 64)
for _ in range(2): out = rows[idx]
torch.cuda.synchronize(); dist.barrier()
e0.record(); out = rows[idx]; e1.record(); torch.cuda.synchronize()
print(rank, f"peer row gather 227k x 256 B: {e0.elapsed_time(e1)*1e3:.1f} us = {227000*256/e0.elapsed_time(e1)/1e6:.1f} GB/s", flush=True)
dist.barrier(); dist.destroy_process_group()
