"""torchrun --nproc-per-node N tools/profile_host_sharded.py : host-side cost of one sharded step on rank 0 (dev tool)."""
import cProfile, os, pstats, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import to_device
from tencent_recommendation_2025_b200.sharded import ShardedBaselineEmbedding
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.get_config("c2", 1024); w = synth.SynthWorld(cfg, 0); lay = w.layout
torch.manual_seed(0)
m = ShardedBaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(),
                             types.SimpleNamespace(device=str(dev), hidden_units=cfg.H), rank, world, path="factored")
with torch.no_grad():
    m.local_table.normal_(0, 0.05)
dense = [p for p in m.parameters() if p is not m.local_table]
opt = torch.optim.AdamW(dense, lr=1e-3, betas=(0.9, 0.98), fused=True)
sts = [w.make_step(1000 * rank + i) for i in range(2)]
B = [([to_device(lay, pc, dev) for pc in st.calls], [torch.from_numpy(r).to(dev) for r in st.upstream]) for st in sts]
flat_grad = torch.zeros(sum(p.numel() for p in dense), device=dev)
o = 0
for p in dense:
    p.grad = flat_grad[o:o + p.numel()].view_as(p); o += p.numel()
def step(i):
    pbs, ups = B[i % 2]
    flat_grad.zero_()
    m.prefetch(pbs)
    m.prepare_next(B[(i + 1) % 2][0])
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups)
    dist.all_reduce(flat_grad); flat_grad.div_(world)
    opt.step()
    m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    m.finish_prepare()
for i in range(6): step(i)
torch.cuda.synchronize(); dist.barrier()
N = 20
t = 0.0
for i in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(i); t += time.perf_counter() - t0
if rank == 0: print(f"host enqueue per step (queue empty at step start) {1e3*t/N:.3f} ms at W={world}, cores {os.cpu_count()}")
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for i in range(N): step(i)
torch.cuda.synchronize(); t1 = time.perf_counter()
if rank == 0: print(f"free-running {1e3*(t1-t0)/N:.3f} ms/step")
pr = cProfile.Profile()
for i in range(N):
    torch.cuda.synchronize(); pr.enable(); step(i); pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
dist.barrier(); dist.destroy_process_group()
