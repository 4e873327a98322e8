"""torchrun --nproc-per-node N tools/profile_sharded.py : torch-profiler table of a sharded step on rank 0 (dev tool)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import to_device
from tencent_recommendation_2025_b200.sharded import ShardedBaselineEmbedding

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = True
cfg = bench.get_config("c2", 1024); w = synth.SynthWorld(cfg, 0); lay = w.layout
torch.manual_seed(0)
m = ShardedBaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(),
                             types.SimpleNamespace(device=str(dev), hidden_units=cfg.H), rank, world)
with torch.no_grad():
    m.local_table.normal_(0, 0.05)
dense = [p for p in m.parameters() if p is not m.local_table]
opt = torch.optim.AdamW(dense, lr=1e-3, betas=(0.9, 0.98))
st = w.make_step(1000 * rank)
pbs = [to_device(lay, pc, dev) for pc in st.calls]; ups = [torch.from_numpy(r).to(dev) for r in st.upstream]

flat_grad = torch.zeros(sum(p.numel() for p in dense), device=dev)
o = 0
for p in dense:
    p.grad = flat_grad[o:o + p.numel()].view_as(p); o += p.numel()
LOOK = os.environ.get("LOOKAHEAD", "1") == "1"

def step():
    flat_grad.zero_()
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups)
    if LOOK: m.prepare_next(pbs)
    dist.all_reduce(flat_grad); flat_grad.div_(world)
    opt.step()
    m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    if LOOK: m.finish_prepare()

for _ in range(4): step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); t1 = time.perf_counter()
if rank == 0: print(f"wall {1e3*(t1-t0)/10:.3f} ms/step at W={world}")
# phase timing with CUDA events (GPU time between phase boundaries on the compute stream) and host time
names = ["prefetch", "fwd x3", "backward", "dense allreduce+opt", "fused_step"]
acc = {k: [0.0, 0.0] for k in names}
def phase(name, f):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record(); r = f(); e1.record(); acc[name][1] += time.perf_counter() - t
    evs.append((name, e0, e1)); return r
def dense_sync():
    flat = torch.cat([p.grad.reshape(-1) for p in dense]); dist.all_reduce(flat); flat /= world
    o = 0
    for p in dense:
        p.grad.copy_(flat[o:o + p.numel()].view_as(p)); o += p.numel()
    opt.step()
evs = []
N = 10
torch.cuda.synchronize(); dist.barrier()
for _ in range(N):
    opt.zero_grad(set_to_none=True)
    phase("prefetch", lambda: m.prefetch(pbs))
    outs = phase("fwd x3", lambda: [m.feat2emb_packed(pb) for pb in pbs])
    phase("backward", lambda: torch.autograd.backward(outs, ups))
    phase("dense allreduce+opt", dense_sync)
    phase("fused_step", lambda: m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2))
torch.cuda.synchronize()
for name, e0, e1 in evs: acc[name][0] += e0.elapsed_time(e1)
if rank == 0:
    for k in names: print(f"  {k:22s} gpu {acc[k][0]/N:.3f} ms   host {1e3*acc[k][1]/N:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
if rank == 0:
    prof.export_chrome_trace("gpurun_out/trace_sharded_r0.json")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
dist.barrier(); dist.destroy_process_group()
