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
                             types.SimpleNamespace(device=str(dev), hidden_units=cfg.H), rank, world,
                             path=os.environ.get("TGR_PATH", "factored"))
with torch.no_grad():
    m.local_table.normal_(0, 0.05)
dense = [p for p in m.parameters() if p is not m.local_table]
opt = None
NB = int(os.environ.get("BATCHES", "1"))          # distinct batches cycled (bench.py cycles 4)
sts = [w.make_step(1000 * rank + s) for s in range(NB)]
all_pbs = [[to_device(lay, pc, dev) for pc in st.calls] for st in sts]
all_ups = [[torch.from_numpy(r).to(dev) for r in st.upstream] for st in sts]
pbs, ups = all_pbs[0], all_ups[0]
step_i = [0]

flat_grad = m.symm_empty(sum(p.numel() for p in dense))
o = 0
for p in dense:
    p.grad = flat_grad[o:o + p.numel()].view_as(p); o += p.numel()
LOOK = os.environ.get("LOOKAHEAD", "1") == "1"

def step():
    k = step_i[0] % NB; step_i[0] += 1
    pbs, ups, nxt = all_pbs[k], all_ups[k], all_pbs[(k + 1) % NB]
    flat_grad.zero_()
    m.prefetch(pbs)
    if LOOK: m.prepare_next(nxt)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups)
    m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    m.allreduce_dense_(flat_grad)
    m.dense_adam_(dense, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    if LOOK: m.finish_prepare()

for _ in range(4): step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize(); t1 = time.perf_counter()
if rank == 0: print(f"wall {1e3*(t1-t0)/10:.3f} ms/step at W={world}")
if os.environ.get("CPROFILE") == "1":
    import cProfile, pstats, io
    def timed(n):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n): step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        return 1e3 * (t1 - t0) / n, 1e3 * (time.perf_counter() - t0) / n
    e, w = timed(20)
    if rank == 0: print(f"enqueue {e:.3f} ms/step, wall {w:.3f} ms/step")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): step()
    pr.disable(); torch.cuda.synchronize()
    if rank == 0:
        for key in ("cumulative", "tottime"):
            sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats(key).print_stats(45)
            print("==== cProfile by", key); print(sio.getvalue()[:7000])
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    prof.export_chrome_trace("gpurun_out/trace_sharded_r0.json")
    import json
    ev = [e for e in json.load(open("gpurun_out/trace_sharded_r0.json"))["traceEvents"]
          if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ev.sort(key=lambda e: e["ts"])
    t_first = ev[0]["ts"]
    span = ev[-1]["ts"] + ev[-1]["dur"] - t_first
    print(f"GPU timeline of 3 steps on rank 0: {span/3:.0f} us/step, {len(ev)//3} device ops/step")
    lo = t_first + 2 * span / 3          # last step
    last_end = None
    for e in ev:
        if e["ts"] < lo: continue
        gap = 0 if last_end is None else e["ts"] - last_end
        nm = e["name"][:60]
        print(f"  +{e['ts']-lo:8.1f} us  dur {e['dur']:7.1f}  gap {gap:7.1f}  s{e['args'].get('stream','?')}  {nm}")
        last_end = max(last_end or 0, e["ts"] + e["dur"])
dist.barrier(); dist.destroy_process_group()
