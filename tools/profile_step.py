"""Host/GPU time breakdown of one bench step (dev tool): python tools/profile_step.py [--steps 5]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import to_device

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--config", default="c2"); ap.add_argument("--path", default="factored")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
cfg = bench.get_config(a.config, 1024)
w = synth.SynthWorld(cfg, 0); lay = w.layout
m = bench.init_module(cfg, dev, "fused", a.path)
opt = torch.optim.AdamW(m.dense_parameters(), lr=1e-3, betas=(0.9, 0.98), fused=True)
st = w.make_step(0)
pbs = [to_device(lay, pc, dev) for pc in st.calls]; ups = [torch.from_numpy(r).to(dev) for r in st.upstream]

def step():
    opt.zero_grad(set_to_none=True)
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups)
    opt.step()
    m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/a.steps:.3f} ms/step, wall {1e3*(t2-t0)/a.steps:.3f} ms/step")
# phase timing (host, unsynced)
import collections
ph = collections.OrderedDict()
def T(name, f):
    t = time.perf_counter(); r = f(); ph[name] = ph.get(name, 0) + time.perf_counter() - t; return r
for _ in range(a.steps):
    T("zero_grad", lambda: opt.zero_grad(set_to_none=True))
    T("prefetch", lambda: m.prefetch(pbs))
    outs = T("fwd x3", lambda: [m.feat2emb_packed(pb) for pb in pbs])
    T("backward", lambda: torch.autograd.backward(outs, ups))
    T("dense opt", lambda: opt.step())
    T("fused_step", lambda: m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2))
torch.cuda.synchronize()
for k, v in ph.items(): print(f"  host {k:12s} {1e3*v/a.steps:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
