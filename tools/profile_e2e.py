"""Host-side breakdown of the e2e loop (dev tool): python tools/profile_e2e.py [--path factored]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import stage_pinned

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=8); ap.add_argument("--path", default="factored")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
cfg = bench.get_config("c2", 1024)
w = synth.SynthWorld(cfg, 0); lay = w.layout
m = bench.init_module(cfg, dev, "fused", a.path)
opt = torch.optim.AdamW(m.dense_parameters(), lr=1e-3, betas=(0.9, 0.98), fused=True)
sts = [w.make_step(i) for i in range(4)]
host = [[stage_pinned(lay, pc) for pc in st.calls] for st in sts]
ups = [[torch.from_numpy(r).to(dev) for r in st.upstream] for st in sts]
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for i in range(a.steps):
    k = i % 4
    t0 = sync()
    pbs = [hp.upload(dev) for hp in host[k]]
    t1 = sync()
    opt.zero_grad(set_to_none=True)
    m.prefetch(pbs)
    t2 = sync()
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    t3 = sync()
    torch.autograd.backward(outs, ups[k])
    t4 = sync()
    opt.step(); m.fused_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    t5 = sync()
    loss = float(sum(o.sum() for o in outs).item())
    t6 = sync()
    ms = torch.cuda.memory_stats()
    print(f"step {i}: h2d {1e3*(t1-t0):.2f} prefetch {1e3*(t2-t1):.2f} fwd {1e3*(t3-t2):.2f} bwd {1e3*(t4-t3):.2f} "
          f"opt {1e3*(t5-t4):.2f} loss {1e3*(t6-t5):.2f} | total {1e3*(t6-t0):.2f} ms | device_alloc {ms['num_device_alloc']} "
          f"free {ms['num_device_free']} reserved {ms['reserved_bytes.all.current']/2**30:.1f} GiB")
