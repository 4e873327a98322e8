"""cProfile of the host side of one bench step (dev tool): python tools/profile_host.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import to_device
dev = torch.device("cuda", 0)
cfg = bench.get_config("c2", 1024)
w = synth.SynthWorld(cfg, 0); lay = w.layout
m = bench.init_module(cfg, dev, "fused", "factored")
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
for _ in range(5): step()
torch.cuda.synchronize()
N = 30
# unthrottled host time: sync after every step so the launch queue never back-pressures
t = 0.0
for _ in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t += time.perf_counter() - t0
print(f"host enqueue (empty queue) {1e3*t/N:.3f} ms/step")
pr = cProfile.Profile()
for _ in range(N):
    torch.cuda.synchronize(); pr.enable(); step(); pr.disable()
ps = pstats.Stats(pr); ps.sort_stats("cumulative").print_stats(35)
