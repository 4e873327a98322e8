"""python tools/host_profile.py : how much HOST time one step costs to enqueue (dev tool; N=1, C2).

Prints the CPU-side enqueue time per step (no synchronize inside the loop) for the device-resident loop and for the two
e2e feeds, then a cProfile table of the resident-feed loop. If enqueue time >= device time the step is launch-bound."""
import cProfile, os, pstats, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import HostPrefetcher, stage_pinned, to_device
from tencent_recommendation_2025_b200.resident import ResidentFeeder, ResidentItemFeatures

dev = torch.device("cuda", 0)
cfg = bench.get_config("c2", 1024)
w = synth.SynthWorld(cfg, 0); lay = w.layout
m = bench.init_module(cfg, dev, "fused", "factored")
dense_opt = torch.optim.AdamW(m.dense_parameters(), lr=1e-3, betas=(0.9, 0.98), fused=True)
steps_np = [w.make_step(s) for s in range(4)]
dev_steps = [([to_device(lay, pc, dev) for pc in st.calls], [torch.from_numpy(r).to(dev) for r in st.upstream]) for st in steps_np]
hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)

def one_step(pbs, ups):
    dense_opt.zero_grad(set_to_none=True)
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups)
    dense_opt.step()
    m.fused_step(**hyper)
    return outs

def timed(fn, n=40, warm=6):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn(warm + i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return 1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n

print("resident loop: enqueue %.3f ms/step, wall %.3f ms/step" % timed(lambda i: one_step(*dev_steps[i % 4])))

def make_feed_loop(feeder, host_steps):
    loss_host = torch.zeros(2, dtype=torch.float32, pin_memory=True)
    st = {"ev": [None, None]}
    feeder.submit(host_steps[0])
    def fn(i):
        pbs = feeder.take()
        feeder.submit(host_steps[(i + 1) % 4])
        outs = one_step(pbs, dev_steps[i % 4][1])
        feeder.retire()
        with torch.no_grad():
            loss = sum(o.detach().sum() for o in outs)
        slot = i & 1
        if st["ev"][slot] is not None:
            st["ev"][slot].synchronize(); float(loss_host[slot])
        loss_host[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)
        st["ev"][slot] = torch.cuda.Event(); st["ev"][slot].record()
    return fn

host_steps = [[stage_pinned(lay, pc) for pc in st.calls] for st in steps_np]
f1 = make_feed_loop(HostPrefetcher(dev), host_steps)
print("packed feed : enqueue %.3f ms/step, wall %.3f ms/step" % timed(f1))
store = ResidentItemFeatures.from_world(w, dev, torch.float32)
slim_steps = [[store.slim(pc) for pc in st.calls] for st in steps_np]
f2 = make_feed_loop(ResidentFeeder(store), slim_steps)
print("slim feed   : enqueue %.3f ms/step, wall %.3f ms/step" % timed(f2))
for name, fn in (("slim feed", f2), ("resident loop", lambda i: one_step(*dev_steps[i % 4]))):
    pr = cProfile.Profile(); pr.enable()
    for i in range(46, 86): fn(i)
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print("==== cProfile", name, "(40 steps)"); print(s.getvalue()[:9000])
