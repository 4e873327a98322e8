"""python tools/graph_timeline.py : device timeline of one replay of the graphed C2 step (torch profiler; dev tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.graphed import GraphedStep, PipelinedStep
from tencent_recommendation_2025_b200.resident import CallShape, ResidentItemFeatures

dev = torch.device("cuda", 0)
cfg = bench.get_config(os.environ.get("CFG", "c2"), 1024)
w = synth.SynthWorld(cfg, 0)
m = bench.init_module(cfg, dev, "fused", "factored")
steps = [w.make_step(s) for s in range(4)]
store = ResidentItemFeatures.from_world(w, dev)
shapes = [CallShape.covering([st.calls[i] for st in steps]) for i in range(3)]
fixed = [store.slim_step(st.calls, shapes) for st in steps]
dev_fixed = [f.ints.to(dev) for f in fixed]
ups = [torch.from_numpy(r).to(dev) for r in steps[0].upstream]
m.own_dense_parameters()
hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)

def body(pbs):
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups)
    m.fused_step(**hyper, dense=True)
    with torch.no_grad():
        return torch.stack([o.detach()[0, -1] for o in outs]).sum()

PIPE = os.environ.get("PIPE", "1") == "1"
r = (PipelinedStep if PIPE else GraphedStep)(m, store, fixed[0], body, hyper=hyper)
if PIPE:
    r.prime(dev_fixed[0])
for i in range(8):
    r.load(dev_fixed[(i + 1) % 4]); r.run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4):
        r.load(dev_fixed[(i + 1) % 4]); r.run()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/trace_graph.json")
ev = [e for e in json.load(open("gpurun_out/trace_graph.json"))["traceEvents"]
      if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print(f"{len(ev)//4} device ops/step, {span/4:.1f} us/step over 4 replays")
lo = ev[0]["ts"] + 3 * span / 4
last = None; busy = 0.0; gaps = 0.0
for e in ev:
    if e["ts"] < lo: continue
    gap = 0 if last is None else e["ts"] - last
    print(f"  +{e['ts']-lo:8.1f} us  dur {e['dur']:7.1f}  gap {gap:6.1f}  s{e['args'].get('stream','?')}  {e['name'][:70]}")
    busy += e["dur"]; gaps += max(gap, 0)
    last = max(last or 0, e["ts"] + e["dur"])
print(f"last replay: busy {busy:.1f} us, gaps {gaps:.1f} us")
