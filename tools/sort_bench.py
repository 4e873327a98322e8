"""Micro-benchmark of tgr_sort_pairs (dev tool): python tools/sort_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tencent_recommendation_2025_b200 import _lib
lib = _lib.load()
_lib.timing_enable(False)
for n, bits in ((2_850_000, 24), (227_000, 21), (227_000, 24)):
    g = torch.Generator(device="cuda").manual_seed(0)
    keys = (torch.rand(n, device="cuda", generator=g, dtype=torch.float64) ** 3 * ((1 << bits) - 1)).to(torch.int64).to(torch.int32)
    vals = torch.arange(n, device="cuda", dtype=torch.int32)
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    ws = torch.empty(lib.tgr_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def run():
        _lib.check(lib.tgr_sort_pairs(keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), n, bits, ws.data_ptr(), ws.numel(), st))
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"sort n={n} bits={bits}: median {1e3*ts[5]:.1f} us")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(); torch.cuda.synchronize()
    for e in prof.key_averages():
        print(f"    {e.key[:60]:60s} {e.device_time_total:8.1f} us x{e.count}")
