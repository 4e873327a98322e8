"""Micro-benchmark of tgr_fact_project_rows / unique_backward on synthetic sorted keys (dev tool)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tencent_recommendation_2025_b200 import _lib
lib = _lib.load()
H, rows, U = 64, 5_000_001, int(sys.argv[1]) if len(sys.argv) > 1 else 227_000
dev = "cuda"
torch.manual_seed(0)
tab = torch.randn(rows, H, device=dev)
W = torch.randn(H, 1024, device=dev)
uniq = torch.sort(torch.randperm(rows - 1, device=dev)[:U] + 1).values.to(torch.int32)
nU = torch.tensor([U], dtype=torch.int32, device=dev)
P = torch.empty(U, H, device=dev)
tabs = (_lib.Table * 1)(); tabs[0].weight = tab.data_ptr(); tabs[0].rows = rows; tabs[0].key_base = 0
dnn = _lib.Dnn(); dnn.w_item = W.data_ptr(); dnn.item_ld = 1024; dnn.table_side[0] = 0; dnn.table_col[0] = 0
st = torch.cuda.current_stream().cuda_stream
dW = torch.zeros(H, 1024, device=dev)
ws = torch.empty(lib.tgr_fact_backward_workspace_bytes(1, H), dtype=torch.uint8, device=dev)
def run(which):
    if which == "fwd":
        _lib.check(lib.tgr_fact_project_rows(tabs, 1, H, C.byref(dnn), uniq.data_ptr(), nU.data_ptr(), U, None, P.data_ptr(), st))
    else:
        _lib.check(lib.tgr_fact_unique_backward(tabs, 1, H, C.byref(dnn), uniq.data_ptr(), nU.data_ptr(), U, None, P.data_ptr(), dW.data_ptr(), None, ws.data_ptr(), ws.numel(), st))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for which in ("fwd", "bwd"):
    for _ in range(3): run(which)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(which); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{which} dbg={os.environ.get('TGR_ROWS_DBG','0')} U={U}: median {1e3*ts[len(ts)//2]:.1f} us  min {1e3*ts[0]:.1f} us")
