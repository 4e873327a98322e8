"""tcgen05 / TMA kernels against plain torch (GPU): the bf16 tensor-core mm projection (csrc/tgr_mm_tc.cu) and the
tcgen05 forward row projection (csrc/tgr_factored.cu fact_rows_tc_kernel vs the mma.sync and FFMA variants)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("T,mm_dim,H", [(1000, 128, 64), (4099, 1024, 64), (777, 1024, 32), (513, 256, 128), (128, 3584, 64)])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_mm_proj_fwd_tc_matches_torch(T, mm_dim, H, out_bf16):
    from tencent_recommendation_2025_b200 import _lib
    lib = _lib.load()
    assert lib.tgr_mm_proj_fwd_tc_supported(_lib.DTYPE_BF16, mm_dim, H) == 1
    g = torch.Generator(device="cuda").manual_seed(T + mm_dim + H)
    x = torch.randn(T, mm_dim, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(H, mm_dim, device="cuda", generator=g) * 0.05)
    b = torch.randn(H, device="cuda", generator=g) * 0.1
    Wb = torch.empty(H, mm_dim, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.tgr_cast_bf16(W.data_ptr(), W.numel(), Wb.data_ptr(), _stream()), "cast")
    assert torch.equal(Wb, W.to(torch.bfloat16))                      # RNE, bit-exact with torch's cast
    ld = H + 8                                                        # strided output rows (a slot of a wider buffer)
    out = torch.full((T, ld), -7.0, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _lib.check(lib.tgr_mm_proj_fwd_tc(x.data_ptr(), T, mm_dim, Wb.data_ptr(), b.data_ptr(), H, out.data_ptr(), ld,
                                      _lib.DTYPE_BF16 if out_bf16 else _lib.DTYPE_F32, _stream()), "mm_proj_fwd_tc")
    torch.cuda.synchronize()
    ref = x.double() @ Wb.double().t() + b.double()                   # exact products of the bf16 operands
    got = out[:, :H].double()
    scale = ref.abs().max().item()
    tol = 1e-2 if out_bf16 else 1e-5                                  # north_star: 1e-2 (bf16) / 1e-5 (fp32) of tensor scale
    assert (got - ref).abs().max().item() <= tol * scale
    assert torch.all(out[:, H:] == -7.0)                              # nothing written outside the slot


@pytest.mark.parametrize("H", [32, 64])
def test_rows_projection_variants_agree(H):
    """tcgen05 (default), mma.sync (TGR_ROWS_TC=0) and FFMA (TGR_ROWS_FFMA=1) forward projections against fp64; the
    variants are selected by environment at library load, so each runs in its own interpreter."""
    code = f"""
import ctypes as C, os, sys, torch
sys.path.insert(0, {ROOT!r})
from tencent_recommendation_2025_b200 import _lib
lib = _lib.load()
H, rows = {H}, 40000
torch.manual_seed(0)
tabs_t = [torch.randn(r, H, device='cuda') * 0.1 for r in (rows, 300, 5000)]
W = torch.randn(H, 3 * H, device='cuda') * 0.2
keys = []
base = 0
for i, t in enumerate(tabs_t):
    n = min(t.shape[0] - 1, (3001, 299, 1777)[i])
    keys.append(torch.sort(torch.randperm(t.shape[0] - 1, device='cuda')[:n] + 1).values + base)
    base += t.shape[0]
uniq = torch.cat(keys).to(torch.int32)
U = uniq.numel()
nU = torch.tensor([U], dtype=torch.int32, device='cuda')
P = torch.zeros(U + 5, H, device='cuda')
tabs = (_lib.Table * 3)()
base = 0
for i, t in enumerate(tabs_t):
    tabs[i].weight, tabs[i].rows, tabs[i].key_base = t.data_ptr(), t.shape[0], base
    base += t.shape[0]
dnn = _lib.Dnn(); dnn.w_item = W.data_ptr(); dnn.item_ld = 3 * H
for i in range(3):
    dnn.table_side[i] = 0; dnn.table_col[i] = i * H
_lib.check(lib.tgr_fact_project_rows(tabs, 3, H, C.byref(dnn), uniq.data_ptr(), nU.data_ptr(), U + 5, None, P.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
ref = []
for i, (t, k) in enumerate(zip(tabs_t, keys)):
    b0 = int(tabs[i].key_base)
    ref.append(t[(k - b0).long()].double() @ W[:, i * H:(i + 1) * H].double().t())
ref = torch.cat(ref)
err = (P[:U].double() - ref).abs().max().item() / ref.abs().max().item()
assert torch.all(P[U:] == 0), 'rows past n_unique were written'
print('ERR', err)
"""
    for env in ({}, {"TGR_ROWS_TC": "0"}, {"TGR_ROWS_FFMA": "1"}):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        err = float(r.stdout.strip().split("ERR")[-1])
        assert err <= 1e-5, (env, err)
