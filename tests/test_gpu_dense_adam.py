"""tgr_adam_dense (the one-launch AdamW of the path's Linear layers) against torch.optim.AdamW, through both callers:
``FactoredEngine`` (own_dense_parameters, tests/test_gpu_graphed.py covers the step) and ``ShardedBaselineEmbedding.dense_adam_``."""
import ctypes as C

import pytest
import torch

from tencent_recommendation_2025_b200 import _lib

pytestmark = pytest.mark.gpu


def _reference(ws, gs_per_step, lr, betas, eps, wd):
    ps = [torch.nn.Parameter(w.clone()) for w in ws]
    opt = torch.optim.AdamW(ps, lr=lr, betas=betas, eps=eps, weight_decay=wd)      # the single-tensor reference formula
    for gs in gs_per_step:
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
    return [p.detach() for p in ps]


@pytest.mark.parametrize("from_device_block", [False, True])
def test_adam_dense_matches_torch_adamw_over_steps(from_device_block):
    lib = _lib.load()
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [(64, 1024), (64,), (64, 576), (64,), (32, 32), (32,), (1000,), (3, 5, 7)]
    ws = [torch.randn(s, device=dev) * 0.1 for s in shapes]
    steps = [[torch.randn(s, device=dev) * (10.0 ** (i % 3 - 2)) for i, s in enumerate(shapes)] for _ in range(4)]
    lr, betas, eps, wd = 1e-3, (0.9, 0.98), 1e-8, 1e-2
    want = _reference(ws, steps, lr, betas, eps, wd)
    got = [w.clone() for w in ws]
    m = [torch.zeros_like(w) for w in ws]
    v = [torch.zeros_like(w) for w in ws]
    g = [torch.empty_like(w) for w in ws]
    dl = _lib.DenseList()
    dl.n = len(ws)
    for i in range(len(ws)):
        dl.w[i], dl.g[i], dl.m[i], dl.v[i], dl.numel[i] = got[i].data_ptr(), g[i].data_ptr(), m[i].data_ptr(), v[i].data_ptr(), got[i].numel()
    block = torch.zeros(C.sizeof(_lib.Adam) // 4, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for t, gs in enumerate(steps, start=1):
        for dst, src in zip(g, gs):
            dst.copy_(src)
        adam = _lib.make_adam(lr, betas[0], betas[1], eps, wd, t, 1.0)
        if from_device_block:
            host = torch.zeros_like(block, device="cpu")
            C.memmove(host.data_ptr(), C.addressof(adam), C.sizeof(adam))
            block.copy_(host)
            _lib.check(lib.tgr_adam_dense(C.byref(dl), None, block.data_ptr(), stream), "tgr_adam_dense")
        else:
            _lib.check(lib.tgr_adam_dense(C.byref(dl), C.addressof(adam), None, stream), "tgr_adam_dense")
    torch.cuda.synchronize()
    for a, b, s in zip(got, want, shapes):
        err = (a - b).abs().max().item()
        assert err <= 1e-6 * max(b.abs().max().item(), 1e-30), f"{s}: {err:.3e}"


def test_adam_dense_rejects_bad_lists():
    lib = _lib.load()
    dl = _lib.DenseList()
    dl.n = 0
    adam = _lib.make_adam(1e-3, 0.9, 0.98, 1e-8, 1e-2, 1, 1.0)
    assert lib.tgr_adam_dense(C.byref(dl), C.addressof(adam), None, None) != 0          # empty list
    w = torch.zeros(8, device="cuda")
    dl.n = 1
    dl.w[0], dl.g[0], dl.m[0], dl.v[0], dl.numel[0] = w.data_ptr(), w.data_ptr(), w.data_ptr(), w.data_ptr(), 8
    assert lib.tgr_adam_dense(C.byref(dl), None, None, None) != 0                        # neither host nor device block
    assert lib.tgr_adam_dense(C.byref(dl), C.addressof(adam), w.data_ptr(), None) != 0   # both
