"""Parity at the SHAPE the benchmark times ("parity only at toy scale"): B = 256 sequences x L = 101,
H = 64, a 1 M-row Zipf item table, the default feature vocabularies (100 ... 10^6 rows), ~0.7 M lookups and ~10^5 unique
rows per step — so the persistent row kernels run several tiles per CTA and cross table boundaries inside a CTA, the
radix sort runs full tiles, and the segmented reduce stitches runs over many CTA tiles — plus one planted 19.5 k-duplicate
row (the heavy segment).

Checked against (a) numpy / torch.unique for everything integer (keys, dedup, counts: bit-exact) and (b) the torch oracle
evaluated in FLOAT64 as ground truth for every floating-point result: outputs, gradients of every parameter and the
AdamW-updated rows, all at the north_star bar of 1e-5 of tensor scale. Updated rows: AdamW's first step moves an element
by lr * g / (|g| + eps), which amplifies ANY fp32 rounding of g when |g| ~ eps; elements whose true gradient is below
the gradient tolerance itself (1e-5 of the table's gradient scale) are therefore only required to stay within the
2 * lr a sign flip can cost — for them the reference's own fp32 result is as far from the truth as ours.
"""
import types

import numpy as np
import pytest
import torch

from oracle import feat2emb_numpy as onp
from oracle.feat2emb_torch import TorchOracle, tensors_to_torch
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld

pytestmark = pytest.mark.gpu

RTOL = 1e-5
LR, WD = 1e-3, 1e-2
HOT_ID, HOT_COUNT = 4242, 19_500


@pytest.fixture(scope="module")
def scale_case():
    cfg = SynthConfig(B=256, L=101, H=64, item_num=1_000_000, user_num=200_000, alpha=1.05, mm_ids=("81",))
    world = SynthWorld(cfg, 3)
    lay = world.layout
    st = world.make_step(0)
    # plant the heavy segment: 19.5 k valid item tokens of the step look up the same item_emb row
    rng = np.random.default_rng(7)
    left = HOT_COUNT
    for pc in st.calls:
        valid = np.nonzero(pc.ids[:, 0] > 0)[0]
        take = min(left, valid.size // 2)
        pc.ids[rng.choice(valid, take, replace=False), 0] = HOT_ID
        pc.n_valid = None
        left -= take
    assert left == 0
    torch.manual_seed(11)
    orc = TorchOracle(lay).double()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in orc.parameters():
            p.copy_((torch.randn(p.shape, generator=g) * (0.05 if p.dim() >= 2 else 0.1)).double())
        for p in orc.table_params():
            p[0].zero_()
    params32 = {k: v.astype(np.float32) for k, v in orc.to_numpy().items()}
    with torch.no_grad():            # the truth is evaluated on the fp32-representable parameters
        for k, v in params32.items():
            orc.param(k).copy_(torch.from_numpy(v).double())
    n_is = len(lay.item_sparse)
    outs = []
    for pc in st.calls:
        # keep the call self-consistent after planting: ids column 0 is (mask == 1) * seq, the user-id column (mask == 2) * seq
        item_col = pc.ids[:, 0].reshape(pc.B, pc.L)
        pc.seq = (item_col + pc.ids[:, 1 + n_is].reshape(pc.B, pc.L) if pc.include_user else item_col).astype(pc.seq.dtype)
        t = {k: (v.double() if v.dtype.is_floating_point else v)
             for k, v in tensors_to_torch(onp.tensors_from_packed(lay, pc)).items()}
        seq = torch.from_numpy(pc.seq.astype(np.int64))
        mask = torch.from_numpy(pc.mask.astype(np.int64)) if pc.include_user else None
        outs.append(orc.feat2emb(seq, t, mask, pc.include_user))
    ups = [torch.from_numpy(r).double() for r in st.upstream]
    torch.autograd.backward(outs, ups)
    truth = {"out": [o.detach().numpy() for o in outs], "grad": orc.grads_numpy()}
    return cfg, lay, st, params32, truth


def _module(cfg, params32, mode):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    args = types.SimpleNamespace(device="cuda", hidden_units=cfg.H)
    m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, mode, path="factored").to("cuda")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params32.items()})
    return m


def _err(a, b):
    return float(np.abs(a.astype(np.float64) - b).max()), float(max(np.abs(b).max(), 1e-30))


def test_keys_sort_dedup_bit_exact_at_scale(scale_case):
    from tencent_recommendation_2025_b200.packed import to_device
    cfg, lay, st, params32, _ = scale_case
    m = _module(cfg, params32, "fused")
    pbs = [to_device(lay, pc, "cuda") for pc in st.calls]
    grp = m.engine.prepare(pbs)
    torch.cuda.synchronize()
    keys, srcs = onp.build_keys(lay, st.calls)
    uniq, counts = np.unique(keys, return_counts=True)
    U = int(grp.n_unique.item())
    assert U == uniq.size and grp.n == keys.size
    assert U > 3 * 296 * 128 // 4, "the case must give the persistent row kernels several tiles per CTA"
    assert np.array_equal(grp.uniq[:U].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, uniq.astype(np.int64))
    seg_off = grp._view(grp.c.seg_off, (U + 1,), torch.int32).cpu().numpy()
    assert np.array_equal(np.diff(seg_off), counts)
    assert counts.max() >= HOT_COUNT, "heavy segment missing"
    # stable: inside a key's run the sources stay in (call, token, slot) emission order
    ks = grp._view(grp.c.keys, (grp.n,), torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    ss = grp._view(grp.c.srcs, (grp.n,), torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ks, keys[order].astype(np.int64)) and np.array_equal(ss, srcs[order].astype(np.int64))


def test_forward_backward_vs_fp64_truth_at_scale(scale_case):
    from tencent_recommendation_2025_b200.packed import to_device
    cfg, lay, st, params32, truth = scale_case
    m = _module(cfg, params32, "parity")
    pbs = [to_device(lay, pc, "cuda") for pc in st.calls]
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    for c, (o, t) in enumerate(zip(outs, truth["out"])):
        e, s = _err(o.detach().cpu().numpy(), t)
        assert e <= RTOL * s, f"out call {c}: {e:.3e} vs scale {s:.3e}"
    torch.autograd.backward(outs, [torch.from_numpy(r).cuda() for r in st.upstream])
    torch.cuda.synchronize()
    named = dict(m.named_parameters())
    for k, gt in truth["grad"].items():
        p = named[k]
        if not gt.any():
            assert p.grad is None or not bool(p.grad.any()), k
            continue
        e, s = _err(p.grad.cpu().numpy(), gt)
        assert e <= RTOL * s, f"grad {k}: {e:.3e} vs scale {s:.3e}"
    hot = named["item_emb.weight"].grad[HOT_ID].cpu().numpy()
    e, s = _err(hot, truth["grad"]["item_emb.weight"][HOT_ID])
    assert e <= RTOL * s, f"heavy-segment row: {e:.3e} vs {s:.3e}"


def test_fused_row_update_vs_fp64_truth_at_scale(scale_case):
    from tencent_recommendation_2025_b200.packed import to_device
    cfg, lay, st, params32, truth = scale_case
    m = _module(cfg, params32, "fused")
    pbs = [to_device(lay, pc, "cuda") for pc in st.calls]
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, [torch.from_numpy(r).cuda() for r in st.upstream])
    m.fused_step(lr=LR, betas=(0.9, 0.98), eps=1e-8, weight_decay=WD)
    torch.cuda.synchronize()
    b1, b2, eps = 0.9, 0.98, 1e-8
    for i, t in enumerate(lay.tables):
        k = f"{t.name}.weight"
        g = truth["grad"][k]
        rows = np.nonzero(np.any(g != 0, axis=1))[0]
        got = m.engine.tables[i].detach().cpu().numpy()
        w0 = params32[k].astype(np.float64)
        untouched = np.setdiff1d(np.arange(got.shape[0]), rows)
        assert np.array_equal(got[untouched], params32[k][untouched]), f"{k}: untouched rows moved"
        if rows.size == 0:
            continue
        gr = g[rows]
        m1 = (1 - b1) * gr
        v1 = (1 - b2) * gr * gr
        want = w0[rows] * (1 - LR * WD) - (LR / (1 - b1)) * m1 / (np.sqrt(v1) / np.sqrt(1 - b2) + eps)   # torch/optim/adam.py step 1
        d = np.abs(got[rows].astype(np.float64) - want)
        gscale, wscale = np.abs(g).max(), max(np.abs(want).max(), 1e-30)
        firm = np.abs(gr) >= RTOL * gscale          # the gradient itself is resolved at the bar
        assert d[firm].max(initial=0.0) <= RTOL * wscale, f"{k}: updated rows off by {d[firm].max():.3e} (scale {wscale:.3e})"
        assert d[~firm].max(initial=0.0) <= 2.0 * LR * 1.001 + RTOL * wscale, f"{k}: sub-resolution gradients moved {d[~firm].max():.3e}"
        assert firm.mean() > 0.99, f"{k}: guard excuses too many elements ({1 - firm.mean():.4f})"
