"""GPU parity tests of the FACTORED path (csrc/tgr_factored.cu, factored.py): itemdnn/userdnn folded into the
deduplicated rows. Same oracle and bars as test_gpu_parity.py — the golden fixtures made from the unmodified
reference (outputs, gradients of EVERY parameter incl. itemdnn/userdnn/emb_transform, AdamW-updated parameters):
1e-5 of tensor scale in fp32, 1e-2 relative Frobenius under bf16 autocast; bitwise run-to-run reproducibility.
"""
import os
import types

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, Golden, TRAIN_CASES, assert_rows_updated
from test_gpu_parity import _close, _close_fro, dev_batch
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld, packed_to_dicts

pytestmark = pytest.mark.gpu


def make_module(g: Golden, mode="parity"):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    args = types.SimpleNamespace(device="cuda", hidden_units=g.H)
    m = BaselineEmbedding(g.user_num, g.item_num, g.feat_statistics, g.feat_types, args, mode, path="factored").to("cuda")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g.params0().items()})
    return m


@pytest.mark.parametrize("prefetch", [False, True])
@pytest.mark.parametrize("name", TRAIN_CASES)
def test_forward_matches_reference(name, prefetch):
    g = Golden(name)
    m = make_module(g)
    pbs = [dev_batch(m, pc) for pc in g.calls(0)]
    with torch.no_grad():
        if prefetch:
            m.prefetch(pbs)
        for c, pb in enumerate(pbs):
            out = m.feat2emb_packed(pb)
            assert out.shape == (pb.B, pb.L, g.H)
            _close(out, g.outs(0)[c], what=f"{name} out c{c} prefetch={prefetch}")


@pytest.mark.parametrize("name", ["baseline_h32", "o1_h64_mm2"])
def test_dict_signature_matches_reference(name):
    g = Golden(name)
    m = make_module(g)
    for c, pc in enumerate(g.calls(0)):
        dicts = packed_to_dicts(g.layout, pc)
        seq = torch.from_numpy(pc.seq)
        mask = torch.from_numpy(pc.mask) if pc.include_user else None
        out = m.feat2emb(seq, dicts, mask=mask, include_user=pc.include_user)
        _close(out, g.outs(0)[c], what=f"{name} dict call {c}")


def test_item_sweep_call_shape():
    g = Golden("item_sweep")
    m = make_module(g)
    dicts = packed_to_dicts(g.layout, g.sweep_call())
    with torch.no_grad():
        out = m.feat2emb(torch.from_numpy(g.z["seq"]).cuda(), dicts, include_user=False).squeeze(0)
    _close(out, g.z["out"][0], what="item sweep")


@pytest.mark.parametrize("prefetch", [False, True])
@pytest.mark.parametrize("name", TRAIN_CASES)
def test_backward_parity_mode_all_grads(name, prefetch):
    """Dense table gradients AND the itemdnn / userdnn / emb_transform gradients vs the reference's autograd."""
    g = Golden(name)
    m = make_module(g, "parity")
    pbs = [dev_batch(m, pc) for pc in g.calls(0)]
    if prefetch:
        m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    loss = sum((o * torch.from_numpy(r).cuda()).sum() for o, r in zip(outs, g.upstream(0)))
    loss.backward()
    ref = g.group("s0/grad/")
    named = dict(m.named_parameters())
    assert any(k.startswith("itemdnn") for k in ref), "fixture must hold the DNN gradients"
    for k, v in ref.items():
        p = named[k]
        if v.size == 0:
            assert p.grad is None or not bool(p.grad.any())
            continue
        assert p.grad is not None, k
        _close(p.grad, v, what=f"{name} grad {k} prefetch={prefetch}")
        if k.split(".")[0] in ("item_emb", "user_emb", "sparse_emb"):
            assert not bool(p.grad[0].any()), "padding row must get exactly zero grad"


def _fused_step(g, prefetch):
    m = make_module(g, "fused")
    dense_opt = torch.optim.AdamW(m.dense_parameters(), lr=g.lr, betas=(0.9, 0.98), weight_decay=g.wd)
    pbs = [dev_batch(m, pc) for pc in g.calls(0)]
    if prefetch:
        m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    loss = sum((o * torch.from_numpy(r).cuda()).sum() for o, r in zip(outs, g.upstream(0)))
    loss.backward()
    for p in m.engine.tables:
        assert p.grad is None
    dense_opt.step()
    m.fused_step(lr=g.lr, betas=(0.9, 0.98), eps=1e-8, weight_decay=g.wd)
    torch.cuda.synchronize()
    return m


@pytest.mark.parametrize("prefetch", [False, True])
@pytest.mark.parametrize("name", ["baseline_h32", "o1_h64"])
def test_fused_row_update_matches_reference_adamw(name, prefetch):
    g = Golden(name)
    m = _fused_step(g, prefetch)
    ref_g, ref_p, p0 = g.group("s0/grad/"), g.group("s0/param/"), g.params0()
    for k, p in m.named_parameters():
        got = p.detach().cpu().numpy()
        if k.split(".")[0] in ("item_emb", "user_emb", "sparse_emb"):
            touched = np.nonzero(np.any(ref_g[k] != 0, axis=1))[0]
            untouched = np.setdiff1d(np.arange(got.shape[0]), touched)
            assert_rows_updated(got[touched], ref_p[k][touched], ref_g[k][touched], g.lr, what=f"{name} updated rows {k}")
            assert np.array_equal(got[untouched], p0[k][untouched]), f"{k}: untouched rows must not move"
            assert not got[0].any()
        else:
            _close(got, ref_p[k], rtol=2e-4, what=f"{name} dense param {k}")


def test_fused_step_bitwise_reproducible():
    g = Golden("o1_h64")
    a = _fused_step(g, True)
    b = _fused_step(g, True)
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), f"{k} differs between two identical runs"


def test_multi_step_parity_mode_tracks_reference():
    g = Golden("baseline_h32")
    m = make_module(g, "parity")
    opt = torch.optim.AdamW(m.parameters(), lr=g.lr, betas=(0.9, 0.98), weight_decay=g.wd)
    named = dict(m.named_parameters())
    for step in range(g.n_steps):
        opt.zero_grad(set_to_none=True)
        pbs = [dev_batch(m, pc) for pc in g.calls(step)]
        m.prefetch(pbs)
        outs = [m.feat2emb_packed(pb) for pb in pbs]
        for c, o in enumerate(outs):
            _close(o, g.outs(step)[c], rtol=5e-5, what=f"step {step} out c{c}")
        loss = sum((o * torch.from_numpy(r).cuda()).sum() for o, r in zip(outs, g.upstream(step)))
        loss.backward()
        opt.step()
        ref_p = g.group(f"s{step}/param/")
        for k, p in named.items():
            _close(p, ref_p[k], rtol=3e-4, what=f"step {step} param {k}")


def test_bf16_autocast_within_1e2():
    g = Golden("o1_h64")
    ref = np.load(os.path.join(GOLDEN_DIR, "o1_h64_bf16.npz"))
    m = make_module(g, "parity")
    pbs = [dev_batch(m, pc) for pc in g.calls(0)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = [m.feat2emb_packed(pb) for pb in pbs]
    for c, o in enumerate(outs):
        assert o.dtype == torch.bfloat16
        _close_fro(o, ref[f"out{c}"], 1e-2, what=f"bf16 out c{c}")
    loss = sum((o.float() * torch.from_numpy(r).cuda()).sum() for o, r in zip(outs, g.upstream(0)))
    loss.backward()
    for k, p in m.named_parameters():
        key = f"grad/{k}"
        if key in ref.files:
            _close_fro(p.grad, ref[key], 1.5e-2, what=f"bf16 grad {k}")


@pytest.mark.parametrize("H", [32, 64, 128])
def test_hidden_sizes_vs_concat_path(H):
    """H in {32, 64, 128}: factored == concat path (itself checked against the oracle) incl. edge ids, empty
    arrays and an all-padding token, on outputs, all gradients and updated rows."""
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.packed import to_device
    stats = {k: 9 for k in ["103", "104", "105", "109", "100", "117", "111", "118", "101", "102", "119", "120", "114",
                            "112", "121", "115", "122", "116", "106", "107", "108", "110"]}
    cfg = SynthConfig(B=5, L=9, H=H, item_num=50, user_num=9, mm_ids=("81",), min_len=2, feat_statistics=stats)
    w = SynthWorld(cfg, 3)
    st = w.make_step(0)
    pc = st.calls[0]
    pc.ids[0, :] = 0
    pc.ids[1, 0] = cfg.item_num
    pc.ids[2, 1:15] = 9
    pc.n_valid = None
    args = types.SimpleNamespace(device="cuda", hidden_units=H)
    mods = []
    for path in ("concat", "factored"):
        torch.manual_seed(0)
        m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "parity", path=path).cuda()
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() == 1:
                    p.normal_(0, 0.1)
            for p in m.engine.tables:
                p[0].zero_()
        mods.append(m)
    mods[1].load_state_dict(mods[0].state_dict())
    res = []
    for m in mods:
        pbs = [to_device(m.layout, c, "cuda") for c in st.calls]
        m.prefetch(pbs)
        outs = [m.feat2emb_packed(pb) for pb in pbs]
        loss = sum((o * torch.from_numpy(r).cuda()).sum() for o, r in zip(outs, st.upstream))
        loss.backward()
        res.append((outs, {k: p.grad for k, p in m.named_parameters()}))
    for c, (a, b) in enumerate(zip(res[1][0], res[0][0])):
        _close(a, b, what=f"H={H} out c{c}")
    for k, gb in res[0][1].items():
        ga = res[1][1][k]
        if gb is None:
            assert ga is None or not bool(ga.any()), k
            continue
        _close(ga, gb, what=f"H={H} grad {k}")


def test_unsupported_hidden_raises():
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    g = Golden("baseline_h32")
    args = types.SimpleNamespace(device="cuda", hidden_units=48)
    with pytest.raises(ValueError):
        BaselineEmbedding(g.user_num, g.item_num, g.feat_statistics, g.feat_types, args, "parity", path="factored")


def test_incomplete_prefetched_group_raises():
    """A prefetched group whose calls do not all reach the loss cannot produce the group-level gradients."""
    g = Golden("baseline_h32")
    m = make_module(g, "parity")
    pbs = [dev_batch(m, pc) for pc in g.calls(0)]
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs[:2]]
    loss = sum(o.sum() for o in outs)
    with pytest.raises(RuntimeError):
        loss.backward()


@pytest.mark.parametrize("path", ["concat", "factored"])
def test_save_item_emb_writes_reference_format(path, tmp_path):
    """save_item_emb(item_ids, retrieval_ids, feat_dict, save_path, batch_size) (model.py:402-433): embedding.fbin holds
    the [1, n] sweep outputs of the reference fixture, id.u64bin the retrieval ids, in the reference's wire format."""
    from tencent_recommendation_2025_b200 import binfmt
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    g = Golden("item_sweep")
    args = types.SimpleNamespace(device="cuda", hidden_units=g.H)
    m = BaselineEmbedding(g.user_num, g.item_num, g.feat_statistics, g.feat_types, args, "parity", path=path).to("cuda")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g.params0().items()})
    dicts = packed_to_dicts(g.layout, g.sweep_call())[0]
    item_ids = [int(x) for x in g.z["seq"].reshape(-1)]
    n = len(item_ids)
    retrieval = list(range(1000, 1000 + n))
    m.save_item_emb(item_ids, retrieval, {i: dicts[i] for i in range(n)}, str(tmp_path), batch_size=max(1, n // 3 + 1))
    emb = binfmt.load_emb(tmp_path / "embedding.fbin", np.float32)
    ids = binfmt.load_emb(tmp_path / "id.u64bin", np.uint64)
    assert emb.shape == (n, g.H) and ids.shape == (n, 1)
    assert ids.reshape(-1).tolist() == retrieval
    _close(emb, g.z["out"][0], what=f"save_item_emb [{path}]")
