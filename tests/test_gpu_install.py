"""The 3-line graft onto a REAL reference model (``install(model)`` had zero tests).

The unmodified ``BaselineModel`` (staged copy of /root/reference/model/BaseLine/model.py under baseline/_ref, see
baseline/ref_loader.py) is built on cuda, initialised with the reference's own loop (main.py:95-111), and run through its
own ``forward`` (model.py:354-384: log2feats -> feat2emb x3 -> logits) with list-of-dict features. A deep copy gets
``tgr.install(...)`` in each of the three documented forms and must reproduce the stock model: logits, gradients of every
parameter (parity mode) and the parameters after one optimizer step (fused mode, tables against the reference's dense
AdamW on the rows it touched). Also: ``torch.compile(model)`` (main.py:114-115) over the grafted model, packed calls in the
feature_array position, and the GradScaler-skipped step."""
import copy
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import ref_loader  # noqa: E402
from golden_util import assert_rows_updated  # noqa: E402
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld, packed_to_dicts  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not staged")]

STATS = {k: 40 for k in ["103", "104", "105", "109", "100", "117", "111", "118", "101", "102", "119", "120", "114", "112",
                         "121", "115", "122", "116", "106", "107", "108", "110"]}
LR = 1e-3
# These comparisons run THROUGH the reference's trunk (attention + LayerNorm, torch fp32 kernels on both sides): a 1e-6
# difference at feat2emb's output is amplified by the normalisations on the way to the logits and back. The graded 1e-5 bar
# is enforced AT feat2emb's boundary with injected upstream gradients (test_gpu_parity / _factored / _scale); here the bar
# is 1e-4 of tensor scale end to end.
E2E_RTOL = 1e-4


def _case(variant="BaseLine", H=64):
    cfg = SynthConfig(B=6, L=17, H=H, item_num=400, user_num=30, alpha=1.2, mm_ids=("81",), min_len=4, feat_statistics=STATS)
    world = SynthWorld(cfg, 2)
    st = world.make_step(0)
    Model = ref_loader.load_model_class(variant)
    args = types.SimpleNamespace(device="cuda", norm_first=False, maxlen=cfg.L - 1, hidden_units=H, num_blocks=1, num_heads=1,
                                 dropout_rate=0.0)
    torch.manual_seed(0)
    model = Model(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args).to("cuda")
    # the reference's init loop (main.py:95-111) ...
    for _, p in model.named_parameters():
        if p.dim() >= 2:
            torch.nn.init.xavier_normal_(p.data)
        elif p.dim() == 1:
            torch.nn.init.constant_(p.data, 0.0)
    model.pos_emb.weight.data[0, :] = 0
    model.item_emb.weight.data[0, :] = 0
    model.user_emb.weight.data[0, :] = 0
    for k in model.sparse_emb:
        model.sparse_emb[k].weight.data[0, :] = 0
    # ... then non-zero biases so the bias gradients are exercised
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.normal_(0.0, 0.1)
    lay = world.layout
    seqc, posc, negc = st.calls
    feats = [packed_to_dicts(lay, pc) for pc in st.calls]
    seq = torch.from_numpy(seqc.seq).cuda()                         # main.py:174-176 moves the id tensors, the masks stay on the host
    mask = torch.from_numpy(seqc.mask)
    pos, neg = torch.from_numpy(posc.seq).cuda(), torch.from_numpy(negc.seq).cuda()
    next_mask = torch.from_numpy((posc.seq != 0).astype(np.int32))
    batch = (seq, pos, neg, mask, next_mask, None, feats[0], feats[1], feats[2])
    return cfg, world, st, model, batch


def _loss(model, batch):
    pos_logits, neg_logits = model(*batch)
    return (pos_logits * 0.7 - neg_logits * 0.3).sum() + (pos_logits ** 2).sum() * 0.1


def _hot(name):
    return name.split(".")[0] in ("item_emb", "user_emb", "sparse_emb")


@pytest.mark.parametrize("path", ["concat", "factored"])
def test_install_parity_mode_reproduces_the_reference_model(path):
    import tencent_recommendation_2025_b200 as tgr
    cfg, world, st, ref, batch = _case()
    mine = copy.deepcopy(ref)
    tgr.install(mine, path=path)                                    # form 1 / 2: parity mode, the reference optimizer stays
    mine.check_padding_rows()
    opt_r = torch.optim.AdamW(ref.parameters(), lr=LR, betas=(0.9, 0.98))
    opt_m = torch.optim.AdamW(mine.parameters(), lr=LR, betas=(0.9, 0.98))
    lr_, lm_ = _loss(ref, batch), _loss(mine, batch)
    assert abs(lr_.item() - lm_.item()) <= E2E_RTOL * max(abs(lr_.item()), 1.0)
    lr_.backward()
    lm_.backward()
    gm = dict(mine.named_parameters())
    # a gradient that is zero in exact arithmetic (the key bias under a softmax: every score of a row shifts by the same
    # amount) is pure rounding noise on both sides, so each tensor's scale is floored at 1e-3 of the largest gradient
    gscale = max(p.grad.abs().max().item() for p in ref.parameters() if p.grad is not None)
    for k, p in ref.named_parameters():
        if p.grad is None:
            assert gm[k].grad is None or not bool(gm[k].grad.any()), k
            continue
        assert gm[k].grad is not None, k
        err = (gm[k].grad - p.grad).abs().max().item()
        assert err <= E2E_RTOL * max(p.grad.abs().max().item(), 1e-3 * gscale), f"grad {k}: {err:.3e}"
    opt_r.step()
    opt_m.step()
    for k, p in ref.named_parameters():
        if _hot(k):
            g = p.grad.cpu().numpy()
            rows = np.nonzero(np.any(g != 0, axis=1))[0]
            got, want = gm[k].detach().cpu().numpy(), p.detach().cpu().numpy()
            assert_rows_updated(got[rows], want[rows], g[rows], LR, rtol=E2E_RTOL, what=f"{path} {k}")
            rest = np.setdiff1d(np.arange(got.shape[0]), rows)
            np.testing.assert_allclose(got[rest], want[rest], rtol=0, atol=E2E_RTOL * max(np.abs(want).max(), 1e-30))
    assert set(mine.state_dict()) == set(ref.state_dict()), "state_dict keys must stay the reference's"


def test_install_fused_mode_row_update_and_scaler_skip():
    import tencent_recommendation_2025_b200 as tgr
    cfg, world, st, ref, batch = _case()
    mine = copy.deepcopy(ref)
    p0 = {k: p.detach().clone() for k, p in ref.named_parameters()}
    opt_r = torch.optim.AdamW(ref.parameters(), lr=LR, betas=(0.9, 0.98))
    tables = {id(p) for k, p in mine.named_parameters() if _hot(k)}
    opt_m = torch.optim.AdamW([p for p in mine.parameters()], lr=LR, betas=(0.9, 0.98))
    scaler = torch.amp.GradScaler("cuda", enabled=True, init_scale=1.0)
    tgr.install(mine, opt_m, mode="fused", path="factored", scaler=scaler)   # form 3
    _loss(ref, batch).backward()
    opt_r.step()
    scaler.scale(_loss(mine, batch)).backward()
    for k, p in mine.named_parameters():
        if _hot(k):
            assert p.grad is None, f"{k}: fused mode must not materialise dense table gradients"
    scaler.step(opt_m)
    scaler.update()
    torch.cuda.synchronize()
    gm = dict(mine.named_parameters())
    for k, p in ref.named_parameters():
        if not _hot(k):
            continue
        g = p.grad.cpu().numpy()
        rows = np.nonzero(np.any(g != 0, axis=1))[0]
        got, want = gm[k].detach().cpu().numpy(), p.detach().cpu().numpy()
        assert_rows_updated(got[rows], want[rows], g[rows], LR, rtol=E2E_RTOL, what=f"fused {k}")
        untouched = np.setdiff1d(np.arange(got.shape[0]), rows)
        assert np.array_equal(got[untouched], p0[k].cpu().numpy()[untouched]), f"{k}: lazy rows must not move"
    # a step the scaler skips (inf gradients) must not leave row gradients queued for the next one
    loss = _loss(mine, batch)
    (loss * float("inf")).backward()
    before = {k: p.detach().clone() for k, p in mine.named_parameters() if _hot(k)}
    scaler.step(opt_m)        # optimizer.step() is skipped: the post-hook never fires
    scaler.update()
    assert not mine._tgr_engine.ready and not mine._tgr_engine.pending
    for k, p in mine.named_parameters():
        if _hot(k):
            assert torch.equal(p, before[k]), f"{k} moved in a skipped step"


def test_install_survives_torch_compile_and_packed_calls():
    import tencent_recommendation_2025_b200 as tgr
    from tencent_recommendation_2025_b200.packed import PackingCollate, pack_from_dicts, stage_pinned, StepGroup
    cfg, world, st, ref, batch = _case()
    mine = copy.deepcopy(ref)
    tgr.install(mine, path="factored")
    with torch.no_grad():
        want = ref(*batch)
    compiled = torch.compile(mine)                                  # main.py:114-115
    with torch.no_grad():
        got = compiled(*batch)
    for a, b in zip(got, want):
        assert (a - b).abs().max().item() <= E2E_RTOL * max(b.abs().max().item(), 1e-30)
    # packed calls in the feature_array position (what PackingCollate hands over): same logits, no dict walk
    lay = mine._tgr_layout
    seq, pos, neg, mask = batch[0], batch[1], batch[2], batch[3]
    calls = [pack_from_dicts(lay, seq, batch[6], mask, True), pack_from_dicts(lay, pos, batch[7], None, False),
             pack_from_dicts(lay, neg, batch[8], None, False)]
    hps = [stage_pinned(lay, pc) for pc in calls]
    grp = StepGroup(hps)
    for hp in hps:
        hp.group = grp
    with torch.no_grad():
        got2 = mine(seq, pos, neg, mask, batch[4], None, hps[0], hps[1], hps[2])
    for a, b in zip(got2, want):
        assert (a - b).abs().max().item() <= E2E_RTOL * max(b.abs().max().item(), 1e-30)
