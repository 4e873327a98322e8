"""CPU-only checks of the factored / sharded host logic: size queries of the step driver, loud failure without a GPU,
argument validation, the default-layout facts the specialised kernels rely on. No compute calls."""
import ctypes as C
import types

import numpy as np
import pytest
import torch

from tencent_recommendation_2025_b200 import _lib, build
from tencent_recommendation_2025_b200.layout import DEFAULT_FEAT_TYPES, KIND_MM, KIND_SINGLE, FeatureLayout, default_feat_statistics


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _layout(H=64, mm=("81",)):
    ft = {k: list(v) for k, v in DEFAULT_FEAT_TYPES.items()}
    ft["item_emb"] = list(mm)
    return FeatureLayout(1000, 5000, default_feat_statistics(), ft, H)


def _group(lay, T=1000, n=12345):
    g = _lib.FactGroup()
    g.n_calls, g.H, g.key_bits, g.n_mm, g.n = 3, lay.H, lay.key_bits, len(lay.item_emb_feat), n
    for f, d in enumerate(lay.item_emb_feat.values()):
        g.mm_dim[f] = d
    for i, inc in enumerate((True, False, False)):
        cl = lay.calls[inc]
        c = g.calls[i]
        c.T, c.n_slots, c.n_single, c.n_arrays = T, len(cl.slots), cl.n_single, cl.n_array
        for j, s in enumerate(cl.slots):
            c.slots[j].kind, c.slots[j].side, c.slots[j].col, c.slots[j].table, c.slots[j].src = s.kind, s.side, s.col, s.table, s.src
        for a in range(cl.n_array):
            c.arr_begin[a], c.arr_nnz[a] = 10 * a, 10
    return g


def test_group_arena_size_query(lib):
    lay = _layout()
    small = lib.tgr_fact_group_bytes(C.byref(_group(lay, n=1000)), len(lay.tables))
    big = lib.tgr_fact_group_bytes(C.byref(_group(lay, n=100000)), len(lay.tables))
    assert 0 < small < big
    # P, G, rows_local and the four pair arrays scale with n: >= 3 * n * H * 4 bytes
    assert big >= 3 * 100000 * lay.H * 4
    bad = _group(lay)
    bad.H = 48
    assert lib.tgr_fact_group_bytes(C.byref(bad), len(lay.tables)) == 0
    assert b"{32, 64, 128}" in lib.tgr_last_error()


def test_step_driver_argument_errors(lib):
    lay = _layout()
    g = _group(lay)
    assert lib.tgr_fact_prepare(None, 24, C.byref(g), None, 0, None) < 0
    assert lib.tgr_fact_call_forward(None, 24, None, C.byref(g), 0, None, None) < 0
    assert lib.tgr_fact_call_backward(None, 24, None, C.byref(g), 0, None, None, 0, None) < 0
    assert lib.tgr_fetch_peer_rows(None, 2, 64, None, None, 10, None, None) < 0
    assert lib.tgr_bwd_reduce_rows(None, 1, 64, None, 0, None, None, 10, None, None, 0, None) < 0
    assert lib.tgr_sort_pairs(None, None, None, None, 10, 24, None, 0, None) < 0
    assert lib.tgr_sort_workspace_bytes(3_000_000) >= 2 * 3_000_000 * 4


def test_default_layout_matches_the_specialised_forward():
    """fact_forward_kernel<.., 15, 0> / <.., 15, 5>: with the reference's default feature lists (dataset.py:191-212)
    the SINGLE slots are ids columns 0..14 (item side) then 15..19 (user side), each table feeding one slot."""
    lay = _layout()
    for inc, (nsi, nsu) in ((False, (15, 0)), (True, (15, 5))):
        cl = lay.calls[inc]
        singles = [s for s in cl.slots if s.kind == KIND_SINGLE]
        item = [s.src for s in singles if s.side == 0]
        user = [s.src for s in singles if s.side == 1]
        assert item == list(range(nsi)) and user == list(range(nsi, nsi + nsu))
    tables = [s.table for s in lay.calls[True].slots if s.kind != KIND_MM]
    assert len(tables) == len(set(tables)) == len(lay.tables)


def test_factored_engine_has_no_cpu_fallback(lib):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.packed import to_device
    from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld
    stats = {k: 9 for k in default_feat_statistics()}
    cfg = SynthConfig(B=2, L=5, H=32, item_num=50, user_num=9, mm_ids=("81",), min_len=2, feat_statistics=stats)
    w = SynthWorld(cfg, 0)
    args = types.SimpleNamespace(device="cpu", hidden_units=32)
    m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "fused", path="factored")
    pb = to_device(m.layout, w.make_step(0).calls[0], "cpu", pin=False)
    with pytest.raises(_lib.TgrError):
        m.feat2emb_packed(pb)
    with pytest.raises(ValueError):
        BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(),
                          types.SimpleNamespace(device="cpu", hidden_units=48), "fused", path="factored")
    with pytest.raises(ValueError):
        BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "fused", path="nope")


def test_sharded_module_rejects_unknown_path():
    from tencent_recommendation_2025_b200.sharded import ShardedBaselineEmbedding
    stats = {k: 9 for k in default_feat_statistics()}
    ft = {k: list(v) for k, v in DEFAULT_FEAT_TYPES.items()}
    with pytest.raises(ValueError):
        ShardedBaselineEmbedding(9, 50, stats, ft, types.SimpleNamespace(device="cpu", hidden_units=32), 0, 1, path="nope")


def _ref_style_batch(lay, st):
    """What the reference's MyDataset.__getitem__ yields per sequence (dataset.py:96-168): id arrays + dict arrays."""
    from tencent_recommendation_2025_b200.synth import packed_to_dicts
    seq_c, pos_c, neg_c = st.calls
    ds, dp, dn = (packed_to_dicts(lay, pc) for pc in st.calls)
    batch = []
    for b in range(seq_c.B):
        tt = seq_c.mask[b]
        batch.append((seq_c.seq[b], pos_c.seq[b], neg_c.seq[b], tt, tt, tt,
                      np.array(ds[b], dtype=object), np.array(dp[b], dtype=object), np.array(dn[b], dtype=object)))
    return batch


def _ref_collate(batch):   # the reference's collate_fn, dataset.py:268-293
    seq, pos, neg, token_type, next_token_type, next_action_type, seq_feat, pos_feat, neg_feat = zip(*batch)
    t = lambda x: torch.from_numpy(np.array(x))
    return t(seq), t(pos), t(neg), t(token_type), t(next_token_type), t(next_action_type), list(seq_feat), list(pos_feat), list(neg_feat)


def test_packing_collate_and_packed_dispatch(lib):
    """PackingCollate replaces the three feature lists by packed calls equal to pack_from_dicts'; feat2emb accepts
    them in the feature_array position (here it must get as far as the engine, which refuses CPU tables)."""
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.packed import HostPacked, PackingCollate, pack_from_dicts, to_device
    from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld
    stats = {k: 9 for k in default_feat_statistics()}
    cfg = SynthConfig(B=3, L=6, H=32, item_num=50, user_num=9, mm_ids=("81",), min_len=2, feat_statistics=stats)
    w = SynthWorld(cfg, 1)
    st = w.make_step(0)
    args = types.SimpleNamespace(device="cpu", hidden_units=32)
    m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "fused", path="factored")
    lay = m.layout
    out = PackingCollate(lay, _ref_collate)(_ref_style_batch(lay, st))
    assert len(out) == 9 and all(isinstance(out[i], HostPacked) for i in (6, 7, 8))
    assert out[6].group is out[7].group is out[8].group
    for hp, pc in zip(out[6:9], st.calls):
        pb = hp.upload("cpu")
        assert pb.include_user == pc.include_user and np.array_equal(pb.ids.numpy(), pc.ids)
        assert np.array_equal(pb.arr_val.numpy(), pc.arr_val) and np.array_equal(pb.mm_x[0].numpy(), pc.mm_x[0])
    seq, mask = out[0], out[3]
    for feats, s_, inc in ((out[6], seq, True), (st.calls[1], out[1], False), (to_device(lay, st.calls[2], "cpu", pin=False), out[2], False)):
        with pytest.raises(_lib.TgrError):        # packed input accepted, dict walk skipped, engine reached
            m.feat2emb(s_, feats, mask=mask if inc else None, include_user=inc)
    with pytest.raises(_lib.TgrError):            # the plain list-of-dict signature takes the same way in
        m.feat2emb(seq, _ref_collate(_ref_style_batch(lay, st))[6], mask=mask, include_user=True)
    with pytest.raises(ValueError):
        m.feat2emb(seq, out[7], mask=mask, include_user=True)            # a pos call passed as the seq call
    with pytest.raises(ValueError):
        m.feat2emb(seq[:, :-1], out[6], mask=mask, include_user=True)    # shape disagreement
