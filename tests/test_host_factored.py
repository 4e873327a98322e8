"""CPU-only checks of the factored / sharded host logic: size queries of the step driver, loud failure without a GPU,
argument validation, the default-layout facts the specialised kernels rely on. No compute calls."""
import ctypes as C
import types

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
