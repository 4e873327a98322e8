"""Device-resident item features (resident.py, csrc/tgr_resident.cu): the slim host call carries ids + user tokens only;
its host-side entry count must equal the packed call's, and the device-side expansion must rebuild the packed call bit for
bit (ids, arrays, mm inputs) — so every parity result of the packed path carries over unchanged."""
import numpy as np
import pytest
import torch

from tencent_recommendation_2025_b200.packed import count_valid
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld

STATS = {"103": 7, "104": 20, "105": 50, "109": 90, "100": 10, "117": 40, "111": 90, "118": 150, "101": 300, "102": 12,
         "119": 33, "120": 77, "114": 120, "112": 260, "121": 9, "115": 45, "122": 85, "116": 140, "106": 60, "107": 110,
         "108": 200, "110": 30}


def _world(mm=("81",)):
    cfg = SynthConfig(B=7, L=19, H=32, item_num=500, user_num=60, alpha=1.1, mm_ids=mm, min_len=3, feat_statistics=STATS)
    return SynthWorld(cfg, 4)


def test_slim_call_counts_entries_like_the_packed_call():
    from tencent_recommendation_2025_b200.resident import ResidentItemFeatures
    world = _world()
    store = ResidentItemFeatures.from_world(world, "cpu")
    for step in range(3):
        st = world.make_step(step)
        for pc in st.calls:
            sc = store.slim(pc, pin=False)
            assert sc.n_valid == count_valid(world.layout, pc)
            assert sc.nbytes < 0.25 * (pc.ids.nbytes + sum(x.nbytes for x in pc.mm_x))     # the point of the exercise
    # an out-of-range feature value is not an entry (the kernels skip it and raise the error flag)
    pc = world.make_step(0).calls[1]
    assert store.slim(pc, pin=False).n_valid == count_valid(world.layout, pc)


@pytest.mark.gpu
@pytest.mark.parametrize("mm,dt", [(("81",), torch.float32), (("81", "82"), torch.bfloat16)])
def test_expansion_rebuilds_the_packed_call_bit_for_bit(mm, dt):
    from tencent_recommendation_2025_b200.packed import to_device
    from tencent_recommendation_2025_b200.resident import ResidentFeeder, ResidentItemFeatures
    world = _world(mm)
    store = ResidentItemFeatures.from_world(world, "cuda", mm_dtype=dt)
    feeder = ResidentFeeder(store)
    steps = [world.make_step(s) for s in range(3)]
    feeder.submit([store.slim(pc) for pc in steps[0].calls])
    for i, st in enumerate(steps):
        pbs = feeder.take()
        if i + 1 < len(steps):
            feeder.submit([store.slim(pc) for pc in steps[i + 1].calls])
        for pc, pb in zip(st.calls, pbs):
            ref = to_device(world.layout, pc, "cuda", mm_dtype=dt)
            assert torch.equal(pb.ids, ref.ids)
            assert torch.equal(pb.arr_off, ref.arr_off) and torch.equal(pb.arr_val, ref.arr_val) and torch.equal(pb.arr_tok, ref.arr_tok)
            assert pb.arr_begin == ref.arr_begin and pb.arr_nnz == ref.arr_nnz and pb.n_valid == ref.n_valid
            for a, b in zip(pb.mm_x, ref.mm_x):
                assert a.dtype == b.dtype and torch.equal(a, b)
        feeder.retire()


@pytest.mark.gpu
def test_streaming_item_sweep_writes_the_same_files(tmp_path):
    """save_item_emb_resident (ids only, features resident, pinned ring, no per-chunk sync) writes the files the
    reference-signature save_item_emb (dict walk per 1024 items, model.py:402-433) writes."""
    import types

    from tencent_recommendation_2025_b200 import binfmt
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.resident import ResidentItemFeatures
    world = _world()
    cfg, lay = world.cfg, world.layout
    args = types.SimpleNamespace(device="cuda", hidden_units=cfg.H)
    torch.manual_seed(1)
    m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "parity", path="factored").cuda()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
        for p in m.engine.tables:
            p[0].zero_()
    store = ResidentItemFeatures.from_world(world, "cuda")
    item_ids = np.random.default_rng(0).permutation(np.arange(1, cfg.item_num + 1))[:437]
    retrieval = (item_ids.astype(np.uint64) + 10_000).tolist()
    d1, d2 = tmp_path / "stream", tmp_path / "dicts"
    d1.mkdir(); d2.mkdir()
    info = m.save_item_emb_resident(store, item_ids, retrieval, str(d1), chunk=100)
    assert info["items"] == item_ids.size
    # the reference-signature sweep on the same items: feature dicts from the same feature functions
    feats = world.item_sparse_values(item_ids.astype(np.int64))
    mmv = [world.mm_vectors(item_ids.astype(np.int64), j, d) for j, d in enumerate(lay.item_emb_feat.values())]
    feat_dict = {}
    for i in range(item_ids.size):
        d = {k: int(feats[i, j]) for j, k in enumerate(lay.item_sparse)}
        for k in lay.user_sparse:
            d[k] = 0
        for k in lay.user_array:
            d[k] = [0]
        for j, k in enumerate(lay.item_emb_feat):
            d[k] = mmv[j][i]
        feat_dict[i] = d
    m.save_item_emb(item_ids.tolist(), retrieval, feat_dict, str(d2), batch_size=128)
    a, b = binfmt.load_emb(d1 / "embedding.fbin"), binfmt.load_emb(d2 / "embedding.fbin")
    assert a.shape == b.shape == (item_ids.size, cfg.H)
    assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()
    assert np.array_equal(binfmt.load_emb(d1 / "id.u64bin", np.uint64), binfmt.load_emb(d2 / "id.u64bin", np.uint64))


def test_fixed_shape_slim_calls_pad_without_changing_content():
    """CallShape padding (graphed.GraphedStep): user tokens padded with -1, array values with id 0 behind the last token of
    each array; offsets rebased to the per-array capacity regions; host counts unchanged."""
    from tencent_recommendation_2025_b200.resident import CallShape, ResidentItemFeatures
    cfg = SynthConfig(B=5, L=9, H=32, item_num=200, user_num=40, alpha=1.1, mm_ids=("81",), min_len=2)
    world = SynthWorld(cfg, 1)
    steps = [world.make_step(s) for s in range(3)]
    store = ResidentItemFeatures.from_world(world, "cpu")
    shapes = [CallShape.covering([st.calls[i] for st in steps]) for i in range(3)]
    sigs = set()
    for st in steps:
        exact = store.slim_step(st.calls, pin=False)
        fixed = store.slim_step(st.calls, shapes, pin=False)
        sigs.add(tuple((tuple(c.offs), tuple(c.sizes), c.n_cap) for c in fixed.calls) + (fixed.ints.numel(),))
        for pc, e, f, sh in zip(st.calls, exact.calls, fixed.calls, shapes):
            assert f.n_valid == e.n_valid and e.n_cap is None and f.n_cap >= f.n_valid
            ev, fv = e.ints.numpy(), f.ints.numpy()
            assert np.array_equal(fv[f.offs[0]:f.offs[0] + f.sizes[0]], ev[e.offs[0]:e.offs[0] + e.sizes[0]])   # item ids
            tok = fv[f.offs[1]:f.offs[1] + f.sizes[1]]
            assert f.sizes[1] == sh.n_user_cap and np.array_equal(tok[:e.sizes[1]], ev[e.offs[1]:e.offs[1] + e.sizes[1]])
            assert (tok[e.sizes[1]:] == -1).all()
            n_arr = pc.arr_off.shape[0]
            off = fv[f.offs[3]:f.offs[3] + f.sizes[3]].reshape(n_arr, pc.T + 1)
            val = fv[f.offs[4]:f.offs[4] + f.sizes[4]]
            atok = fv[f.offs[5]:f.offs[5] + f.sizes[5]]
            for a in range(n_arr):
                assert off[a, 0] == f.arr_begin[a] and f.arr_nnz[a] == sh.arr_caps[a]
                for t in range(pc.T):
                    want = pc.arr_val[pc.arr_off[a, t]:pc.arr_off[a, t + 1]]
                    assert np.array_equal(val[off[a, t]:off[a, t + 1]], want)
                    assert (atok[off[a, t]:off[a, t + 1]] == t).all()
                assert (val[off[a, -1]:f.arr_begin[a] + f.arr_nnz[a]] == 0).all()
    assert len(sigs) == 1, "every step must have the same buffer layout"
