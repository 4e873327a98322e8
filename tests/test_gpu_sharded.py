"""Row-sharded path on ONE GPU with W emulated ranks (run_emulated): the CUDA kernels + the exchange logic.
Bit-exact: routing / ownership / forward rows; W=1 equals the unsharded fused step bitwise; W>1 updated
rows within 1e-5 of the oracle (fp64 reduction -> fp32 AdamW rows)."""
import types

import numpy as np
import pytest
import torch

from golden_util import assert_rows_updated

from oracle import feat2emb_numpy as onp
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld

pytestmark = pytest.mark.gpu

STATS = {"103": 7, "104": 20, "105": 50, "109": 90, "100": 10, "117": 40, "111": 90, "118": 150, "101": 3000,
         "102": 12, "119": 33, "120": 77, "114": 120, "112": 2600, "121": 9, "115": 45, "122": 85, "116": 140,
         "106": 60, "107": 110, "108": 200, "110": 30}
HYPER = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)


def setup(W, B=16, L=33, H=64, seed=3):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.packed import to_device
    from tencent_recommendation_2025_b200.sharded import CudaShardOps, ShardedRank, shard_of_tables
    cfg = SynthConfig(B=B, L=L, H=H, item_num=5000, user_num=300, alpha=1.2, mm_ids=("81",), min_len=5,
                      feat_statistics=STATS)
    world = SynthWorld(cfg, seed)
    args = types.SimpleNamespace(device="cuda", hidden_units=H)
    torch.manual_seed(seed)
    full = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "fused").cuda()
    with torch.no_grad():
        for p in full.engine.tables:
            p.normal_(0, 0.1)
            p[0].zero_()
    lay = full.layout
    tables = [p.data for p in full.engine.tables]
    ranks = []
    for r in range(W):
        ops = CudaShardOps(lay, shard_of_tables(tables, r, W), dict(full.emb_transform.items()), W)
        ranks.append(ShardedRank(lay, ops, r, W))
    steps = [world.make_step(r) for r in range(W)]                       # rank r's data-parallel share
    pbs = [[to_device(lay, pc, "cuda") for pc in st.calls] for st in steps]
    return cfg, full, lay, ranks, steps, pbs


@pytest.mark.parametrize("W", [1, 2, 4, 8])
def test_sharded_forward_bit_exact(W):
    from tencent_recommendation_2025_b200.sharded import run_emulated
    cfg, full, lay, ranks, steps, pbs = setup(W)
    for c in range(3):
        outs = run_emulated([ranks[r].forward_gen(pbs[r][c]) for r in range(W)])
        for r in range(W):
            ref_item, ref_user = full.engine.forward(pbs[r][c])
            assert torch.equal(outs[r][0], ref_item), f"W={W} rank {r} call {c}: item concat differs"
            if ref_user is not None:
                assert torch.equal(outs[r][1], ref_user)
            # routing facts vs the numpy restatement
            keys, _ = onp.build_keys(lay, [steps[r].calls[c]])
            uniq = np.unique(keys)
            owner, local, counts, order = onp.route(uniq, W)
            assert ranks[r].last_fwd["send_counts"] == counts.tolist()
            assert ranks[r].last_fwd["U"] == uniq.size
        for r in range(W):   # what r receives from s is what s sends to r
            assert ranks[r].last_fwd["recv_counts"] == [ranks[s].last_fwd["send_counts"][r] for s in range(W)]


def _dcats(lay, st, seed):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for pc in st.calls:
        cl = lay.calls[pc.include_user]
        di = torch.randn((pc.T, cl.item_dim), generator=gen)
        du = torch.randn((pc.T, cl.user_dim), generator=gen) if pc.include_user else None
        out.append((di, du))
    return out


def test_sharded_w1_step_equals_unsharded_bitwise():
    from tencent_recommendation_2025_b200.sharded import run_emulated, tables_from_shards
    cfg, full, lay, ranks, steps, pbs = setup(1)
    d = _dcats(lay, steps[0], 7)
    for pb, (di, du) in zip(pbs[0], d):
        full.engine.queue(pb, di.cuda(), None if du is None else du.cuda())
        ranks[0].queue(pb, di.cuda(), None if du is None else du.cuda())
    full.fused_step(**HYPER)
    run_emulated([ranks[0].step_gen(dict(HYPER))])
    got = tables_from_shards(lay, [ranks[0].ops.local])
    for g, p, t in zip(got, full.engine.tables, lay.tables):
        assert torch.equal(g, p.data), t.name


@pytest.mark.parametrize("W", [2, 4])
def test_sharded_step_matches_oracle_and_is_deterministic(W):
    from tencent_recommendation_2025_b200.sharded import run_emulated, tables_from_shards
    results = []
    for rep in range(2):
        cfg, full, lay, ranks, steps, pbs = setup(W)
        before = [p.data.clone() for p in full.engine.tables]
        dc = [_dcats(lay, steps[r], 100 + r) for r in range(W)]
        for r in range(W):
            for pb, (di, du) in zip(pbs[r], dc[r]):
                ranks[r].queue(pb, di.cuda(), None if du is None else du.cuda())
        run_emulated([ranks[r].step_gen(dict(HYPER)) for r in range(W)])
        got = tables_from_shards(lay, [rk.ops.local for rk in ranks])
        results.append(got)
        if rep:
            continue
        # oracle: fp64 per-key gradient summed over ranks, then the fp32 AdamW row formula
        tot = {}
        for r in range(W):
            dnp = [(di.numpy(), None if du is None else du.numpy()) for di, du in dc[r]]
            uq, rows = onp.segment_reduce_fp64(lay, steps[r].calls, dnp)
            for k, row in zip(uq.tolist(), rows):
                tot[k] = tot.get(k, 0) + row
        keys = np.array(sorted(tot), np.int64)
        g = np.stack([tot[k] for k in keys]).astype(np.float32)
        m_tabs = tables_from_shards(lay, [rk.ops.exp_avg for rk in ranks])
        for ti, (t, b, a) in enumerate(zip(lay.tables, before, got)):
            sel = (keys >= t.key_base) & (keys < t.key_base + t.rows)
            rows_t = keys[sel] - t.key_base
            w = b.cpu().numpy().copy()
            m, v = np.zeros_like(w), np.zeros_like(w)
            onp.adamw_rows(w, m, v, rows_t, g[sel], 1, lr=1e-3, wd=1e-2)
            an = a.cpu().numpy()
            assert_rows_updated(an[rows_t], w[rows_t], g[sel], 1e-3, what=t.name)   # 1e-5, |g| guard (golden_util)
            # exp_avg is linear in the reduced gradient: the strict 1e-5 check of the reduction
            mg = m_tabs[ti].cpu().numpy()
            assert np.abs(mg - m).max() <= 1e-5 * max(np.abs(m).max(), 1e-30), t.name + " exp_avg"
            untouched = np.setdiff1d(np.arange(t.rows), rows_t)
            assert np.array_equal(an[untouched], b.cpu().numpy()[untouched]), f"{t.name}: untouched rows moved"
        # ownership: every key was updated by exactly its owner (rows of other shards untouched there)
        for r in range(W):
            assert ranks[r].last_step["recv_counts"] == [ranks[s].last_step["send_counts"][r] for s in range(W)]
    for a, b in zip(*results):
        assert torch.equal(a, b), "sharded step must be bitwise reproducible"


@pytest.mark.parametrize("W", [1, 2, 4])
def test_prefetch_protocol_equals_per_call_protocol(W):
    """Step-level prefetch (one sort/dedup + one exchange; backward sends gradient rows only) must give the
    same forward rows bit-exactly and the same updated tables as the per-call protocol; W=1 == unsharded bitwise."""
    from tencent_recommendation_2025_b200.sharded import run_emulated, tables_from_shards
    outs_tabs = []
    for prefetch in (False, True):
        cfg, full, lay, ranks, steps, pbs = setup(W)
        if prefetch:
            if W == 2:   # look-ahead form: phases issued separately
                run_emulated([ranks[r].prepare_gen(pbs[r]) for r in range(W)])
                run_emulated([ranks[r].finish_prepare_gen() for r in range(W)])
            run_emulated([ranks[r].prefetch_gen(pbs[r]) for r in range(W)])
        fwd = []
        for c in range(3):
            o = run_emulated([ranks[r].forward_gen(pbs[r][c]) for r in range(W)])
            for r in range(W):
                ref_item, ref_user = full.engine.forward(pbs[r][c])
                assert torch.equal(o[r][0], ref_item)
                if ref_user is not None:
                    assert torch.equal(o[r][1], ref_user)
        dc = [_dcats(lay, steps[r], 100 + r) for r in range(W)]
        for r in range(W):
            for pb, (di, du) in reversed(list(zip(pbs[r], dc[r]))):      # autograd queues the later calls first
                ranks[r].queue(pb, di.cuda(), None if du is None else du.cuda())
        run_emulated([ranks[r].step_gen(dict(HYPER)) for r in range(W)])
        outs_tabs.append(tables_from_shards(lay, [rk.ops.local for rk in ranks]))
        if W == 1 and prefetch:
            for pb, (di, du) in zip(pbs[0], dc[0]):
                full.engine.queue(pb, di.cuda(), None if du is None else du.cuda())
            full.fused_step(**HYPER)
            for g, p in zip(outs_tabs[-1], full.engine.tables):
                assert torch.equal(g, p.data)
    for a, b in zip(*outs_tabs):
        # per-call protocol sorts the queued calls in backward order, prefetch in forward order: the per-row
        # sums associate differently, so compare like any two fp32 reductions (plus Adam's sign-of-zero caveat)
        d = (a - b).abs()
        assert float(d.max()) <= 2.2e-3 and float((d > 1e-5 * float(b.abs().max())).float().mean()) < 2e-4
