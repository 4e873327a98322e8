"""Row-sharded FACTORED path on ONE GPU with W emulated ranks (run_emulated): the factored kernels working on rows
fetched from their owners + the unchanged exchange protocol. W=1 equals the unsharded factored step bitwise; forward
outputs equal the unsharded factored module bitwise for every W (same unique list, same rows, same kernels); W>1 updated
rows / AdamW state and the summed dense-parameter gradients match the oracle (reference op sequence in numpy)."""
import types

import numpy as np
import pytest
import torch

from oracle import feat2emb_numpy as onp
from golden_util import assert_rows_updated
from test_gpu_sharded import HYPER, STATS
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld

pytestmark = pytest.mark.gpu


def setup(W, B=16, L=33, H=64, seed=3, p2p=False):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.packed import to_device
    from tencent_recommendation_2025_b200.sharded import FactShardOps, ShardedRank, shard_of_tables
    cfg = SynthConfig(B=B, L=L, H=H, item_num=5000, user_num=300, alpha=1.2, mm_ids=("81",), min_len=5,
                      feat_statistics=STATS)
    world = SynthWorld(cfg, seed)
    args = types.SimpleNamespace(device="cuda", hidden_units=H)
    torch.manual_seed(seed)
    full = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, "fused",
                             path="factored").cuda()
    with torch.no_grad():
        for p in full.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
        for p in full.engine.tables:
            p.normal_(0, 0.1)
            p[0].zero_()
    lay = full.layout
    tables = [p.data for p in full.engine.tables]
    dnn = {"item": full.itemdnn, "user": full.userdnn}
    ranks = []
    for r in range(W):
        ops = FactShardOps(lay, shard_of_tables(tables, r, W), dict(full.emb_transform.items()), dnn, W)
        ranks.append(ShardedRank(lay, ops, r, W))
    if p2p:   # emulated ranks share one address space: every rank reads the others' shards / gradient windows in place
        ptrs = [rk.ops.local.data_ptr() for rk in ranks]
        wins = [torch.zeros((1 << 14, H), device="cuda") for _ in ranks]
        for rk, w_ in zip(ranks, wins):
            rk.ops.peers = list(ptrs)
            rk.ops.grad_win = w_
            rk.ops.grad_peers = [x.data_ptr() for x in wins]
        from tencent_recommendation_2025_b200.sharded import emulate_io
        emulate_io(ranks, 1 << 16)      # counts / ids through the peer-memory mailboxes (tgr_peer_put / _pull / merge)
    steps = [world.make_step(r) for r in range(W)]                       # rank r's data-parallel share
    pbs = [[to_device(lay, pc, "cuda") for pc in st.calls] for st in steps]
    return cfg, full, lay, ranks, steps, pbs


def run_step(W, ranks, pbs, steps):
    """prefetch -> 3 forwards -> 3 backwards (later calls first, as autograd) -> exchange + owner AdamW."""
    from tencent_recommendation_2025_b200.sharded import run_emulated
    run_emulated([ranks[r].prefetch_gen(pbs[r]) for r in range(W)])
    outs = [[None] * 3 for _ in range(W)]
    for c in range(3):
        o = run_emulated([ranks[r].forward_gen(pbs[r][c]) for r in range(W)])
        for r in range(W):
            outs[r][c] = o[r]
    accs = []
    for r in range(W):
        g = ranks[r].pf["pf"]["group"]
        g.n_fwd = 3
        for c in (2, 1, 0):
            up = torch.from_numpy(steps[r].upstream[c]).cuda()
            done = ranks[r].ops.feng.fact_backward(g, pbs[r][c], up)
            ranks[r].queue(pbs[r][c], None, None)
        assert done
        accs.append(g.acc)
    run_emulated([ranks[r].step_gen(dict(HYPER)) for r in range(W)])
    return outs, accs


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("W", [1, 2, 4, 8])
def test_forward_equals_unsharded_factored_bitwise(W, p2p):
    from tencent_recommendation_2025_b200.sharded import run_emulated
    cfg, full, lay, ranks, steps, pbs = setup(W, p2p=p2p)
    run_emulated([ranks[r].prefetch_gen(pbs[r]) for r in range(W)])
    for r in range(W):
        with torch.no_grad():
            full.prefetch(pbs[r])
            ref = [full.feat2emb_packed(pb) for pb in pbs[r]]
        for c in range(3):
            out = run_emulated([ranks[q].forward_gen(pbs[q][c]) for q in range(W)])[r]
            assert torch.equal(out.view_as(ref[c]), ref[c]), f"W={W} rank {r} call {c}"


@pytest.mark.parametrize("p2p", [False, True])
def test_w1_step_equals_unsharded_factored_bitwise(p2p):
    from tencent_recommendation_2025_b200.sharded import tables_from_shards
    cfg, full, lay, ranks, steps, pbs = setup(1, p2p=p2p)
    dense = [p for p in full.dense_parameters()]
    full.prefetch(pbs[0])
    outs = [full.feat2emb_packed(pb) for pb in pbs[0]]
    torch.autograd.backward(outs, [torch.from_numpy(u).cuda() for u in steps[0].upstream])
    ref_dense = {k: p.grad.clone() for k, p in full.named_parameters() if p.grad is not None}
    souts, accs = run_step(1, ranks, pbs, steps)          # reads the tables' values before the unsharded update
    full.fused_step(**HYPER)
    got = tables_from_shards(lay, [ranks[0].ops.local])
    for g, p, t in zip(got, full.engine.tables, lay.tables):
        assert torch.equal(g, p.data), t.name
    for c in range(3):
        assert torch.equal(souts[0][c].view_as(outs[c]), outs[c])
    a = accs[0]
    assert torch.equal(a["dW_item"], ref_dense["itemdnn.weight"]) and torch.equal(a["db_item"], ref_dense["itemdnn.bias"])
    assert torch.equal(a["dW_user"], ref_dense["userdnn.weight"]) and torch.equal(a["db_user"], ref_dense["userdnn.bias"])
    assert torch.equal(a["dWmm/81"], ref_dense["emb_transform.81.weight"])


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("W", [2, 4])
def test_sharded_factored_step_matches_oracle_and_is_deterministic(W, p2p):
    from tencent_recommendation_2025_b200.sharded import tables_from_shards
    results = []
    for rep in range(2):
        cfg, full, lay, ranks, steps, pbs = setup(W, p2p=p2p)
        params = {k: v.detach().cpu().numpy().copy() for k, v in full.named_parameters()}
        outs, accs = run_step(W, ranks, pbs, steps)
        got = tables_from_shards(lay, [rk.ops.local for rk in ranks])
        results.append(got + [sum(a[k] for a in accs) for k in ("dW_item", "dW_user", "db_item", "db_user", "dWmm/81", "dbmm/81")])
        if rep:
            continue
        # oracle: the reference op sequence per rank, gradients summed over ranks, one AdamW row update
        tot = None
        for r in range(W):
            for c, pc in enumerate(steps[r].calls):
                ref, cache = onp.feat2emb_forward(params, lay, pc.seq, onp.tensors_from_packed(lay, pc), pc.mask, pc.include_user)
                o = outs[r][c].cpu().numpy().reshape(ref.shape)
                assert np.abs(o - ref).max() <= 1e-5 * np.abs(ref).max(), f"rank {r} call {c} forward"
                g = onp.feat2emb_backward(params, lay, cache, steps[r].upstream[c])
                tot = g if tot is None else {k: tot[k] + g[k] if k in g else tot[k] for k in set(tot) | set(g)} | {k: g[k] for k in g if k not in tot}
        m_tabs = tables_from_shards(lay, [rk.ops.exp_avg for rk in ranks])
        for ti, (t, a) in enumerate(zip(lay.tables, got)):
            k = f"{t.name}.weight"
            g = tot[k]
            rows_t = np.nonzero(np.any(g != 0, axis=1))[0]
            w = params[k].copy()
            m, v = np.zeros_like(w), np.zeros_like(w)
            onp.adamw_rows(w, m, v, rows_t, g[rows_t], 1, lr=1e-3, wd=1e-2)
            an = a.cpu().numpy()
            assert_rows_updated(an[rows_t], w[rows_t], g[rows_t], 1e-3, what=t.name)
            untouched = np.setdiff1d(np.arange(w.shape[0]), rows_t)
            assert np.array_equal(an[untouched], params[k][untouched]), t.name + ": untouched rows moved"
            mg = m_tabs[ti].cpu().numpy()
            assert np.abs(mg - m).max() <= 1e-5 * max(np.abs(m).max(), 1e-30), t.name + " exp_avg"
        for key, name in (("dW_item", "itemdnn.weight"), ("dW_user", "userdnn.weight"), ("db_item", "itemdnn.bias"),
                          ("db_user", "userdnn.bias"), ("dWmm/81", "emb_transform.81.weight"), ("dbmm/81", "emb_transform.81.bias")):
            s = sum(a[key] for a in accs).cpu().numpy()
            assert np.abs(s - tot[name]).max() <= 1e-5 * np.abs(tot[name]).max(), name
    for a, b in zip(*results):
        assert torch.equal(a, b), "sharded factored step must be bitwise reproducible"


def test_peer_source_projection_equals_table_source():
    """tgr_row_source_t.peer_rows (rows read in place from W shards inside the projection kernel) and
    tgr_fetch_peer_rows + fetched_rows give the table-source projection bit for bit."""
    import ctypes as C
    from tencent_recommendation_2025_b200 import _lib
    from tencent_recommendation_2025_b200.engine import _stream
    from tencent_recommendation_2025_b200.sharded import shard_of_tables
    W = 4
    cfg, full, lay, ranks, steps, pbs = setup(1)
    eng = full.engine
    g = eng.prepare(pbs[0])
    lib, H, nt = eng.lib, lay.H, len(eng.tables)
    cap = int(g.c.cap)
    shards = [shard_of_tables([p.data for p in eng.tables], r, W) for r in range(W)]
    outs = []
    for mode in ("table", "peer", "fetched"):
        P = torch.zeros((cap, H), device="cuda")
        src = _lib.RowSource()
        keep = None
        if mode == "peer":
            src.n_peers = W
            for r in range(W):
                src.peer_rows[r] = shards[r].data_ptr()
        elif mode == "fetched":
            keep = torch.zeros((cap, H), device="cuda")
            ptrs = (C.c_void_p * W)(*[s.data_ptr() for s in shards])
            _lib.check(lib.tgr_fetch_peer_rows(ptrs, W, H, g.c.uniq, g.c.n_unique, cap, keep.data_ptr(), _stream()))
            src.fetched_rows = keep.data_ptr()
        _lib.check(lib.tgr_fact_project_rows(eng._table_array(), nt, H, C.byref(eng._params().dnn), g.c.uniq, g.c.n_unique, cap,
                                             C.byref(src), P.data_ptr(), _stream()))
        torch.cuda.synchronize()
        outs.append(P[: int(g.n_unique.item())].clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
