"""world_size-2 gloo run (CPU) of the row-sharded exchange: the same generators the NCCL path drives, with the
numpy ShardOps standing in for the CUDA kernels. W=2 must reproduce the W=1 result: forward rows bit-exact,
routing facts exact, updated rows within fp32 re-association."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shard_numpy_ops import NumpyPeerShardOps, NumpyShardOps
from tencent_recommendation_2025_b200.packed import to_device
from tencent_recommendation_2025_b200.sharded import (ShardedRank, run_distributed, run_emulated, shard_of_tables,
                                                      tables_from_shards)
from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld

STATS = {"103": 7, "104": 20, "105": 50, "109": 90, "100": 10, "117": 40, "111": 90, "118": 150, "101": 300,
         "102": 12, "119": 33, "120": 77, "114": 120, "112": 260, "121": 9, "115": 45, "122": 85, "116": 140,
         "106": 60, "107": 110, "108": 200, "110": 30}
HYPER = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)


def assert_adam_close(a, b, lr=1e-3, what=""):
    """AdamW step 1 moves every element by ~lr*sign(g): an element whose summed gradient is within rounding of
    zero may legitimately land 2*lr apart under a different fp32 association. Everything else must agree."""
    d = (a - b).abs()
    tol = 2e-6 * float(b.abs().max())
    assert float(d.max()) <= 2.2 * lr, what
    assert float((d > tol).float().mean()) < 2e-4, what


def make_world():
    cfg = SynthConfig(B=4, L=12, H=32, item_num=400, user_num=40, alpha=1.2, mm_ids=("81",), min_len=3,
                      feat_statistics=STATS)
    world = SynthWorld(cfg, 11)
    lay = world.layout
    g = torch.Generator().manual_seed(5)
    tables = [0.1 * torch.randn((t.rows, lay.H), generator=g) for t in lay.tables]
    for t in tables:
        t[0] = 0
    mm = {k: (0.1 * torch.randn((lay.H, d), generator=g).numpy(), 0.1 * torch.randn(lay.H, generator=g).numpy())
          for k, d in lay.item_emb_feat.items()}
    return cfg, world, lay, tables, mm


def dcats(lay, st, seed):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for pc in st.calls:
        cl = lay.calls[pc.include_user]
        out.append((torch.randn((pc.T, cl.item_dim), generator=gen),
                    torch.randn((pc.T, cl.user_dim), generator=gen) if pc.include_user else None))
    return out


def rank_job(rank, W, lay, world, tables, mm, run, prefetch=False):
    ops = NumpyShardOps(lay, shard_of_tables(tables, rank, W), mm, W)
    rk = ShardedRank(lay, ops, rank, W)
    st = world.make_step(rank)
    pbs = [to_device(lay, pc, "cpu", pin=False) for pc in st.calls]
    if prefetch:
        # look-ahead form: phases A and B issued separately (as one step ahead), then phase C
        run(rk.prepare_gen(pbs))
        run(rk.finish_prepare_gen())
        run(rk.prefetch_gen(pbs))
    outs = [run(rk.forward_gen(pb)) for pb in pbs]
    facts = [dict(rk.last_fwd)]
    for pb, (di, du) in zip(pbs, dcats(lay, st, 50 + rank)):
        rk.queue(pb, di, du)
    run(rk.step_gen(dict(HYPER)))
    facts.append(dict(rk.last_step))
    return outs, ops.local, facts


def _worker(rank, W, port, tmp, prefetch=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=W)
    torch.set_num_threads(1)
    cfg, world, lay, tables, mm = make_world()
    outs, local, facts = rank_job(rank, W, lay, world, tables, mm, lambda g: run_distributed(g), prefetch)

    def extra():   # the request kinds only the peer-memory protocol issues, through the real process group
        m = yield ("allgather", torch.tensor([rank, 10 + rank], dtype=torch.int32))
        yield ("barrier", torch.ones(1))
        return m

    m = run_distributed(extra())
    assert m.tolist() == [[r, 10 + r] for r in range(W)]
    torch.save({"outs": outs, "local": local, "facts": facts}, os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("prefetch", [False, True])
def test_world_size_2_gloo_matches_single_rank(prefetch):
    """prefetch=True: the step-level protocol (one exchange for all calls; backward sends gradient rows only)."""
    W = 2
    port = 29500 + (os.getpid() % 400) + (17 if prefetch else 0)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(W, port, tmp, prefetch), nprocs=W, join=True)
        res = [torch.load(os.path.join(tmp, f"r{r}.pt"), weights_only=False) for r in range(W)]
    cfg, world, lay, tables, mm = make_world()
    # single-rank reference: one rank owns everything and processes both data-parallel shares
    ops1 = NumpyShardOps(lay, shard_of_tables(tables, 0, 1), mm, 1)
    rk1 = ShardedRank(lay, ops1, 0, 1)
    for r in range(W):
        st = world.make_step(r)
        pbs = [to_device(lay, pc, "cpu", pin=False) for pc in st.calls]
        for c, pb in enumerate(pbs):
            ref = run_emulated([rk1.forward_gen(pb)])[0]
            got = res[r]["outs"][c]
            assert torch.equal(got[0], ref[0]), f"rank {r} call {c}: forward rows must be bit-exact"
            if ref[1] is not None:
                assert torch.equal(got[1], ref[1])
        for pb, (di, du) in zip(pbs, dcats(lay, st, 50 + r)):
            rk1.queue(pb, di, du)
    run_emulated([rk1.step_gen(dict(HYPER))])
    ref_tables = tables_from_shards(lay, [ops1.local])
    got_tables = tables_from_shards(lay, [res[r]["local"] for r in range(W)])
    moved = 0
    for a, b, t0 in zip(got_tables, ref_tables, tables):
        assert_adam_close(a, b, what="updated rows differ")
        moved += int((b != t0).any(dim=1).sum())
    assert moved > 50
    # routing facts: what r receives from s is what s sent to r, in both exchanges
    for phase in (0, 1):
        for r in range(W):
            assert res[r]["facts"][phase]["recv_counts"] == [res[s]["facts"][phase]["send_counts"][r] for s in range(W)]
            assert sum(res[r]["facts"][phase]["send_counts"]) == res[r]["facts"][phase]["U"]
    if prefetch:   # the backward reuses the forward's routing verbatim
        for r in range(W):
            assert res[r]["facts"][0] == res[r]["facts"][1]


def test_emulated_w4_equals_w1_numpy():
    """Same check without a process group: 4 emulated ranks (lockstep generators) vs 1."""
    cfg, world, lay, tables, mm = make_world()
    W = 4
    ranks = [ShardedRank(lay, NumpyShardOps(lay, shard_of_tables(tables, r, W), mm, W), r, W) for r in range(W)]
    rk1 = ShardedRank(lay, NumpyShardOps(lay, shard_of_tables(tables, 0, 1), mm, 1), 0, 1)
    steps = [world.make_step(r) for r in range(W)]
    pbs = [[to_device(lay, pc, "cpu", pin=False) for pc in st.calls] for st in steps]
    for c in range(3):
        outs = run_emulated([ranks[r].forward_gen(pbs[r][c]) for r in range(W)])
        for r in range(W):
            ref = run_emulated([rk1.forward_gen(pbs[r][c])])[0]
            assert torch.equal(outs[r][0], ref[0])
    for r in range(W):
        for pb, (di, du) in zip(pbs[r], dcats(lay, steps[r], 50 + r)):
            ranks[r].queue(pb, di, du)
            rk1.queue(pb, di, du)
    run_emulated([rk.step_gen(dict(HYPER)) for rk in ranks])
    run_emulated([rk1.step_gen(dict(HYPER))])
    a = tables_from_shards(lay, [rk.ops.local for rk in ranks])
    b = tables_from_shards(lay, [rk1.ops.local])
    for x, y in zip(a, b):
        assert_adam_close(x, y)
    # the first moment is linear in the summed gradient: a strict check of the reduction itself
    m4 = tables_from_shards(lay, [torch.from_numpy(rk.ops.m) for rk in ranks])
    m1 = tables_from_shards(lay, [torch.from_numpy(rk1.ops.m)])
    for x, y in zip(m4, m1):
        assert torch.allclose(x, y, rtol=0, atol=1e-6 * max(float(y.abs().max()), 1e-30))


@pytest.mark.parametrize("window_rows", [4096, 1])
def test_emulated_peer_memory_protocol_equals_exchange_protocol(window_rows):
    """The peer-memory protocol (rows read in place from the owners, gradient rows pulled from the sources' windows,
    W x W count matrix by all-gather) gives bit-identical forward rows and updated tables to the all-to-all protocol;
    a window that is too small falls back to the gradient all-to-all on every rank alike."""
    cfg, world, lay, tables, mm = make_world()
    W = 4
    steps = [world.make_step(r) for r in range(W)]
    results = []
    for peer in (False, True):
        if peer:
            ops = [NumpyPeerShardOps(lay, shard_of_tables(tables, r, W), mm, W, window_rows) for r in range(W)]
            for o in ops:
                o.link(ops)
        else:
            ops = [NumpyShardOps(lay, shard_of_tables(tables, r, W), mm, W) for r in range(W)]
        ranks = [ShardedRank(lay, ops[r], r, W) for r in range(W)]
        pbs = [[to_device(lay, pc, "cpu", pin=False) for pc in st.calls] for st in steps]
        run_emulated([ranks[r].prepare_gen(pbs[r]) for r in range(W)])
        run_emulated([ranks[r].finish_prepare_gen() for r in range(W)])
        run_emulated([ranks[r].prefetch_gen(pbs[r]) for r in range(W)])
        outs = [run_emulated([ranks[r].forward_gen(pbs[r][c]) for r in range(W)]) for c in range(3)]
        for r in range(W):
            for pb, (di, du) in zip(pbs[r], dcats(lay, steps[r], 50 + r)):
                ranks[r].queue(pb, di, du)
        run_emulated([rk.step_gen(dict(HYPER)) for rk in ranks])
        results.append((outs, tables_from_shards(lay, [o.local for o in ops])))
        if peer:
            assert all(o.pulled == (1 if window_rows > 1 else 0) for o in ops)
    for c in range(3):
        for r in range(W):
            assert torch.equal(results[0][0][c][r][0], results[1][0][c][r][0])
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.equal(a, b)


def test_shard_layout_round_trip():
    cfg, world, lay, tables, mm = make_world()
    for W in (1, 2, 3, 8):
        shards = [shard_of_tables(tables, r, W) for r in range(W)]
        back = tables_from_shards(lay, shards)
        for a, b in zip(back, tables):
            assert torch.equal(a, b)
