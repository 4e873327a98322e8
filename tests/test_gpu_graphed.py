"""The step as a CUDA graph (graphed.GraphedStep) against the eager step, bit for bit, on batches of DIFFERENT content and
lookup counts; plus the device-count variants of sort / dedup / remap / reduce (``n_is_capacity``) on their own."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld  # noqa: E402

pytestmark = pytest.mark.gpu

HYPER = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)


def _module(cfg, dev):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    args = types.SimpleNamespace(device=str(dev), hidden_units=cfg.H)
    torch.manual_seed(0)
    with torch.device(dev):
        m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, mode="fused", path="factored")
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0.0, 0.05 if p.dim() >= 2 else 0.1, generator=g)
        for p in m.engine.tables:
            p[0].zero_()
    return m


def _setup(B=24, L=29, H=64, n_steps=5):
    from tencent_recommendation_2025_b200.resident import CallShape, ResidentItemFeatures
    dev = torch.device("cuda", 0)
    cfg = SynthConfig(B=B, L=L, H=H, item_num=3000, user_num=400, alpha=1.1, mm_ids=("81",), min_len=3)
    world = SynthWorld(cfg, 3)
    steps = [world.make_step(s) for s in range(n_steps)]
    store = ResidentItemFeatures.from_world(world, dev)
    shapes = [CallShape.covering([st.calls[i] for st in steps]) for i in range(3)]
    return dev, cfg, world, steps, store, shapes


def _make_body(m, opt, ups):
    def body(pbs):
        opt.zero_grad(set_to_none=True)
        m.prefetch(pbs)
        outs = [m.feat2emb_packed(pb) for pb in pbs]
        torch.autograd.backward(outs, ups)
        opt.step()
        m.fused_step(**HYPER)
        return [o.detach() for o in outs]
    return body


def test_fixed_shape_calls_are_the_same_calls():
    """Padding to a CallShape changes nothing but buffer sizes: same outputs, same updated tables as the exact-shape calls
    (this also runs every kernel in its device-count mode against its host-count mode)."""
    dev, cfg, world, steps, store, shapes = _setup()
    ref, fix = _module(cfg, dev), _module(cfg, dev)      # same seed: identical parameters
    for st in steps[:3]:
        ups = [torch.from_numpy(r).to(dev) for r in st.upstream]
        outs = {}
        for name, m, shp in (("exact", ref, [None] * 3), ("fixed", fix, shapes)):
            slim = store.slim_step(st.calls, shp)
            dints = slim.ints.to(dev)
            pbs = [store.expand(sc, dints[b:b + sc.ints.numel()]) for sc, b in zip(slim.calls, slim.bases)]
            assert (pbs[0].n_cap is None) == (name == "exact")
            m.prefetch(pbs)
            o = [m.feat2emb_packed(pb) for pb in pbs]
            torch.autograd.backward(o, ups)
            m.fused_step(**HYPER)
            outs[name] = o
        for a, b in zip(outs["exact"], outs["fixed"]):
            assert torch.equal(a, b)
    for (k, p), (_, q) in zip(ref.named_parameters(), fix.named_parameters()):
        assert torch.equal(p, q), k
        if p.grad is not None:
            assert torch.equal(p.grad, q.grad), k


def test_graphed_step_equals_eager_step_bitwise():
    from tencent_recommendation_2025_b200.graphed import GraphedStep
    dev, cfg, world, steps, store, shapes = _setup()
    eager, graphed = _module(cfg, dev), _module(cfg, dev)
    ups = [torch.from_numpy(r).to(dev) for r in steps[0].upstream]        # static upstream gradients (the trunk's stand-in)
    slims = [store.slim_step(st.calls, shapes) for st in steps]
    assert len({st.n_valid for st in slims}) > 1, "the batches must differ in their lookup counts"

    def dense(m):
        return [p for p in m.parameters() if not any(p is t for t in m.engine.tables)]

    opt_e = torch.optim.AdamW(dense(eager), lr=1e-3, betas=(0.9, 0.98), fused=True, capturable=True)
    opt_g = torch.optim.AdamW(dense(graphed), lr=1e-3, betas=(0.9, 0.98), fused=True, capturable=True)
    body_e = _make_body(eager, opt_e, ups)
    n_warm = 2
    runner = GraphedStep(graphed, store, slims[0], _make_body(graphed, opt_g, ups), hyper=HYPER, warmup=n_warm)
    # the runner warmed up on slims[0] n_warm times: do the same eagerly
    d0 = slims[0].ints.to(dev)
    for _ in range(n_warm):
        body_e([store.expand(sc, d0[b:b + sc.ints.numel()]) for sc, b in zip(slims[0].calls, slims[0].bases)])
    assert eager.engine.step == graphed.engine.step == n_warm
    for k, st in enumerate(slims + slims[:2]):
        dints = st.ints.to(dev)
        want = body_e([store.expand(sc, dints[b:b + sc.ints.numel()]) for sc, b in zip(st.calls, st.bases)])
        if k % 2 == 0:
            runner.submit(st)            # through the staging slots / copy stream
        else:
            runner.load(dints)           # inputs already in HBM
        got = runner.run()
        for a, b in zip(got, want):
            assert torch.equal(a, b), f"step {k}: outputs differ"
    torch.cuda.synchronize()
    assert eager.engine.step == graphed.engine.step
    for (k, p), (_, q) in zip(eager.named_parameters(), graphed.named_parameters()):
        assert torch.equal(p, q), f"{k} differs after {runner.replays} replays"
    for t in range(len(eager.engine.tables)):
        for nm in ("exp_avg", "exp_avg_sq"):
            a, b = getattr(eager.engine, nm, None), getattr(graphed.engine, nm, None)
            if a is not None:
                assert torch.equal(a[t], b[t]), f"{nm}[{t}]"
    runner.close()


def test_submit_rejects_other_shapes():
    from tencent_recommendation_2025_b200.graphed import GraphedStep
    from tencent_recommendation_2025_b200.resident import CallShape
    dev, cfg, world, steps, store, shapes = _setup(n_steps=2)
    m = _module(cfg, dev)
    ups = [torch.from_numpy(r).to(dev) for r in steps[0].upstream]
    opt = torch.optim.AdamW([m.itemdnn.weight], lr=1e-3, fused=True, capturable=True)
    runner = GraphedStep(m, store, store.slim_step(steps[0].calls, shapes), _make_body(m, opt, ups), hyper=HYPER, warmup=1)
    other = [CallShape(s.B, s.L, s.include_user, s.n_user_cap, tuple(c + 1024 for c in s.arr_caps)) for s in shapes]
    if any(s.arr_caps for s in shapes):
        with pytest.raises(ValueError):
            runner.submit(store.slim_step(steps[1].calls, other))
    with pytest.raises(ValueError):
        GraphedStep(m, store, store.slim_step(steps[0].calls), _make_body(m, opt, ups))   # exact-shape calls cannot be captured
    runner.close()


def test_engine_owned_dense_adamw_matches_torch_and_is_graphable():
    """own_dense_parameters(): itemdnn / userdnn / emb_transform updated by tgr_adam_dense from the kernels' accumulators ==
    torch.optim.AdamW on the autograd gradients (1e-5 of tensor scale over three steps), and the graphed step of that body ==
    its eager step bit for bit."""
    from tencent_recommendation_2025_b200.graphed import GraphedStep
    dev, cfg, world, steps, store, shapes = _setup()
    tref, own, gown = _module(cfg, dev), _module(cfg, dev), _module(cfg, dev)
    own.own_dense_parameters()
    gown.own_dense_parameters()
    ups = [torch.from_numpy(r).to(dev) for r in steps[0].upstream]
    dense = [p for p in tref.parameters() if not any(p is t for t in tref.engine.tables)]
    opt = torch.optim.AdamW(dense, lr=HYPER["lr"], betas=HYPER["betas"], eps=HYPER["eps"], weight_decay=HYPER["weight_decay"])
    slims = [store.slim_step(st.calls, shapes) for st in steps]

    def body_own(m):
        def body(pbs):
            m.prefetch(pbs)
            outs = [m.feat2emb_packed(pb) for pb in pbs]
            torch.autograd.backward(outs, ups)
            m.fused_step(**HYPER, dense=True)
            return [o.detach() for o in outs]
        return body

    runner = GraphedStep(gown, store, slims[0], body_own(gown), hyper=HYPER, warmup=1)
    b_ref, b_own = _make_body(tref, opt, ups), body_own(own)
    for k, st in enumerate([slims[0]] + slims[:3]):
        dints = st.ints.to(dev)
        pbs = lambda: [store.expand(sc, dints[b:b + sc.ints.numel()]) for sc, b in zip(st.calls, st.bases)]   # noqa: E731
        o_ref, o_own = b_ref(pbs()), b_own(pbs())
        if k > 0:                       # the runner's own warm-up was step 0
            runner.load(dints)
            o_g = runner.run()
            for a, b in zip(o_g, o_own):
                assert torch.equal(a, b), f"step {k}: graphed != eager"
        for a, b in zip(o_own, o_ref):
            assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-30), f"step {k}"
    for (k, p), (_, q), (_, r) in zip(tref.named_parameters(), own.named_parameters(), gown.named_parameters()):
        assert torch.equal(q, r), f"{k}: graphed != eager"
        assert q.grad is None or any(q is t for t in own.engine.tables) or True
        assert (p - q).abs().max().item() <= 1e-5 * max(p.abs().max().item(), 1e-30), k
    for p in own.dense_parameters():
        assert p.grad is None, "own_dense: the Linear gradients stay inside the engine"
    runner.close()


def test_pipelined_step_equals_eager_step_bitwise():
    """PipelinedStep: replay r computes batch r and, on a forked branch of the same graph, runs the expansion / key
    processing of batch r + 1. Outputs and every parameter after 8 replays == the eager steps, bit for bit."""
    from tencent_recommendation_2025_b200.graphed import PipelinedStep
    dev, cfg, world, steps, store, shapes = _setup()
    eager, piped = _module(cfg, dev), _module(cfg, dev)
    eager.own_dense_parameters()
    piped.own_dense_parameters()
    ups = [torch.from_numpy(r).to(dev) for r in steps[0].upstream]
    slims = [store.slim_step(st.calls, shapes) for st in steps]

    def body_of(m):
        def body(pbs):
            m.prefetch(pbs)
            outs = [m.feat2emb_packed(pb) for pb in pbs]
            torch.autograd.backward(outs, ups)
            m.fused_step(**HYPER, dense=True)
            return [o.detach() for o in outs]
        return body

    n_warm = 2
    runner = PipelinedStep(piped, store, slims[0], body_of(piped), hyper=HYPER, warmup=n_warm)
    b_e = body_of(eager)

    def eager_step(st):
        dints = st.ints.to(dev)
        return b_e([store.expand(sc, dints[b:b + sc.ints.numel()]) for sc, b in zip(st.calls, st.bases)])

    for _ in range(n_warm):
        eager_step(slims[0])
    assert eager.engine.step == piped.engine.step == n_warm
    seq = slims[1:] + slims + slims[:2]
    runner.prime(seq[0])
    for k, st in enumerate(seq):
        if k + 1 < len(seq):
            if k % 2 == 0:
                runner.submit(seq[k + 1])
            else:
                runner.load(seq[k + 1].ints.to(dev))
        got = runner.run()
        want = eager_step(st)
        for a, b in zip(got, want):
            assert torch.equal(a, b), f"step {k}: outputs differ"
    torch.cuda.synchronize()
    assert eager.engine.step == piped.engine.step
    for (k, p), (_, q) in zip(eager.named_parameters(), piped.named_parameters()):
        assert torch.equal(p, q), f"{k} differs after {runner.replays} replays"
    runner.close()
