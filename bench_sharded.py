"""Multi-GPU leg of bench.py: row-sharded tables, one process per GPU (torchrun), NCCL all-to-all of deduplicated
ids / rows / gradient rows. Weak scaling: every rank runs B sequences per step against the
SAME global tables as the 1-GPU line, now split W ways (owner = key mod W)."""
from __future__ import annotations

import json
import os
import time
import types

import numpy as np
import torch
import torch.distributed as dist

import bench
from tencent_recommendation_2025_b200 import synth
from tencent_recommendation_2025_b200.packed import to_device


def live_parity_check(args, rank: int, world: int, dev) -> dict:
    """W-rank result == emulated-W result, checked on the wires the benchmark times (the NCCL +
    symmetric-memory path had no value check). Every rank builds the SAME small problem (all ranks' batches, identical
    parameters), runs the W-rank step twice — (a) with W emulated ranks inside its own process (`run_emulated`: the
    configuration the single-GPU tests pin to the oracle and to the unsharded step, tests/test_gpu_sharded_factored.py) and
    (b) as one rank of the live process group (NCCL collectives, peer-memory row reads, gradient windows) — and compares
    its forward outputs, its dense-parameter gradients, its updated table shard and AdamW state BIT FOR BIT. The
    arithmetic is deterministic (fixed reduction orders), so anything but equality is a transport / ordering bug."""
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    from tencent_recommendation_2025_b200.sharded import (FactShardOps, ShardedBaselineEmbedding, ShardedRank, run_emulated,
                                                            shard_of_tables)
    W = world
    stats = {k: 50 for k in ["103", "104", "105", "109", "100", "117", "111", "118", "101", "102", "119", "120", "114", "112",
                             "121", "115", "122", "116", "106", "107", "108", "110"]}
    cfg = synth.SynthConfig(B=16, L=33, H=64, item_num=5000, user_num=300, alpha=1.2, mm_ids=("81",), min_len=5,
                            feat_statistics=stats)
    gen = synth.SynthWorld(cfg, 3)
    steps = [gen.make_step(r) for r in range(W)]
    margs = types.SimpleNamespace(device=str(dev), hidden_units=cfg.H)
    torch.manual_seed(3)
    full = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), margs, "fused", path="factored").to(dev)
    g = torch.Generator(device=dev).manual_seed(17)
    with torch.no_grad():
        for p in full.parameters():
            p.normal_(0.0, 0.1, generator=g)
        for p in full.engine.tables:
            p[0].zero_()
    lay = full.layout
    tables = [p.data for p in full.engine.tables]
    hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    pbs_all = [[to_device(lay, pc, dev) for pc in st.calls] for st in steps]
    ups_all = [[torch.from_numpy(u).to(dev) for u in st.upstream] for st in steps]
    # ---- (a) emulated ranks in this process ----
    ranks = []
    for r in range(W):
        ops = FactShardOps(lay, shard_of_tables(tables, r, W), dict(full.emb_transform.items()),
                           {"item": full.itemdnn, "user": full.userdnn}, W)
        ranks.append(ShardedRank(lay, ops, r, W))
    if not args.no_p2p:
        wins = [torch.zeros((1 << 14, cfg.H), device=dev) for _ in ranks]
        for rk, w_ in zip(ranks, wins):
            rk.ops.peers = [q.ops.local.data_ptr() for q in ranks]
            rk.ops.grad_win = w_
            rk.ops.grad_peers = [x.data_ptr() for x in wins]
        from tencent_recommendation_2025_b200.sharded import emulate_io
        emulate_io(ranks, 1 << 16)
    run_emulated([ranks[r].prefetch_gen(pbs_all[r]) for r in range(W)])
    exp_out = []
    for c in range(3):
        exp_out.append(run_emulated([ranks[r].forward_gen(pbs_all[r][c]) for r in range(W)])[rank].clone())
    for r in range(W):
        grp = ranks[r].pf["pf"]["group"]
        grp.n_fwd = 3
        for c in (2, 1, 0):
            ranks[r].ops.feng.fact_backward(grp, pbs_all[r][c], ups_all[r][c])
            ranks[r].queue(pbs_all[r][c], None, None)
        if r == rank:
            exp_acc = {k: v.clone() for k, v in grp.acc.items() if isinstance(v, torch.Tensor)}
    run_emulated([ranks[r].step_gen(dict(hyper)) for r in range(W)])
    exp_shard, exp_m = ranks[rank].ops.local.clone(), ranks[rank].ops.exp_avg.clone()
    # ---- (b) this process as one rank of the live group ----
    m = ShardedBaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), margs, rank, W,
                                 path="factored", p2p=not args.no_p2p, grad_window_rows=1 << 14)
    m.load_full_tables(tables)
    with torch.no_grad():
        for k in lay.item_emb_feat:
            m.emb_transform[k].weight.copy_(full.emb_transform[k].weight)
            m.emb_transform[k].bias.copy_(full.emb_transform[k].bias)
        m.itemdnn.weight.copy_(full.itemdnn.weight); m.itemdnn.bias.copy_(full.itemdnn.bias)
        m.userdnn.weight.copy_(full.userdnn.weight); m.userdnn.bias.copy_(full.userdnn.bias)
    torch.cuda.synchronize()
    dist.barrier()
    pbs = pbs_all[rank]
    m.prefetch(pbs)
    outs = [m.feat2emb_packed(pb) for pb in pbs]
    torch.autograd.backward(outs, ups_all[rank])
    m.fused_step(**hyper)
    torch.cuda.synchronize()
    checks = {
        "forward": all(torch.equal(o.reshape(e.shape), e) for o, e in zip(outs, exp_out)),
        "dense_grads": (torch.equal(m.itemdnn.weight.grad, exp_acc["dW_item"]) and torch.equal(m.userdnn.weight.grad, exp_acc["dW_user"])
                        and torch.equal(m.emb_transform["81"].weight.grad, exp_acc["dWmm/81"])),
        "updated_shard": torch.equal(m.local_table.data, exp_shard),
        "adam_state": torch.equal(m.ops.exp_avg, exp_m),
    }
    flags = torch.tensor([1.0 if v else 0.0 for v in checks.values()], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out = {k: bool(f) for k, f in zip(checks, flags.tolist())}
    out["parity_ok"] = all(out.values())
    out["how"] = (f"W={W} live ranks (NCCL + {'peer memory' if getattr(m.ops, 'peers', None) is not None else 'all-to-all'}) vs "
                  f"{W} emulated ranks in one process, bit for bit, min over ranks; B=16 x L=33 per rank, 5000-item tables")
    del m, ranks, full
    torch.cuda.empty_cache()
    return out


def run_sharded(args, rank: int, world: int, local_rank: int):
    from tencent_recommendation_2025_b200.sharded import ShardedBaselineEmbedding
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = args.dnn_matmul == "tf32"
    hbm_peak, peak_src = bench.load_peaks()
    parity = live_parity_check(args, rank, world, dev) if args.path == "factored" else None
    if parity is not None and not parity["parity_ok"] and rank == 0:
        import sys
        print(f"WARNING: live sharded parity check failed: {parity}", file=sys.stderr)
    cfg = bench.get_config(args.config, args.batch)
    worldgen = synth.SynthWorld(cfg, 0)
    lay = worldgen.layout
    margs = types.SimpleNamespace(device=str(dev), hidden_units=cfg.H)
    n_batches = max(1, min(args.batches, args.steps + args.warmup))
    steps_np = [worldgen.make_step(1000 * rank + s) for s in range(n_batches)]
    # gradient window (symmetric memory, same size on every rank): half the largest step's lookup count bounds the
    # unique rows of this workload with a wide margin; a step that does not fit falls back to the NCCL all-to-all
    from tencent_recommendation_2025_b200.packed import count_valid
    n_max = torch.tensor([max(sum(count_valid(lay, pc) for pc in st.calls) for st in steps_np)], device=dev)
    dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
    win_rows = max(1 << 20, int(n_max.item()) // 2)
    torch.manual_seed(0)                                   # identical dense parameters on every rank
    m = ShardedBaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), margs, rank, world,
                                 path=args.path, p2p=not args.no_p2p, grad_window_rows=win_rows,
                                 max_step_entries=int(n_max.item()) + 1024, symm_io=not args.no_symm_io)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    with torch.no_grad():
        for p in m.parameters():
            if p is not m.local_table:
                p.normal_(0.0, 0.05)
        m.local_table.normal_(0.0, 0.05, generator=g)
        for t in lay.tables:                               # padding rows are zero wherever they live
            if t.key_base % world == rank:
                m.local_table[t.key_base // world].zero_()
    dense = [p for p in m.parameters() if p is not m.local_table]
    # replicated dense parameters: one tgr_adam_dense launch (m.dense_adam_) unless --torch-dense-opt
    dense_opt = torch.optim.AdamW(dense, lr=1e-3, betas=(0.9, 0.98), fused=True) if args.torch_dense_opt else None
    dev_steps = [([to_device(lay, pc, dev) for pc in st.calls], [torch.from_numpy(r).to(dev) for r in st.upstream])
                 for st in steps_np]
    hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    lookups = [st.n_lookups() for st in steps_np]   # host-side row counting stays out of the timed regions
    # replicated dense parameters: their .grad tensors are views of ONE flat buffer, so the data-parallel
    # all-reduce is a single collective with no flatten / copy-back kernels
    flat_grad = m.symm_empty(sum(p.numel() for p in dense))      # symmetric memory when available: one-shot pull all-reduce
    o = 0
    for p in dense:
        p.grad = flat_grad[o:o + p.numel()].view_as(p)
        o += p.numel()

    def one_step(pbs, ups, next_pbs=None):
        flat_grad.zero_()
        if not args.no_prefetch:
            m.prefetch(pbs)                      # one dedup + one exchange for the step's three calls
        if next_pbs is not None and not args.no_lookahead:
            # next step's key processing (value independent) is enqueued EARLY: its per-owner counts reach the host
            # long before finish_prepare needs them, so the step never waits on a device->host copy
            m.prepare_next(next_pbs)
        outs = [m.feat2emb_packed(pb) for pb in pbs]
        torch.autograd.backward(outs, ups)
        m.fused_step(**hyper)                    # row gradients to their owners, owners' AdamW row update
        m.allreduce_dense_(flat_grad, pre_barrier=False)   # replicated dense parameters: mean over ranks (peer memory or NCCL)
        if dense_opt is not None:
            dense_opt.step()
        else:
            m.dense_adam_(dense, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
        if next_pbs is not None and not args.no_lookahead:
            m.finish_prepare()
        return outs

    from tencent_recommendation_2025_b200 import _lib
    from tencent_recommendation_2025_b200.packed import HostPrefetcher, stage_pinned
    clocks = bench.ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    nw = max(args.warmup, 3 * n_batches)     # every distinct batch shape several times: the caching allocator has seen every size
    for i in range(nw):
        one_step(*dev_steps[i % n_batches], next_pbs=dev_steps[(i + 1) % n_batches][0])
    torch.cuda.synchronize()
    dist.barrier()
    if clocks:
        clocks.mark()

    def n_launch():
        return _lib.launch_count()   # counted inside the library at every launch site

    l0 = n_launch()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the sharded step is issued eagerly and every rank waits for the slowest at four barriers per step: keep the Python
    # garbage collector's pauses (thousands of short-lived ctypes / tensor objects per step) out of the loop
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()
    torch.cuda.synchronize()
    ev0.record()
    rows = 0
    for i in range(args.steps):
        k = (nw + i) % n_batches
        one_step(*dev_steps[k], next_pbs=dev_steps[(k + 1) % n_batches][0])   # the data loader is one batch ahead
        rows += lookups[k]
    ev1.record()
    torch.cuda.synchronize()
    gc.enable()
    dist.barrier()
    clk = clocks.stop() if clocks else None
    ms_local = ev0.elapsed_time(ev1)
    launches = n_launch() - l0
    t = torch.tensor([ms_local, float(rows)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms = float(tmax[0])
    total_rows = float(tsum[1])

    # ---- per-entry CUDA-event timing inside the library: a second pass over the same steps (all ranks in lockstep) ----
    used = [(nw + i) % n_batches for i in range(args.steps)]
    _lib.timing_enable(True)
    for k in used:
        one_step(*dev_steps[k], next_pbs=dev_steps[(k + 1) % n_batches][0])
    torch.cuda.synchronize()
    kern_ms = _lib.timing_collect()
    _lib.timing_enable(False)
    dist.barrier()

    # ---- e2e from pinned host buffers: copy stream two steps ahead (the look-ahead key processing needs step k+1's
    #      ids on the device during step k), loss of every step copied back and read on the host ----
    use_resident = args.resident_items == "on" or (args.resident_items == "auto" and args.config in ("c1", "c2"))
    feed_how = "packed calls (every token's feature ids and mm vectors cross PCIe)"
    if use_resident:
        # item-side features resident in every rank's HBM (they are functions of the item id): the host hands over ids + user
        # tokens only (3 MB instead of 62 MB per step and rank), the packed calls are rebuilt on the device (resident.py)
        from tencent_recommendation_2025_b200.resident import ResidentFeeder, ResidentItemFeatures
        try:
            store = ResidentItemFeatures.from_world(worldgen, dev)
            host_steps = [[store.slim(pc) for pc in st.calls] for st in steps_np]
            feeder = ResidentFeeder(store, slots=5)
            feed_how = "slim calls (ids + user tokens; item feature / mm tables resident in HBM, expanded on the device)"
        except Exception as exc:
            use_resident = False
            feed_how += f" [resident feed failed: {exc!r}]"
    if not use_resident:
        host_steps = [[stage_pinned(lay, pc) for pc in st.calls] for st in steps_np]   # as a pin_memory DataLoader would
        feeder = HostPrefetcher(dev, slots=5)
    e2e_steps = max(3, min(args.steps, 50))
    e2e_warm = 2 * n_batches + 1
    m.rank_state.prep = None
    loss_host = torch.zeros(2, dtype=torch.float32, pin_memory=True)
    loss_ev = [None, None]
    losses = []
    rows_e2e = 0
    h2d = 0
    feeder.submit(host_steps[0])
    feeder.submit(host_steps[1 % n_batches])
    cur = feeder.take()
    t0 = None
    for i in range(e2e_warm + e2e_steps):
        k = i % n_batches
        if i == e2e_warm:
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
        nxt = feeder.take()
        feeder.submit(host_steps[(i + 2) % n_batches])
        outs = one_step(cur, dev_steps[k][1], next_pbs=nxt)
        feeder.retire()
        with torch.no_grad():
            loss = torch.stack([o.detach()[0, -1] for o in outs]).sum()   # checksum of the step's outputs (as bench.py step_result)
        slot = i & 1
        if loss_ev[slot] is not None:
            loss_ev[slot].synchronize()
            losses.append(float(loss_host[slot]))
        loss_host[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)
        loss_ev[slot] = torch.cuda.Event()
        loss_ev[slot].record()
        if i >= e2e_warm:
            rows_e2e += lookups[k]
            h2d += sum(pb.h2d_bytes for pb in cur)
        cur = nxt
    for ev in loss_ev:
        if ev is not None:
            ev.synchronize()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e, float(rows_e2e)], dtype=torch.float64, device=dev)
    te_max = te.clone()
    dist.all_reduce(te_max, op=dist.ReduceOp.MAX)
    te_sum = te.clone()
    dist.all_reduce(te_sum, op=dist.ReduceOp.SUM)

    if rank == 0:
        # rank 0's kernels against their compulsory bytes (rank 0's batches)
        kbytes = {}
        for k in used:
            st = steps_np[k]
            pu, su = bench.unique_counts(lay, st.calls)
            for name, v in bench.kernel_bytes_step(lay, st, dev_steps[k][0], pu, su, args.path).items():
                kbytes[name] = kbytes.get(name, 0) + v
        if args.path == "factored":
            kbytes.pop("adam_rows", None)          # the owner-side update is bwd_reduce (mode 1) over received rows
            kbytes.pop("bwd_reduce", None)         # two different launches share this entry name on the sharded path
        table = {}
        for name, (t_ms, cnt) in kern_ms.items():
            e = {"ms_per_step": round(t_ms / args.steps, 4), "launches_per_step": round(cnt / args.steps, 2)}
            if name in kbytes and t_ms > 0:
                gbs = kbytes[name] / (t_ms * 1e-3) / 1e9
                e.update({"alg_MB_per_step": round(kbytes[name] / args.steps / 1e6, 1), "GBs": round(gbs, 1),
                          "frac": round(gbs / hbm_peak, 4)})
            table[name] = e
        cand = [n_ for n_ in table if "GBs" in table[n_]]
        dom = max(cand, key=lambda n_: table[n_]["ms_per_step"]) if cand else None
        roofline = {"bound": "hbm", "kernel": (dom or "") + " (rank 0)", "achieved": table[dom]["GBs"] if dom else 0.0,
                    "peak": hbm_peak, "unit": "GB/s", "frac": table[dom]["frac"] if dom else 0.0,
                    "traffic": bench.load_traffic(dom, table[dom]["launches_per_step"] if dom else 1), "peak_source": peak_src, "kernels": table,
                    "kernels_ms_per_step_sum": round(sum(v["ms_per_step"] for v in table.values()), 4)}
        line = {"metric": bench.METRIC, "value": total_rows / (ms * 1e-3), "unit": bench.UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": bench.WORKLOADS[args.config] + f", tables row-sharded over {world} GPUs "
                           "(owner = key mod W), dedup-then-all-to-all of ids/rows/grad rows over NCCL",
                           "batch_per_gpu": cfg.B, "seq_len": cfg.L, "hidden": cfg.H, "item_rows": cfg.item_num + 1,
                           "user_rows": cfg.user_num + 1, "zipf_alpha": cfg.alpha, "mm_features": list(cfg.mm_ids),
                           "row_update": "fused sparse AdamW on the owner", "path": args.path,
                           "row_fetch": ("in place from the owners' shards over NVLink peer memory, fused into the projection kernel"
                                         if getattr(m.ops, "peers", None) is not None else "NCCL all-to-all of deduplicated rows"),
                           "small_messages": ("counts / ids / dense gradients by peer-memory kernels + device-side barriers (no NCCL "
                                              "collective in the step)" if getattr(m.ops, "io", None) is not None
                                              else "NCCL all-gather / all-to-all / all-reduce"),
                           "grad_exchange": ("owners pull the bucketed gradient rows in place over NVLink peer memory inside the "
                                             "segmented reduce (+ one barrier)" if getattr(m.ops, "grad_peers", None) is not None
                                             else "NCCL all-to-all of locally reduced gradient rows"),
                           "rows_per_step": int(total_rows // args.steps),
                           "l2": "per-step working set >> 126 MB L2; distinct batches cycled"},
                "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": float(te_sum[1]) / float(te_max[0]), "unit": bench.UNIT,
                        "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": 4 + 8 * world * 4,
                        "ms_per_step": round(float(te_max[0]) / e2e_steps * 1e3, 3),
                        "entry": "prefetch + feat2emb_packed x3 + backward + all-reduce + fused_step per rank from pinned "
                                 "host buffers (copy stream two steps ahead), loss read back every step; wall clock, max over "
                                 "ranks; feed: " + feed_how,
                        "loss_finite": bool(np.all(np.isfinite(losses)))},
                "gpu_launches": launches, "clocks": clk, "parity": parity,
                "parity_ok": None if parity is None else parity["parity_ok"]}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
