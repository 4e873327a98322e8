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


def run_sharded(args, rank: int, world: int, local_rank: int):
    from tencent_recommendation_2025_b200.sharded import ShardedBaselineEmbedding
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = args.dnn_matmul == "tf32"
    hbm_peak, peak_src = bench.load_peaks()
    cfg = bench.get_config(args.config, args.batch)
    worldgen = synth.SynthWorld(cfg, 0)
    lay = worldgen.layout
    margs = types.SimpleNamespace(device=str(dev), hidden_units=cfg.H)
    torch.manual_seed(0)                                   # identical dense parameters on every rank
    m = ShardedBaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), margs, rank, world)
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
    dense_opt = torch.optim.AdamW(dense, lr=1e-3, betas=(0.9, 0.98))
    n_batches = max(1, min(args.batches, args.steps + args.warmup))
    steps_np = [worldgen.make_step(1000 * rank + s) for s in range(n_batches)]
    dev_steps = [([to_device(lay, pc, dev) for pc in st.calls], [torch.from_numpy(r).to(dev) for r in st.upstream])
                 for st in steps_np]
    hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    lookups = [st.n_lookups() for st in steps_np]   # host-side row counting stays out of the timed regions
    # replicated dense parameters: their .grad tensors are views of ONE flat buffer, so the data-parallel
    # all-reduce is a single collective with no flatten / copy-back kernels
    flat_grad = torch.zeros(sum(p.numel() for p in dense), device=dev)
    o = 0
    for p in dense:
        p.grad = flat_grad[o:o + p.numel()].view_as(p)
        o += p.numel()

    def one_step(pbs, ups, next_pbs=None):
        flat_grad.zero_()
        if not args.no_prefetch:
            m.prefetch(pbs)                      # one dedup + one exchange for the step's three calls
        outs = [m.feat2emb_packed(pb) for pb in pbs]
        torch.autograd.backward(outs, ups)
        if next_pbs is not None and not args.no_lookahead:
            m.prepare_next(next_pbs)             # next step's key processing overlaps this step's tail
        dist.all_reduce(flat_grad)
        flat_grad.div_(world)
        dense_opt.step()
        m.fused_step(**hyper)
        if next_pbs is not None and not args.no_lookahead:
            m.finish_prepare()
        return outs

    eng = m.ops._eng
    clocks = bench.ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    for i in range(args.warmup):
        one_step(*dev_steps[i % n_batches], next_pbs=dev_steps[(i + 1) % n_batches][0])
    torch.cuda.synchronize()
    dist.barrier()
    eng.timing = {}
    m.ops.timing = eng.timing
    if clocks:
        clocks.mark()
    l0 = m.ops.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    rows = 0
    for i in range(args.steps):
        k = (args.warmup + i) % n_batches
        one_step(*dev_steps[k], next_pbs=dev_steps[(k + 1) % n_batches][0])   # the data loader is one batch ahead
        rows += lookups[k]
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clk = clocks.stop() if clocks else None
    ms_local = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_local, float(rows)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms = float(tmax[0])
    total_rows = float(tsum[1])
    kern_ms = eng.timing_summary()
    eng.timing = None
    launches = m.ops.launches - l0

    # ---- e2e from host buffers ----
    from tencent_recommendation_2025_b200.packed import stage_pinned
    host_steps = [[stage_pinned(lay, pc) for pc in st.calls] for st in steps_np]   # as a pin_memory DataLoader would
    e2e_steps = max(3, min(args.steps, 10))
    t_e2e = 0.0
    rows_e2e = 0
    h2d = 0
    nxt = None
    m.rank_state.prep = None
    for i in range(e2e_steps + 1):
        k = i % n_batches
        st = steps_np[k]
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        pbs = nxt if nxt is not None else [hp.upload(dev) for hp in host_steps[k]]
        nxt = [hp.upload(dev) for hp in host_steps[(k + 1) % n_batches]]     # next batch's H2D, also inside the timed region
        outs = one_step(pbs, dev_steps[k][1], next_pbs=nxt)
        _ = float(sum(o.sum() for o in outs).item())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i == 0:
            continue
        t_e2e += dt
        rows_e2e += lookups[k]
        h2d += sum(pb.h2d_bytes for pb in pbs)
    te = torch.tensor([t_e2e, float(rows_e2e)], dtype=torch.float64, device=dev)
    te_max = te.clone()
    dist.all_reduce(te_max, op=dist.ReduceOp.MAX)
    te_sum = te.clone()
    dist.all_reduce(te_sum, op=dist.ReduceOp.SUM)

    if rank == 0:
        H = lay.H
        # rank 0's dominant embedding kernel, against its algorithmic bytes
        used = [(args.warmup + i) % n_batches for i in range(args.steps)]
        kb = {"fwd_gather_pool_concat": 0.0}
        for k in used:
            st = steps_np[k]
            pu, su = bench.unique_counts(lay, st.calls)
            for pc, U in zip(st.calls, pu):
                cl = lay.calls[pc.include_user]
                d_tab = cl.item_dim + cl.user_dim - H * cl.n_mm
                kb["fwd_gather_pool_concat"] += 4 * (pc.ids.size + pc.arr_val.size + pc.arr_off.size) + U * H * 4 + pc.T * d_tab * 4
        dom = "fwd_gather_pool_concat"
        dom_ms, dom_n = kern_ms.get(dom, (0.0, 0))
        achieved = kb[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": dom + " (rank 0)", "achieved": round(achieved, 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": None, "peak_source": peak_src,
                    "kernels_ms_per_step": {n: round(v[0] / args.steps, 4) for n, v in kern_ms.items()}}
        line = {"metric": bench.METRIC, "value": total_rows / (ms * 1e-3), "unit": bench.UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": bench.WORKLOADS[args.config] + f", tables row-sharded over {world} GPUs "
                           "(owner = key mod W), dedup-then-all-to-all of ids/rows/grad rows over NCCL",
                           "batch_per_gpu": cfg.B, "seq_len": cfg.L, "hidden": cfg.H, "item_rows": cfg.item_num + 1,
                           "user_rows": cfg.user_num + 1, "zipf_alpha": cfg.alpha, "mm_features": list(cfg.mm_ids),
                           "row_update": "fused sparse AdamW on the owner", "dnn_matmul": args.dnn_matmul,
                           "rows_per_step": int(total_rows // args.steps),
                           "l2": "per-step working set >> 126 MB L2; distinct batches cycled"},
                "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": float(te_sum[1]) / float(te_max[0]), "unit": bench.UNIT,
                        "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": 4 + 8 * world * 4,
                        "ms_per_step": round(float(te_max[0]) / e2e_steps * 1e3, 3)},
                "gpu_launches": launches, "clocks": clk}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
