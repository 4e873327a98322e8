"""bench.py — embedding fwd+bwd rows/sec of the TencentGR sparse-feature path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c2|c3|c4|c5]
                    [--path factored|concat] [--graph auto|plain|off]

One "step" = one training step's worth of the hot path on one synthetic batch:
  3 x feat2emb forward (seq with users, pos, neg: model.py:324,376-377)
  + backward from injected upstream gradients
  + AdamW on the path's Linear layers + fused sparse AdamW row update of every touched table row.
"row" = one non-padding table-row lookup in a forward call.

N = 1 (default path `factored`: itemdnn / userdnn folded into the step's deduplicated rows):
  `value`  rows/s of the step replayed from a CUDA graph with the NEXT batch's key processing on a branch of the same graph
           (graphed.PipelinedStep), slim inputs resident in HBM, CUDA events around K replays; `eager` in the line = the same
           step issued from Python (host-bound), `--graph plain` = one step per graph, `--graph off` = eager only.
  `e2e`    the same replays fed from pinned HOST buffers (ids + user tokens, one H2D copy per step on a copy stream; the item
           feature / mm tables are resident in HBM) with the step's result copied back and read on the host every step.
  `roofline`  the C-ABI entry with the largest CUDA-event time (events recorded inside the library on the launching stream,
           an eager pass over the same steps); `roofline.kernels` lists every entry.
  `cpu_baseline` / `gpu_eager_baseline`  the UNMODIFIED reference (baseline/_ref) on the host cores / as torch eager on the GPU.
N > 1 (torchrun): bench_sharded.py — tables row-sharded over the ranks, peer-memory exchange, value = all ranks' rows / max time.
`--impl reference` times the unmodified reference's CPU implementation of the same step (full batches, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tencent_recommendation_2025_b200 import synth  # noqa: E402

METRIC = "embedding fwd+bwd rows/sec at 1/2/4/8 B200; % of HBM roofline"
UNIT = "rows/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def get_config(name: str, B: int):
    if name == "c1":
        return synth.config_c1()
    if name == "c2":
        return synth.config_c2(B)
    if name == "c3":
        return synth.config_c3(B)
    if name == "c4":
        return synth.config_c4(B)
    if name == "c5":
        return synth.config_c5(B)
    raise ValueError(name)


WORKLOADS = {
    "c1": "C1 BaseLine tiny batch (B=128, L=101, H=32, 100k items)",
    "c2": "C2 BaseLine feat2emb fwd+bwd, full feature mix (B=1024/GPU, L=101, H=64, 5M-row item table, mm '81')",
    "c3": "C3 BaseLineO1 + frozen mm features '81'+'82' (1024-d) projection",
    "c4": "C4 row-sharded 50M-row item/user tables",
    "c5": "C5 high-skew Zipf 1.2, 50M-row tables",
}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 20):
        self.gpu = gpu_index
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:   # nvidia-smi takes a moment to emit its first sample
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples from here on belong to the timed region."""
        self.first = len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = self.lines[getattr(self, "first", 0):]
        # a timed region shorter than a couple of sampling periods falls back to warm-up + timed samples (same load)
        for ln in (timed if len(timed) >= 3 else self.lines[1:]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def init_module(cfg: synth.SynthConfig, device, mode="fused", path="concat"):
    from tencent_recommendation_2025_b200.module import BaselineEmbedding
    args = types.SimpleNamespace(device=str(device), hidden_units=cfg.H)
    with torch.device(device):
        m = BaselineEmbedding(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args, mode=mode, path=path)
    g = torch.Generator(device=device).manual_seed(0)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.05, generator=g)
            else:
                p.normal_(0.0, 0.1, generator=g)
        for p in m.engine.tables:
            p[0].zero_()
    return m


def algorithmic_bytes(lay, calls, uniq_per_call, uniq_step, cat_esz=4, x_esz=4):
    """SURVEY.md §8(d), counted compulsory from the actual batch (the REFERENCE dataflow: concat buffers written in
    the forward and their gradients read in the backward). Returns (fwd_bytes, bwd_bytes)."""
    H = lay.H
    fwd = bwd = 0
    for pc, U in zip(calls, uniq_per_call):
        cl = lay.calls[pc.include_user]
        S = pc.ids.size + pc.arr_val.size + pc.arr_off.size
        D = cl.item_dim + cl.user_dim
        X = sum(lay.item_emb_feat.values())
        fwd += 4 * S + U * H * 4 + pc.T * D * cat_esz + pc.T * X * x_esz
        bwd += 4 * S + pc.T * D * cat_esz + pc.T * X * x_esz
    bwd += uniq_step * 6 * H * 4
    return fwd, bwd


def unique_counts(lay, calls):
    from oracle import feat2emb_numpy as onp  # checker-side accounting only (exact U), never the timed path
    per = []
    allk = []
    for pc in calls:
        k, _ = onp.build_keys(lay, [pc])
        per.append(int(np.unique(k).size))
        allk.append(k)
    return per, int(np.unique(np.concatenate(allk)).size)


def kernel_bytes_step(lay, st, pbs, per_u, step_u, path, x_esz=4):
    """Compulsory (algorithmic) bytes of OUR kernels for one step, by C-ABI entry name. Every operand is counted once
    per launch that must touch it (a row read by many lookups counts once: cache hits cannot inflate the figure)."""
    H = lay.H
    n = sum(pb.n_valid for pb in pbs)
    X = sum(lay.item_emb_feat.values())
    S = sum(pc.ids.size + pc.arr_val.size for pc in st.calls)
    out = {}
    if path == "concat":
        fw = 0
        for pc, U in zip(st.calls, per_u):
            cl = lay.calls[pc.include_user]
            d_tab = cl.item_dim + cl.user_dim - H * cl.n_mm
            fw += 4 * (pc.ids.size + pc.arr_val.size + pc.arr_off.size) + U * H * 4 + pc.T * d_tab * 4
        out["fwd_gather_pool_concat"] = fw
        out["bwd_reduce"] = 8 * n + n * H * 4 + step_u * 6 * H * 4
        out["build_keys"] = 4 * S + 8 * n
        out["sort_pairs"] = 16 * n
        out["mm_proj_fwd"] = sum(pc.T * (X * 4 + H * 4 * len(lay.item_emb_feat)) for pc in st.calls)
        out["mm_proj_bwd"] = out["mm_proj_fwd"]
        return out
    U = step_u
    out["build_keys"] = 4 * S + 8 * n
    out["sort_pairs"] = 16 * n
    out["dedup"] = 4 * n + 4 * n + 4 * U
    out["remap_scatter"] = 8 * n + 4 * n
    out["fact_project_rows"] = 2 * U * H * 4
    # mm features: wide bf16 ones run on the tcgen05 kernels (forward x read + [T, H] write; backward x + bf16 dz read)
    tc_dims = [d for d in lay.item_emb_feat.values() if x_esz == 2 and d >= 128 and d % 64 == 0]
    sm_dims = [d for d in lay.item_emb_feat.values() if d not in tc_dims]
    Tsum = sum(pc.T for pc in st.calls)
    out["mm_proj_fwd"] = Tsum * sum(d * x_esz + H * 4 for d in sm_dims)
    if tc_dims:
        out["mm_proj_fwd_tc"] = Tsum * sum(d * 2 + H * 4 for d in tc_dims)
        out["mm_proj_bwd_tc"] = Tsum * sum(d * 2 + H * 2 for d in tc_dims)
        out["cast_bf16"] = Tsum * H * 6
    fw = rm = red = 0
    for pc, Uc in zip(st.calls, per_u):
        sides = 2 if pc.include_user else 1
        fw += 4 * pc.ids.size + 4 * pc.arr_val.size + Uc * H * 4 + pc.T * H * 4 * (1 + len(lay.item_emb_feat)) + pc.T * H // 4
        rm += pc.T * H * 4 * (1 + sides) + pc.T * H // 4 + pc.T * 32 * x_esz * (1 if 32 in lay.item_emb_feat.values() else 0)
        red += pc.T * H * 4 * sides
    out["fact_forward"] = fw
    out["fact_relu_mask"] = rm
    out["bwd_reduce"] = 12 * n + red + U * H * 4
    out["fact_unique_backward"] = 3 * U * H * 4
    out["adam_rows"] = 7 * U * H * 4
    return out


def run_gpu(args):
    from tencent_recommendation_2025_b200 import _lib
    from tencent_recommendation_2025_b200.packed import HostPrefetcher, stage_pinned, to_device
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if world > 1:
        from bench_sharded import run_sharded  # row-sharded tables + all-to-all
        return run_sharded(args, rank, world, local_rank)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # concat path only: itemdnn/userdnn stay the caller's torch Linear calls; the reference's launcher runs them with
    # TF32 (run.sh:8 --use_tf32 -> main.py:66-68). The factored path computes them in fp32 inside its own kernels.
    torch.backends.cuda.matmul.allow_tf32 = args.dnn_matmul == "tf32"
    hbm_peak, peak_src = load_peaks()
    cfg = get_config(args.config, args.batch)
    # frozen mm features: fp32 as the reference holds them, or bf16 storage (config 3: the wide features go through the
    # bf16 tcgen05 projection)
    mm_dtype = torch.bfloat16 if (args.mm_dtype or ("bf16" if args.config == "c3" else "f32")) == "bf16" else torch.float32
    worldgen = synth.SynthWorld(cfg, 0)
    lay = worldgen.layout
    m = init_module(cfg, dev, "fused", args.path)
    # factored path: the engine also owns the path's Linear layers (itemdnn / userdnn / emb_transform): one AdamW launch next
    # to the row update instead of torch's multi-tensor AdamW (37 us of device time + 0.14 ms of host time per step)
    own_dense = args.path == "factored" and not args.torch_dense_opt
    if own_dense:
        m.own_dense_parameters()
    dense_opt = None if own_dense else torch.optim.AdamW(m.dense_parameters(), lr=1e-3, betas=(0.9, 0.98), fused=True)
    eng = m.engine
    n_batches = max(1, min(args.batches, args.steps + args.warmup))
    steps_np = [worldgen.make_step(s) for s in range(n_batches)]
    dev_steps = []
    for st in steps_np:
        pbs = [to_device(lay, pc, dev, mm_dtype=mm_dtype) for pc in st.calls]
        ups = [torch.from_numpy(r).to(dev) for r in st.upstream]
        dev_steps.append((pbs, ups))
    torch.cuda.synchronize()
    hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    lookups = [st.n_lookups() for st in steps_np]   # host-side row counting stays out of the timed regions

    def one_step(pbs, ups):
        if dense_opt is not None:
            dense_opt.zero_grad(set_to_none=True)
        m.prefetch(pbs)                          # factored path: one key sort / row projection per step
        outs = [m.feat2emb_packed(pb) for pb in pbs]
        torch.autograd.backward(outs, ups)
        if dense_opt is not None:
            dense_opt.step()
            m.fused_step(**hyper)
        else:
            m.fused_step(**hyper, dense=True)
        return outs

    def step_result(outs):
        """The scalar read back every step in the e2e legs: a checksum of the step's outputs (the last position of the first
        sequence of every call). The real loss comes from the trunk, which is not on this path."""
        with torch.no_grad():
            return torch.stack([o.detach()[0, -1] for o in outs]).sum()

    # ---- device-resident timing (value) -------------------------------------------------------
    clocks = ClockSampler(local_rank, period_ms=args.clock_period_ms)
    if not args.no_clocks:
        clocks.start()
    for i in range(max(args.warmup, 3 * n_batches)):   # every distinct batch shape several times (allocator warm)
        one_step(*dev_steps[i % n_batches])
    torch.cuda.synchronize()
    clocks.mark()
    l0 = _lib.launch_count()    # counted inside the library at every launch site (tgr_launch_count)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    rows = 0
    for i in range(args.steps):
        k = (args.warmup + i) % n_batches
        one_step(*dev_steps[k])
        rows += lookups[k]
    ev1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - l0
    value = rows / (ms * 1e-3)

    # ---- per-kernel CUDA-event timing inside the library (a second pass over the same steps; the event pairs are
    #      recorded on the launching stream around every C-ABI entry) -> roofline of the dominant kernel -------------
    kern_ms = {}
    used = [(args.warmup + i) % n_batches for i in range(args.steps)]
    if not args.no_kernel_timing:
        _lib.timing_enable(True)
        for k in used:
            one_step(*dev_steps[k])
        torch.cuda.synchronize()
        kern_ms = _lib.timing_collect()
        _lib.timing_enable(False)
    alg_f = alg_b = 0
    stats = {}
    for k in sorted(set(used)):
        pu, su = unique_counts(lay, steps_np[k].calls)
        f, b = algorithmic_bytes(lay, steps_np[k].calls, pu, su, x_esz=2 if mm_dtype == torch.bfloat16 else 4)
        stats[k] = (pu, su, f, b, kernel_bytes_step(lay, steps_np[k], dev_steps[k][0], pu, su, args.path,
                                                    2 if mm_dtype == torch.bfloat16 else 4))
    kbytes = {}
    nominal = 0
    for k in used:
        pu, su, f, b, kb = stats[k]
        alg_f += f
        alg_b += b
        for name, v in kb.items():
            kbytes[name] = kbytes.get(name, 0) + v
        nominal += sum(pb.n_valid for pb in dev_steps[k][0]) * lay.H * 4
    table = {}
    for name, (t_ms, cnt) in kern_ms.items():
        e = {"ms_per_step": round(t_ms / args.steps, 4), "launches_per_step": round(cnt / args.steps, 2)}
        if name in kbytes and t_ms > 0:
            gbs = kbytes[name] / (t_ms * 1e-3) / 1e9
            e.update({"alg_MB_per_step": round(kbytes[name] / args.steps / 1e6, 1), "GBs": round(gbs, 1),
                      "frac": round(gbs / hbm_peak, 4)})
        table[name] = e
    # the two row kernels are fp32 FFMA GEMMs over the unique rows (2*U*H*H and 4*U*H*H flop): their own ceiling is the
    # CUDA-core rate (148 SMs x 128 FMA/clk x 2 at the max SM clock), stated next to the HBM figure
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    u_total = sum(stats[k][1] for k in used)
    for name, mult in (("fact_project_rows", 2), ("fact_unique_backward", 4)):
        if name in table and table[name]["ms_per_step"] > 0:
            tf = mult * u_total * lay.H * lay.H / (kern_ms[name][0] * 1e-3) / 1e12
            table[name].update({"fp32_TFLOPs": round(tf, 2), "frac_of_fp32_ffma_peak": round(tf / fp32_peak, 3)})
    cand = [n_ for n_ in table if "GBs" in table[n_]]
    dom = max(cand, key=lambda n_: table[n_]["ms_per_step"]) if cand else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": table[dom]["GBs"] if dom else 0.0, "peak": hbm_peak,
                "unit": "GB/s", "frac": table[dom]["frac"] if dom else 0.0,
                "traffic": load_traffic(dom, table[dom]["launches_per_step"] if dom else 1),
                "peak_source": peak_src,
                "kernel_ms_per_step": table[dom]["ms_per_step"] if dom else 0.0,
                "launches": int(kern_ms[dom][1]) if dom else 0,
                "note": ("achieved = compulsory bytes of the entry / its CUDA-event time; the gather-sum forward and the "
                         "segmented reduce read each row once per LOOKUP out of L2 (nominal per-lookup traffic "
                         f"{round(nominal / args.steps / 1e6)} MB/step each), so their compulsory-HBM fraction is low by "
                         "design") if args.path == "factored" else "",
                "kernels": table,
                "kernels_ms_per_step_sum": round(sum(v["ms_per_step"] for v in table.values()), 4),
                "reference_dataflow_GB_per_step": round((alg_f + alg_b) / args.steps / 1e9, 3),
                "reference_dataflow_frac_of_hbm_peak": round((alg_f + alg_b) / (ms * 1e-3) / 1e9 / hbm_peak, 4)}

    # ---- end to end from host buffers (e2e): pinned host batches -> copy stream -> step -> loss read back ----------
    def e2e_loop(feeder, host_steps, entry):
        e2e_steps = max(3, min(args.steps, 100))
        e2e_warm = n_batches + 1
        loss_host = torch.zeros(2, dtype=torch.float32, pin_memory=True)
        loss_ev = [None, None]
        h2d = d2h = 0
        rows_e2e = 0
        losses = []
        feeder.submit(host_steps[0])
        torch.cuda.synchronize()
        t0 = None
        for i in range(e2e_warm + e2e_steps):
            k = i % n_batches
            if i == e2e_warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            pbs = feeder.take()                                        # waits (on the stream) for this step's H2D copies
            feeder.submit(host_steps[(i + 1) % n_batches])             # next step's copies overlap this step's kernels
            outs = one_step(pbs, dev_steps[k][1])
            feeder.retire()
            loss = step_result(outs)
            slot = i & 1
            if loss_ev[slot] is not None:                              # step i-2's loss has long arrived: read it
                loss_ev[slot].synchronize()
                losses.append(float(loss_host[slot]))
            loss_host[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)   # D2H of the step's result
            loss_ev[slot] = torch.cuda.Event()
            loss_ev[slot].record()
            if i >= e2e_warm:
                rows_e2e += lookups[k]
                h2d += sum(pb.h2d_bytes for pb in pbs)
                d2h += 4
        for slot in ((e2e_warm + e2e_steps) & 1, (e2e_warm + e2e_steps + 1) & 1):
            if loss_ev[slot] is not None:
                loss_ev[slot].synchronize()
                losses.append(float(loss_host[slot]))
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        feeder.take()   # drain the look-ahead submission
        feeder.retire()
        return {"value": rows_e2e / t, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps,
                "d2h_bytes_per_step": d2h // e2e_steps, "ms_per_step": round(t / e2e_steps * 1e3, 3), "entry": entry,
                "loss_finite": bool(np.all(np.isfinite(losses)))}, t / e2e_steps

    host_steps = [[stage_pinned(lay, pc, mm_dtype) for pc in st.calls] for st in steps_np]   # as a pin_memory DataLoader would
    feeder = HostPrefetcher(dev)
    e2e_packed, t_e2e_step = e2e_loop(
        feeder, host_steps,
        "BaselineEmbedding.prefetch + feat2emb_packed x3 + backward + fused_step from pinned host PACKED buffers (every "
        "token's feature ids and mm vectors cross PCIe); H2D of step k+1 on a copy stream overlaps step k, every step's loss "
        "is copied back and read on the host (two steps later, so the read never stalls the queue); wall clock")
    e2e = e2e_packed
    e2e_resident = None
    if args.resident_items == "on" or (args.resident_items == "auto" and args.config in ("c1", "c2")):
        # item features resident in HBM (they are functions of the item id, dataset.py:159,260-263): the host hands over
        # ids + user tokens only, the packed calls are rebuilt on the device (resident.py, csrc/tgr_resident.cu)
        from tencent_recommendation_2025_b200.resident import ResidentFeeder, ResidentItemFeatures
        try:
            store = ResidentItemFeatures.from_world(worldgen, dev, mm_dtype)
            slim_steps = [[store.slim(pc) for pc in st.calls] for st in steps_np]
            e2e_resident, t_e2e_step = e2e_loop(
                ResidentFeeder(store), slim_steps,
                "the same step fed from pinned host buffers holding ids + user tokens only: the item-side feature table "
                f"[{store.n_items}, {store.feat_dev.shape[1]}] and mm tables are resident in HBM and the packed calls are expanded on "
                "the device (ResidentItemFeatures / ResidentFeeder); H2D on a copy stream one step ahead, loss read back every "
                "step; wall clock")
            e2e = e2e_resident
        except Exception as exc:   # e.g. not enough host memory for the table build: keep the packed-feed number
            e2e_resident = {"error": repr(exc)}
            store = None
    else:
        store = None
    e2e_steps = 1

    # ---- the step as ONE CUDA graph (graphed.GraphedStep): the eager loop above is host-bound (tools/host_profile.py:
    #      0.99 ms of Python/torch enqueue per 1.03 ms step), a replay costs the host one launch. Fixed-shape slim calls, lookup
    #      count and AdamW bias corrections in device memory; bit-identical to the eager step (tests/test_gpu_graphed.py) -----
    eager = {"value": value, "ms_per_step": ms / args.steps, "gpu_launches": launches, "clocks": clk, "e2e": e2e,
             "how": "the same step issued eagerly from Python (one C-ABI call per phase)"}
    graph_leg = None
    if store is not None and args.graph != "off" and args.path == "factored":
        try:
            from tencent_recommendation_2025_b200.graphed import GraphedStep, PipelinedStep
            piped = args.graph == "auto"       # the next batch's key processing on a forked branch of the same graph
            from tencent_recommendation_2025_b200.resident import CallShape
            shapes = [CallShape.covering([st.calls[i] for st in steps_np]) for i in range(len(steps_np[0].calls))]
            fixed = [store.slim_step(st.calls, shapes) for st in steps_np]        # pinned, one buffer per step
            dev_fixed = [f.ints.to(dev) for f in fixed]
            cap_opt = None if own_dense else torch.optim.AdamW(m.dense_parameters(), lr=1e-3, betas=(0.9, 0.98), fused=True,
                                                               capturable=True)
            ups0 = dev_steps[0][1]     # upstream gradients are the trunk's stand-in: a static buffer inside the graph

            def body(pbs):
                if cap_opt is not None:
                    cap_opt.zero_grad(set_to_none=True)
                m.prefetch(pbs)
                outs = [m.feat2emb_packed(pb) for pb in pbs]
                torch.autograd.backward(outs, ups0)
                if cap_opt is not None:
                    cap_opt.step()
                    m.fused_step(**hyper)
                else:
                    m.fused_step(**hyper, dense=True)
                return step_result(outs)

            g_warm = 3
            if piped:
                runner = PipelinedStep(m, store, fixed[0], body, hyper=hyper, warmup=g_warm)
            else:
                runner = GraphedStep(m, store, fixed[0], body, hyper=hyper, warmup=g_warm)
            per_replay = runner.launches_per_replay      # this library's launches inside one captured step
            clocks2 = ClockSampler(local_rank, period_ms=args.clock_period_ms)
            if not args.no_clocks:
                clocks2.start()
            # pipelined: run() computes the batch loaded one call EARLIER (its key processing ran inside the previous replay)
            # and prepares the batch just loaded; every replay does one full step's work, rows are counted per computed batch
            gw = max(args.warmup, 3)
            if piped:
                runner.prime(dev_fixed[0])
            for i in range(gw):
                runner.load(dev_fixed[(i + 1) % n_batches] if piped else dev_fixed[i % n_batches])
                runner.run()
            torch.cuda.synchronize()
            clocks2.mark()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            g_rows = 0
            for i in range(args.steps):
                k = (gw + i) % n_batches           # the batch this replay computes
                runner.load(dev_fixed[(k + 1) % n_batches] if piped else dev_fixed[k])   # inputs resident in HBM: one 3 MB D2D copy
                runner.run()
                g_rows += lookups[k]
            g1.record()
            torch.cuda.synchronize()
            g_clk = clocks2.stop()
            g_ms = g0.elapsed_time(g1)
            # e2e: pinned host slim buffers -> staging (copy stream, one step ahead) -> replay -> loss read back every step
            e2e_n = max(3, min(args.steps, 100))
            e2e_warm = n_batches + 1
            loss_host = torch.zeros(2, dtype=torch.float32, pin_memory=True)
            loss_ev = [None, None]
            losses, h2d, rows_e2e, t0 = [], 0, 0, None
            if piped:
                torch.cuda.synchronize()
                runner.prime(fixed[0])
                runner.submit(fixed[1 % n_batches])
            else:
                runner.submit(fixed[0])
            for i in range(e2e_warm + e2e_n):
                k = i % n_batches
                if i == e2e_warm:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                # the H2D copy of a later step overlaps this step's replay (pipelined: two ahead, the replay prepares step i + 1)
                runner.submit(fixed[(i + 2) % n_batches] if piped else fixed[(i + 1) % n_batches])
                loss = runner.run()
                slot = i & 1
                if loss_ev[slot] is not None:
                    loss_ev[slot].synchronize()
                    losses.append(float(loss_host[slot]))
                loss_host[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)
                loss_ev[slot] = torch.cuda.Event()
                loss_ev[slot].record()
                if i >= e2e_warm:
                    rows_e2e += lookups[k]
                    h2d += fixed[k].ints.numel() * 4 + 48
            for ev in loss_ev:
                if ev is not None:
                    ev.synchronize()
            torch.cuda.synchronize()
            t_g = time.perf_counter() - t0
            runner.run()                                        # drain the look-ahead submission(s)
            if piped:
                runner.run()
            torch.cuda.synchronize()
            graph_leg = {"value": g_rows / (g_ms * 1e-3), "ms_per_step": g_ms / args.steps,
                         "kernels_per_replay": int(per_replay), "clocks": g_clk,
                         "e2e": {"value": rows_e2e / t_g, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_n, "d2h_bytes_per_step": 4,
                                 "ms_per_step": round(t_g / e2e_n * 1e3, 3),
                                 "entry": ("PipelinedStep" if piped else "GraphedStep") + ".submit (pinned slim buffer: ids + user tokens, one H2D copy on a copy stream, "
                                          "one step ahead) + GraphedStep.run (48-byte AdamW block + one graph replay: expansion from "
                                          "the HBM-resident item tables, prefetch, feat2emb x3, backward, dense AdamW, row update) + "
                                          "the step's loss copied back and read on the host every step; wall clock",
                                 "loss_finite": bool(np.all(np.isfinite(losses)))}}
            runner.close()
            runner = dev_fixed = None
            value, ms, launches, clk = graph_leg["value"], g_ms, int(per_replay) * args.steps, g_clk
            e2e, t_e2e_step = graph_leg["e2e"], t_g / e2e_n
        except Exception as exc:   # the eager numbers stand
            import traceback
            graph_leg = {"error": repr(exc), "trace": traceback.format_exc()[-1500:]}
    store = None

    # ---- the reference itself, beside the number: on the box's host cores and as torch eager on this GPU -----------
    cpu = None
    gpu_eager = None
    if not args.no_cpu_baseline:
        del dev_steps, host_steps, feeder
        m = eng = dense_opt = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            gpu_eager = gpu_eager_reference(cfg, dev)
            gpu_eager["ours_over_eager_device"] = round(gpu_eager["device_ms_per_step"] / (ms / args.steps), 2) \
                if "device_ms_per_step" in gpu_eager else None
            gpu_eager["ours_e2e_over_eager_wall"] = round(gpu_eager["wall_ms_per_step"] / (t_e2e_step * 1e3), 2) \
                if "wall_ms_per_step" in gpu_eager else None
        except Exception as e:   # a side measurement never takes the benchmark line down
            gpu_eager = {"error": repr(e)}
        try:
            cpu = cpu_reference(cfg, steps=1, warmup=1, n_batches=2, prebuilt_steps=1)
        except Exception as e:
            cpu = {"error": repr(e)}

    conf = workload_config(args, cfg)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mm_dtype == torch.float32 else "f32 tables / bf16 mm features (tcgen05 projection)",
            "data": "synthetic", "config": conf,
            "impl_detail": {"row_update": "fused sparse AdamW (lazy rows)", "path": args.path,
                            "dnn": ("itemdnn/userdnn folded into the deduplicated rows (factored kernels; 3xTF32 tensor-core "
                                    "row GEMMs, fp32 accumulate)" if args.path == "factored"
                                    else f"torch F.linear (caller side, unchanged), {args.dnn_matmul} as reference run.sh --use_tf32"),
                            "rows_per_step": rows // args.steps, "distinct_batches": n_batches,
                            "launch": ("one CUDA graph replay per step (graphed.PipelinedStep / GraphedStep)" if graph_leg and "error" not in graph_leg
                                       else "eager: one C-ABI call per phase from Python")},
            "roofline": roofline, "cpu_baseline": cpu, "gpu_eager_baseline": gpu_eager, "e2e": e2e,
            "e2e_host_packed_feed": e2e_packed, "eager": eager, "graph": graph_leg,
            "gpu_launches": launches, "clocks": clk,
            "dict_tensorizer": None if args.no_cpu_baseline else tensorizer_timing(cfg)}
    print(json.dumps(line))


def tensorizer_timing(cfg: synth.SynthConfig, sequences: int = 64):
    """Host cost of the reference-signature entry (list-of-dict features -> packed calls), on a bounded sample of the
    workload: the C tensorizer (libtgr_pack.so) next to the numpy restatement. Reported apart from the metric
    (SURVEY.md §8(d)): a data pipeline packs in its DataLoader workers, the benchmark steps start from packed calls."""
    try:
        import dataclasses
        from tencent_recommendation_2025_b200.packed import pack_from_dicts, pack_from_dicts_py
        scfg = dataclasses.replace(cfg, B=min(sequences, cfg.B))
        world = synth.SynthWorld(scfg, 0)
        lay = world.layout
        st = world.make_step(0)
        out = {"sample": f"{scfg.B} sequences x 3 calls", "tokens": 0}
        t_c = t_py = 0.0
        for pc in st.calls:
            d = synth.packed_to_dicts(lay, pc)
            seq = torch.from_numpy(pc.seq)
            mask = torch.from_numpy(pc.mask) if pc.include_user else None
            t0 = time.perf_counter()
            pack_from_dicts(lay, seq, d, mask, pc.include_user)
            t1 = time.perf_counter()
            pack_from_dicts_py(lay, seq, d, mask, pc.include_user)
            t2 = time.perf_counter()
            t_c += t1 - t0
            t_py += t2 - t1
            out["tokens"] += pc.T
        out["c_us_per_token"] = round(1e6 * t_c / out["tokens"], 3)
        out["numpy_us_per_token"] = round(1e6 * t_py / out["tokens"], 3)
        return out
    except Exception as e:   # never let a host-side side measurement take the benchmark line down
        return {"error": repr(e)}


def load_traffic(kernel, launches_per_step=1):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the C-ABI entry `kernel`, from the committed
    `ncu --set full` capture of one C2 step (profiles/traffic.json holds bytes per STEP per entry), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if kernel is None or not os.path.exists(p):
        return None
    try:
        v = json.load(open(p)).get(kernel)
        return None if v is None else int(v / max(launches_per_step, 1))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def workload_config(args, cfg: synth.SynthConfig):
    """The workload-defining keys — identical in our arm and in the reference arm (same_config)."""
    return {"workload": WORKLOADS[args.config], "batch_per_gpu": cfg.B, "seq_len": cfg.L, "hidden": cfg.H,
            "item_rows": cfg.item_num + 1, "user_rows": cfg.user_num + 1, "zipf_alpha": cfg.alpha,
            "mm_features": list(cfg.mm_ids), "tokens_per_step": 3 * cfg.B * cfg.L,
            "step": "3 x feat2emb (seq+users, pos, neg) + backward from injected upstream grads + AdamW row update",
            "l2": "tables + AdamW state >> 126 MB L2; distinct batches cycled (each step touches different rows)"}


class ReferenceRunner:
    """The UNMODIFIED reference ``BaselineModel`` (staged copy under baseline/_ref, see baseline/ref_loader.py) driven
    exactly as its training step drives the hot path: ``feat2emb`` x3 through the stock list-of-dict signature
    (model/BaseLine/model.py:226-310, called at :324,376-377), upstream gradients injected at feat2emb's output
    (SURVEY.md F13), ``torch.optim.AdamW(betas=(0.9, 0.98))`` over the hot-path parameters (main.py:131,188-190).
    Runs on ``device`` 'cpu' (the --impl reference arm / cpu_baseline) or 'cuda' (gpu_eager_baseline)."""

    def __init__(self, cfg: synth.SynthConfig, device: str, n_batches: int, variant: str = "BaseLine"):
        from baseline import ref_loader
        self.cfg, self.device = cfg, device
        self.model = ref_loader.build_model(cfg, device, variant)
        self.opt = torch.optim.AdamW(ref_loader.hot_params(self.model), lr=1e-3, betas=(0.9, 0.98))
        world = synth.SynthWorld(cfg, 0)
        lay = world.layout
        self.batches = []
        for s in range(n_batches):
            st = world.make_step(s)
            calls = []
            for pc in st.calls:
                seq = torch.from_numpy(pc.seq)                       # int32 [B, L] as dataset.py:123,284 hands it over
                mask = torch.from_numpy(pc.mask) if pc.include_user else None
                calls.append((seq, synth.packed_to_dicts(lay, pc), mask, pc.include_user))
            ups = [torch.from_numpy(r).to(device) for r in st.upstream]
            self.batches.append((calls, ups, st.n_lookups()))

    def step(self, k: int) -> int:
        calls, ups, rows = self.batches[k % len(self.batches)]
        m = self.model
        self.opt.zero_grad(set_to_none=True)
        outs = [m.feat2emb(seq, feats, mask=mask, include_user=iu) for seq, feats, mask, iu in calls]
        torch.autograd.backward(outs, ups)
        self.opt.step()
        return rows

    def prebuild(self):
        """Figure (ii) of SURVEY.md §8(d): ``feat2tensor`` outputs pre-built. The method is replaced ON THE INSTANCE by a
        lookup of tensors the stock method produced once; feat2emb's own mm fill loop (model.py:288-293) still runs."""
        stock = self.model.feat2tensor
        cache = {}
        for calls, _, _ in self.batches:
            for _, feats, _, iu in calls:
                groups = [self.model.ITEM_SPARSE_FEAT, self.model.ITEM_ARRAY_FEAT]
                if iu:
                    groups += [self.model.USER_SPARSE_FEAT, self.model.USER_ARRAY_FEAT]
                for g in groups:
                    for k in g:
                        cache[(id(feats), k)] = stock(feats, k)
        self.model.feat2tensor = lambda feats, k: cache[(id(feats), k)]
        return stock


def cpu_reference(cfg: synth.SynthConfig, steps: int, warmup: int, n_batches: int = 2, prebuilt_steps: int = 1):
    """The reference's own CPU implementation of the path at the SAME configuration as the GPU arm (full batch), all
    host threads. `value` = the stock path (dict walk included, figure (i)); `feat2tensor_prebuilt` = figure (ii)."""
    from baseline import ref_loader
    if not ref_loader.available():
        raise FileNotFoundError("baseline/_ref not staged")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = ReferenceRunner(cfg, "cpu", max(1, min(n_batches, steps + warmup)))
    for i in range(warmup):
        run.step(i)
    t0 = time.perf_counter()
    rows = 0
    for i in range(steps):
        rows += run.step(warmup + i)
    dt = time.perf_counter() - t0
    out = {"value": rows / dt, "unit": UNIT, "cores": cores, "kind": "reference",
           "ms_per_step": round(dt / steps * 1e3, 1),
           "sample": f"{steps} steps of the full B={cfg.B} batch (same tables / feature mix / lookups as the GPU arm) through "
                     "the unmodified reference BaselineModel.feat2emb x3 with list-of-dict inputs (Python feat2tensor walk "
                     "included), backward from injected upstream grads, torch.optim.AdamW over all hot-path rows"}
    if prebuilt_steps > 0:
        run.prebuild()
        t0 = time.perf_counter()
        r2 = 0
        for i in range(prebuilt_steps):
            r2 += run.step(warmup + steps + i)
        d2 = time.perf_counter() - t0
        out["feat2tensor_prebuilt"] = {"value": r2 / d2, "ms_per_step": round(d2 / prebuilt_steps * 1e3, 1),
                                       "steps": prebuilt_steps,
                                       "note": "feat2tensor replaced on the instance by a lookup of its own pre-built "
                                               "outputs; the mm fill loop inside feat2emb still runs"}
    return out


def gpu_eager_reference(cfg: synth.SynthConfig, device, steps: int = 3, warmup: int = 2):
    """SURVEY.md §2.2 bar: the unmodified reference module on device='cuda' (torch eager sm_100 kernels), same batches.
    wall = stock path incl. the Python dict walk; device = summed CUDA kernel + memcpy time of one step from
    torch.profiler (what the GPU actually spends on the reference's kernels, host walk excluded)."""
    from baseline import ref_loader
    if not ref_loader.available():
        return {"unavailable": "baseline/_ref not staged"}
    torch.backends.cuda.matmul.allow_tf32 = True      # the reference launcher's own setting (run.sh --use_tf32)
    run = ReferenceRunner(cfg, str(device), 2)
    for i in range(warmup):
        run.step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = 0
    for i in range(steps):
        rows += run.step(warmup + i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"impl": "unmodified reference BaselineModel on cuda (torch eager, TF32 matmuls as run.sh), dense AdamW",
           "steps": steps, "wall_ms_per_step": round(dt / steps * 1e3, 2), "wall_rows_per_s": rows / dt}
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            r1 = run.step(warmup + steps)
            torch.cuda.synchronize()
        dev_us = 0.0
        n_k = 0
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                dev_us += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
                n_k += 1
        out.update({"device_ms_per_step": round(dev_us / 1e3, 3), "device_rows_per_s": r1 / (dev_us * 1e-6),
                    "device_kernels_per_step": n_k,
                    "device_note": "sum of CUDA kernel + memcpy durations of one step (torch.profiler / CUPTI)"})
    except Exception as e:
        out["device_error"] = repr(e)
    del run
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = get_config(args.config, args.batch)
    # a full-batch step of the reference costs ~2.1 s on 16 host cores (dict walk + dense AdamW over every row): the arm
    # times at most `budget` seconds of them so that any --steps K finishes within a few minutes; the steps stay FULL
    # batches (a smaller batch would charge the reference its per-step dense AdamW against fewer rows)
    budget_s, est_step_s = 150.0, 2.2
    steps_run = max(2, min(args.steps, int(budget_s / est_step_s)))
    warm_run = max(1, min(args.warmup, 2))
    try:
        cpu = cpu_reference(cfg, steps=steps_run, warmup=warm_run, n_batches=2, prebuilt_steps=1)
        cpu["steps_timed"], cpu["warmup_run"] = steps_run, warm_run
    except FileNotFoundError as e:
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg), "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--path", default="factored", choices=["concat", "factored"],
                    help="concat: gather/pool/concat kernels + torch itemdnn/userdnn; factored: DNN folded into unique rows")
    ap.add_argument("--batch", type=int, default=1024, help="sequences per GPU")
    ap.add_argument("--batches", type=int, default=4, help="distinct synthetic batches to cycle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resident-items", default="auto", choices=["auto", "on", "off"],
                    help="e2e leg with the item feature / mm tables resident in HBM (auto: c1 and c2)")
    ap.add_argument("--torch-dense-opt", action="store_true",
                    help="factored path: update itemdnn/userdnn/emb_transform with torch.optim.AdamW instead of the engine's "
                         "own dense AdamW launch")
    ap.add_argument("--graph", default="auto", choices=["auto", "plain", "off"],
                    help="N=1: the step replayed from a CUDA graph (needs the resident item tables). auto = pipelined (the next "
                         "batch's key processing on a forked branch of the same graph), plain = one step per graph, off = eager only")
    ap.add_argument("--mm-dtype", default=None, choices=["f32", "bf16"],
                    help="storage dtype of the frozen mm features (default: bf16 for c3, f32 otherwise)")
    ap.add_argument("--no-prefetch", action="store_true", help="sharded path: per-call exchange instead of step prefetch")
    ap.add_argument("--no-lookahead", action="store_true", help="sharded path: no one-step-ahead key processing")
    ap.add_argument("--no-p2p", action="store_true", help="sharded factored path: fetch rows with the NCCL all-to-all "
                    "instead of reading the owners' shards in place over NVLink peer memory")
    ap.add_argument("--no-symm-io", action="store_true", help="sharded path: counts / ids / dense gradients through NCCL "
                    "collectives instead of the peer-memory transport")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi clocks (debug)")
    ap.add_argument("--no-kernel-timing", action="store_true", help="no per-kernel CUDA events (debug)")
    ap.add_argument("--clock-period-ms", type=int, default=20)
    ap.add_argument("--dnn-matmul", default="tf32", choices=["tf32", "fp32"],
                    help="precision of the caller-side torch itemdnn/userdnn matmuls (reference launcher: tf32)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_gpu(args)


if __name__ == "__main__":
    main()
