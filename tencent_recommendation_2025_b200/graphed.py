"""A whole training step of the embedding path as ONE CUDA graph (B200: the step is ~57 short kernels; issued from Python
they cost 0.99 ms of host time per 1.03 ms of device time at the C2 benchmark, tools/host_profile.py — replayed from a graph
the host cost is one launch).

What has to hold for a replay to be the step of a NEW batch:

* the launch sequence depends on call SHAPES only — fixed-shape slim calls (``resident.CallShape``: user tokens / array values
  padded to capacities) and ``PackedBatch.n_cap``: the kernels read the lookup count from device memory
  (``tgr_fact_group_t.n_is_capacity``), the host count sizes nothing;
* every input the kernels read sits at a fixed address: ONE static int32 buffer holds the step's slim calls, the item-side
  features are expanded from the HBM-resident tables inside the graph (``resident.ResidentItemFeatures.expand``);
* per-step scalars live in device memory: the AdamW bias corrections of the row update are a 48-byte ``tgr_adam_t`` block
  the runner refreshes before each replay (``FactoredEngine.adam_dev`` -> ``tgr_adam_rows_dev``); a dense optimizer inside
  the body must be ``capturable=True``.

``body(pbs)`` is the user's step on the static ``PackedBatch``es — prefetch, ``feat2emb_packed`` x3, trunk + loss (or injected
upstream gradients), ``backward()``, the dense optimizer step, ``fused_step`` — run eagerly a few times, then captured.
Results are identical to the eager step (same kernels, same order; tests/test_gpu_graphed.py compares bit for bit).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import torch

from . import _lib
from .packed import PackedBatch
from .resident import CallShape, ResidentItemFeatures, SlimCall, SlimStep


class GraphedStep:
    def __init__(self, module, store: ResidentItemFeatures, example: SlimStep, body: Callable[[List[PackedBatch]], object],
                 hyper: Optional[dict] = None, warmup: int = 3, staging_slots: int = 2):
        eng = module._tgr_engine if hasattr(module, "_tgr_engine") else module.engine
        if getattr(eng, "path", "") != "factored" or eng.mode != "fused":
            raise ValueError("GraphedStep needs the factored path in fused mode")
        for sc in example.calls:
            if sc.n_cap is None:
                raise ValueError("GraphedStep needs fixed-shape slim calls (ResidentItemFeatures.slim_step(..., shapes=...))")
        self.module, self.engine, self.store, self.body = module, eng, store, body
        self.device = store.device
        self.hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2, grad_scale=1.0)
        self.hyper.update(hyper or {})
        self.template = example
        self._sig = self._signature(example)
        n = example.ints.numel()
        self.static_ints = torch.empty(n, dtype=torch.int32, device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._gather_stream = torch.cuda.Stream(device=self.device)
        self.staging = [{"buf": torch.empty(n, dtype=torch.int32, device=self.device), "filled": None, "free": None}
                        for _ in range(staging_slots)]
        self._next_slot = 0
        self._queue: List[dict] = []
        # AdamW block of the row update: device copy + a ring of pinned sources (a slot is rewritten only after its copy ran)
        self.adam_dev = torch.zeros(C.sizeof(_lib.Adam) // 4, dtype=torch.float32, device=self.device)
        self._adam_ring = [{"host": torch.zeros(C.sizeof(_lib.Adam) // 4, dtype=torch.float32, pin_memory=True), "ev": None}
                           for _ in range(8)]
        self._adam_i = 0
        if eng.adam_dev is not None:
            raise RuntimeError("the engine already belongs to a GraphedStep")
        eng.adam_dev = self.adam_dev
        # static expansion targets + the PackedBatches the body sees (device views of the static buffer)
        lay = store.layout
        self._ids, self._mm = [], []
        for sc in example.calls:
            cl = lay.calls[sc.include_user]
            self._ids.append(torch.empty((sc.T, cl.n_single), dtype=torch.int32, device=self.device))
            self._mm.append([torch.empty((sc.T, t.shape[1]), dtype=t.dtype, device=self.device) for t in store.mm_dev])
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.outputs = None
        self.replays = 0
        # ---- eager warm-up on a side stream (allocations, lazily created state, one-time function attributes) ----
        self.static_ints.copy_(example.ints, non_blocking=True)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._refresh_adam()
                self.outputs = body(self._expand())
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        # ---- capture: the group arena comes out of the graph's private pool (the engine's recycling pool is bypassed) ----
        pool, eng._arena_pool = eng._arena_pool, []
        try:
            g = torch.cuda.CUDAGraph()
            step_before = eng.step
            self._refresh_adam()
            with torch.cuda.graph(g):
                self.outputs = body(self._expand())
            self.graph = g
            # the captured launches have NOT run: the step counter the body advanced is rolled back, run() advances it
            eng.step = step_before
        finally:
            eng._arena_pool = pool

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _signature(st: SlimStep):
        return tuple((c.B, c.L, c.include_user, tuple(c.offs), tuple(c.sizes), c.n_user_tok, tuple(c.arr_begin), tuple(c.arr_nnz),
                      c.n_cap) for c in st.calls) + (tuple(st.bases),)

    def _expand(self) -> List[PackedBatch]:
        """Static slim buffer -> the step's PackedBatches; the mm row gathers run on a forked stream next to the id expansion."""
        cur = torch.cuda.current_stream(self.device)
        side = self._gather_stream if self.store.mm_dev else None
        if side is not None:
            side.wait_stream(cur)
        pbs = []
        for i, (sc, b) in enumerate(zip(self.template.calls, self.template.bases)):
            dev = self.static_ints[b:b + sc.ints.numel()]
            pbs.append(self.store.expand(sc, dev, self._ids[i], self._mm[i], mm_stream=side))
        if side is not None:
            cur.wait_stream(side)
        return pbs

    def _refresh_adam(self):
        """Upload the tgr_adam_t block of the NEXT row update (bias corrections formed on the host in double, as eager)."""
        h = self.hyper
        a = _lib.make_adam(h["lr"], h["betas"][0], h["betas"][1], h["eps"], h["weight_decay"], self.engine.step + 1,
                           h.get("grad_scale", 1.0))
        slot = self._adam_ring[self._adam_i]
        self._adam_i = (self._adam_i + 1) % len(self._adam_ring)
        if slot["ev"] is not None:
            slot["ev"].synchronize()
        C.memmove(slot["host"].data_ptr(), C.addressof(a), C.sizeof(a))
        self.adam_dev.copy_(slot["host"], non_blocking=True)
        slot["ev"] = torch.cuda.Event()
        slot["ev"].record(torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------------ feeding
    def submit(self, step: SlimStep):
        """H2D copy of a step's slim buffer into a staging slot, on the copy stream (overlaps the running step)."""
        if self._signature(step) != self._sig:
            raise ValueError("GraphedStep.submit: the step does not have the captured shape")
        slot = self.staging[self._next_slot]
        self._next_slot = (self._next_slot + 1) % len(self.staging)
        if any(q is slot for q in self._queue):
            raise RuntimeError("GraphedStep.submit: every staging slot holds a step that has not run yet")
        with torch.cuda.stream(self.copy_stream):
            if slot["free"] is not None:
                self.copy_stream.wait_event(slot["free"])       # the step that last used the slot has copied it out
            slot["buf"].copy_(step.ints, non_blocking=True)
            slot["filled"] = torch.cuda.Event()
            slot["filled"].record(self.copy_stream)
        self._queue.append(slot)

    def load(self, dev_ints: torch.Tensor):
        """Inputs already in HBM: one device-to-device copy into the static buffer (bench `value`, tests)."""
        self.static_ints.copy_(dev_ints, non_blocking=True)

    def run(self):
        """Replay the step on the oldest submitted batch (or on whatever ``load`` put into the static buffer)."""
        cur = torch.cuda.current_stream(self.device)
        if self._queue:
            slot = self._queue.pop(0)
            cur.wait_event(slot["filled"])
            self.static_ints.copy_(slot["buf"], non_blocking=True)
            slot["free"] = torch.cuda.Event()
            slot["free"].record(cur)
        self._refresh_adam()
        self.graph.replay()
        self.engine.step += 1
        self.replays += 1
        return self.outputs

    def close(self):
        if self.engine.adam_dev is self.adam_dev:
            self.engine.adam_dev = None
