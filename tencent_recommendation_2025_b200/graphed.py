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
upstream gradients), ``backward()``, the dense optimizer step, ``fused_step`` — run eagerly ``warmup`` times on the example
batch (REAL steps: they update the model), then captured.
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
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self.outputs = body(self._expand())
            self.launches_per_replay = int(_lib.launch_count() - l0)   # this library's kernels in one captured step
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


class PipelinedStep:
    """``GraphedStep`` with the value-independent half of the NEXT step inside the same graph, on a forked branch.

    Expansion, key building, sort, dedup and id remap of a batch (~170 us of issue-bound kernels at the C2 benchmark) depend on
    the batch alone, not on the tables, so replay r runs them for batch r+1 NEXT TO the row-gradient GEMMs and the row update
    of batch r (latency- / HBM-bound kernels; the branch forks after the segmented reduce, see ``_step``) instead of in front
    of its own step. Two slots (static inputs, expansion targets, group arena) alternate roles, hence two graphs: graph[s]
    computes the batch prepared in slot s and prepares slot 1-s.

    Construction runs ``max(2, warmup)`` REAL steps on ``example`` (they update the model, like the warm-up of
    ``GraphedStep``) before capturing.

        runner.prime(b0)            # batch 0 into the slot the first replay computes; its key processing runs eagerly
        runner.submit(b1)           # (or load(dev_ints)): the batch the next run() PREPARES
        out0 = runner.run()         # computes b0, prepares b1
        runner.submit(b2); out1 = runner.run() ...

    Results are bit-identical to the eager step of each batch (tests/test_gpu_graphed.py)."""

    def __init__(self, module, store: ResidentItemFeatures, example: SlimStep, body: Callable[[List[PackedBatch]], object],
                 hyper: Optional[dict] = None, warmup: int = 2, staging_slots: int = 3):
        eng = module._tgr_engine if hasattr(module, "_tgr_engine") else module.engine
        if getattr(eng, "path", "") != "factored" or eng.mode != "fused":
            raise ValueError("PipelinedStep needs the factored path in fused mode")
        for sc in example.calls:
            if sc.n_cap is None:
                raise ValueError("PipelinedStep needs fixed-shape slim calls (ResidentItemFeatures.slim_step(..., shapes=...))")
        if eng.adam_dev is not None:
            raise RuntimeError("the engine already belongs to a graphed step")
        self.module, self.engine, self.store, self.body = module, eng, store, body
        self.device = dev = store.device
        self.hyper = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2, grad_scale=1.0)
        self.hyper.update(hyper or {})
        self.template = example
        self._sig = GraphedStep._signature(example)
        n = example.ints.numel()
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._prep_stream = torch.cuda.Stream(device=dev, priority=-1)   # its short kernels go first when block slots free up
        self._fork_ev = torch.cuda.Event()
        self._fork_ev.record(torch.cuda.current_stream(dev))     # creates the underlying cudaEvent_t
        self._gather_stream = torch.cuda.Stream(device=dev)
        self.staging = [{"buf": torch.empty(n, dtype=torch.int32, device=dev), "filled": None, "free": None}
                        for _ in range(staging_slots)]
        self._next_slot = 0
        self._queue: List[dict] = []
        self.adam_dev = torch.zeros(C.sizeof(_lib.Adam) // 4, dtype=torch.float32, device=dev)
        self._adam_ring = [{"host": torch.zeros(C.sizeof(_lib.Adam) // 4, dtype=torch.float32, pin_memory=True), "ev": None}
                           for _ in range(8)]
        self._adam_i = 0
        eng.adam_dev = self.adam_dev
        lay = store.layout
        self.static_ints = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
        self._ids = [[torch.empty((sc.T, lay.calls[sc.include_user].n_single), dtype=torch.int32, device=dev) for sc in example.calls]
                     for _ in range(2)]
        self._mm = [[[torch.empty((sc.T, t.shape[1]), dtype=t.dtype, device=dev) for t in store.mm_dev] for sc in example.calls]
                    for _ in range(2)]
        for s in range(2):
            self.static_ints[s].copy_(example.ints, non_blocking=True)
        self._pbs = [self._expand(s) for s in range(2)]           # the canonical PackedBatch objects of each slot
        nbytes = eng.group_bytes(self._pbs[0])
        self._arena = [torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.consume = 0
        self.graphs: List[Optional[torch.cuda.CUDAGraph]] = [None, None]
        self.outputs = [None, None]
        self.replays = 0
        # ---- eager warm-up of the pipelined step on a side stream ----
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            eng.stage(eng.prepare(self._pbs[0], arena=self._arena[0]))
            for _ in range(max(2, warmup)):
                GraphedStep._refresh_adam(self)
                self.outputs[self.consume] = self._step(self.consume)
                self.consume ^= 1
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # ---- capture graph[c] then graph[1 - c]: each consumes the group the previous one staged ----
        step_before = eng.step
        pool, eng._arena_pool = eng._arena_pool, []
        try:
            for _ in range(2):
                c = self.consume
                g = torch.cuda.CUDAGraph()
                GraphedStep._refresh_adam(self)
                l0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    self.outputs[c] = self._step(c)
                self.launches_per_replay = int(_lib.launch_count() - l0)   # this library's kernels in one captured step
                self.graphs[c] = g
                self.consume ^= 1
        finally:
            eng._arena_pool = pool
            eng._staged.clear()
            eng.step = step_before      # the captured launches have not run
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ one pipelined step (eager or under capture)
    def _expand(self, s: int, mm_stream=None) -> List[PackedBatch]:
        pbs = []
        for i, (sc, b) in enumerate(zip(self.template.calls, self.template.bases)):
            dev = self.static_ints[s][b:b + sc.ints.numel()]
            pbs.append(self.store.expand(sc, dev, self._ids[s][i], self._mm[s][i], mm_stream=mm_stream))
        return pbs

    def _prepare_slot(self, s: int):
        """Expansion + key processing of the batch in slot s, on the current stream (+ the gather stream)."""
        cur = torch.cuda.current_stream(self.device)
        gs = self._gather_stream if self.store.mm_dev else None
        if gs is not None:
            gs.wait_stream(cur)
        self._expand(s, mm_stream=gs)                 # same static tensors as self._pbs[s]: only the launches matter
        g = self.engine.prepare(self._pbs[s], arena=self._arena[s])
        if gs is not None:
            cur.wait_stream(gs)
        return g

    def _step(self, c: int):
        """body(slot c), and the next batch's value-independent half forked off right after the body's segmented reduce:
        everything before that point (gather-sum forwards, dZ, the reduce) lives on L2 hits that a concurrent sort would
        evict (measured: the reduce took 245 us instead of 111 with the branch forked at the top of the step); what
        follows (row-gradient GEMMs, AdamW row update) is latency- / HBM-bound."""
        cur = torch.cuda.current_stream(self.device)
        prep = self._prep_stream
        self.engine.reduce_done_event = self._fork_ev
        try:
            out = self.body(self._pbs[c])             # prefetch() finds the group staged for slot c
        finally:
            self.engine.reduce_done_event = None
        prep.wait_event(self._fork_ev)                # fork point: recorded inside the finishing backward
        with torch.cuda.stream(prep):
            self.engine.stage(self._prepare_slot(1 - c))
        cur.wait_stream(prep)                         # join
        return out

    # ------------------------------------------------------------------ feeding
    def prime(self, step):
        """The first batch: into the slot the next run() computes, key processing eagerly on the current stream. ``step`` is a
        SlimStep (host) or an int32 device tensor."""
        src = step.ints if isinstance(step, SlimStep) else step
        if isinstance(step, SlimStep) and GraphedStep._signature(step) != self._sig:
            raise ValueError("PipelinedStep.prime: the step does not have the captured shape")
        self.static_ints[self.consume].copy_(src, non_blocking=True)
        self._prepare_slot(self.consume)              # fills arena[consume]; the host-side group object is not needed
        self.engine._staged.clear()

    _signature = staticmethod(GraphedStep._signature)
    submit = GraphedStep.submit

    def load(self, dev_ints: torch.Tensor):
        """Inputs already in HBM: the batch the next run() prepares (computed by the run() after it)."""
        self.static_ints[1 - self.consume].copy_(dev_ints, non_blocking=True)

    def run(self):
        """Replay: computes the batch made current by the previous run() / prime(), prepares the one just loaded / submitted."""
        cur = torch.cuda.current_stream(self.device)
        if self._queue:
            slot = self._queue.pop(0)
            cur.wait_event(slot["filled"])
            self.static_ints[1 - self.consume].copy_(slot["buf"], non_blocking=True)
            slot["free"] = torch.cuda.Event()
            slot["free"].record(cur)
        GraphedStep._refresh_adam(self)
        c = self.consume
        self.graphs[c].replay()
        self.consume ^= 1
        self.engine.step += 1
        self.replays += 1
        return self.outputs[c]

    def close(self):
        if self.engine.adam_dev is self.adam_dev:
            self.engine.adam_dev = None
