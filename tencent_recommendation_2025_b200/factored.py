"""Factored feat2emb: item/user DNN applied to the step's DEDUPLICATED rows (csrc/tgr_factored.cu).

Same contract as the concat path of ``engine.py`` — ``feat2emb`` of model/BaseLine/model.py:226-310 with its
backward and row update — but the ``[T, item_dim]`` / ``[T, user_dim]`` concat buffers never exist:

  prepare(calls)  keys -> sort -> dedup -> ids remapped to unique-row numbers   (shared with the backward)
  forward(call)   P = W_slot . row per unique row (once per group), then a per-token gather-sum of P rows,
                  bias, folded mm projection, ReLU, add -> out [T, H]
  backward(call)  dZ = dOut * relu mask (+ bias grads, mm chain); when the group's last call has its gradient:
                  G[u] = segmented sum of dZ rows by key, row grads = G . W_slot, dW += G^T (x) rows
  fused_step      one AdamW update per touched row

A *group* is the set of calls prepared together: ``prefetch(pbs)`` makes one group of a step's three calls (one
sort, one projection of rows shared between seq / pos / neg); without it every call is its own group. The DNN
weight / bias and emb_transform gradients are returned by the backward of the group's LAST call (autograd sums
the per-call returns, the earlier ones return nothing).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import Call, Dnn, check, make_adam
from .engine import _DEBUG, EmbeddingEngine, _stream
from .layout import KIND_ARRAY, KIND_MM, KIND_SINGLE, SIDE_ITEM, FeatureLayout
from .packed import PackedBatch

SUPPORTED_H = (32, 64, 128)


class FactGroup:
    """The calls prepared together: the C-side group descriptor + the arena every buffer of the group is carved from
    (sorted pairs, unique keys, remapped ids, projected rows P, masks, dZ, row gradients G). Lives until the row update."""

    def __init__(self, pbs: Sequence[PackedBatch]):
        self.pbs = list(pbs)
        self.index = {id(pb): i for i, pb in enumerate(self.pbs)}
        self.c = _lib.FactGroup()
        self.arena: Optional[torch.Tensor] = None
        self.n = 0
        self.n_fwd = 0             # forwards that will get a gradient
        self.n_bwd = 0
        self.acc: Optional[dict] = None   # dW_item, dW_user, db_item, db_user, dWmm{}, dbmm{} + the C struct
        self.done = False
        self.pool: Optional[list] = None  # the engine's free-arena list; the arena goes back when the group dies

    def release(self):
        """Hand the arena back for the next group (all work is stream-ordered, so reuse needs no event)."""
        if self.arena is not None and self.pool is not None:
            self.pool.append(self.arena)
        self.arena = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # -- views for tests / the slow merge path (no copies) -------------------------------------------
    def _view(self, ptr: int, shape, dtype) -> torch.Tensor:
        off = ptr - self.arena.data_ptr()
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.arena[off:off + nbytes].view(dtype).view(shape)

    @property
    def n_unique(self) -> torch.Tensor:
        return self._view(self.c.n_unique, (1,), torch.int32)

    @property
    def uniq(self) -> torch.Tensor:
        return self._view(self.c.uniq, (self.c.cap,), torch.int32)

    def rows(self, which: str = "G") -> torch.Tensor:
        return self._view(getattr(self.c, which), (self.c.cap, self.c.H), torch.float32)


class FactoredEngine(EmbeddingEngine):
    """EmbeddingEngine whose forward/backward run the factored kernels. ``dnn``: {'item': nn.Linear, 'user': nn.Linear}
    (itemdnn / userdnn, model.py:150-151) — borrowed like the tables."""

    def __init__(self, layout: FeatureLayout, tables, mm, dnn: Dict[str, torch.nn.Linear], mode: str = "fused",
                 check_shapes: bool = True):
        super().__init__(layout, tables, mm, mode, check_shapes)
        if layout.H not in SUPPORTED_H:
            raise ValueError(f"the factored path supports hidden_units in {SUPPORTED_H}, got {layout.H}")
        self.dnn = dnn
        self.path = "factored"
        self.current: Optional[FactGroup] = None   # group made by prefetch(), consumed by the next forwards
        self.ready: List[FactGroup] = []           # groups whose row gradients wait for fused_step
        full = layout.calls[True]
        self._prm = _lib.FactParams()
        seen = set()
        for s in full.slots:
            if s.kind == KIND_MM:
                continue
            if s.table in seen:
                raise ValueError("a table feeding two slots cannot be factored")
            seen.add(s.table)
            self._prm.dnn.table_side[s.table] = s.side
            self._prm.dnn.table_col[s.table] = s.col
        self._mm_slots = [s for s in full.slots if s.kind == KIND_MM]
        if len(self._mm_slots) > _lib.MAX_MM:
            raise ValueError(f"at most {_lib.MAX_MM} mm features")
        self._prm.n_mm = len(self._mm_slots)
        for f, s in enumerate(self._mm_slots):
            self._prm.mm[f].mm_dim, self._prm.mm[f].col = s.mm_dim, s.col
        self._prm_key = None
        self._prm_fresh = False
        self._tab_memo: Dict[bool, object] = {}
        self._arena_pool: List[torch.Tensor] = []
        self.adam_dev: Optional[torch.Tensor] = None   # device copy of the tgr_adam_t block (set by graphed.GraphedStep)
        # fused mode may also own the path's Linear layers (itemdnn / userdnn / emb_transform): their gradients then stay in
        # the group's accumulators (no autograd AccumulateGrad copies) and fused_step(dense=True) updates them in one launch
        self.own_dense = False
        self._staged: Dict[tuple, FactGroup] = {}
        self.reduce_done_event = None   # torch.cuda.Event recorded after the finishing backward's segmented reduce
        self.overlap_mm = True      # prefetch(): mm branch on a side stream (tgr_fact_mm_branch)
        self._dense_state: Dict[int, tuple] = {}

    # ------------------------------------------------------------------ structs: pointers are re-read once per group
    def _table_array(self, state: bool = False, grads=None):
        """Parameter pointers are re-validated when a group is prepared (and at the row update), not on each of the
        ~8 C calls of a step: reading 24 ``.data_ptr()`` per call was 0.2 ms of host time per step."""
        if grads is not None:
            return super()._table_array(state, grads)
        hit = self._tab_memo.get(state)
        if hit is None:
            hit = self._tab_memo[state] = super()._table_array(state)
        return hit

    def _params(self) -> "_lib.FactParams":
        if self._prm_fresh:
            return self._prm
        self._prm_fresh = True
        p = self._prm
        wi, wu = self.dnn["item"].weight.data, self.dnn["user"].weight.data
        bi, bu = self.dnn["item"].bias.data, self.dnn["user"].bias.data
        key = (wi.data_ptr(), wu.data_ptr(), bi.data_ptr(), bu.data_ptr()) + tuple(
            self.mm[s.name].weight.data.data_ptr() for s in self._mm_slots)
        if key == self._prm_key:
            return p
        for w in (wi, wu, bi, bu):
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise TypeError("itemdnn / userdnn parameters must be contiguous float32")
        p.dnn.w_item, p.dnn.item_ld = wi.data_ptr(), wi.stride(0)
        p.dnn.w_user, p.dnn.user_ld = wu.data_ptr(), wu.stride(0)
        p.b_item, p.b_user = bi.data_ptr(), bu.data_ptr()
        for f, s in enumerate(self._mm_slots):
            lin = self.mm[s.name]
            w = lin.weight.data
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise TypeError("emb_transform weight must be contiguous float32")
            p.mm[f].w = w.data_ptr()
            p.mm[f].b = None if lin.bias is None else lin.bias.data.data_ptr()
        self._prm_key = key
        return p

    # ------------------------------------------------------------------ group preparation (value independent)
    def _describe(self, pbs: Sequence[PackedBatch]):
        """-> (group with its C descriptor filled from the calls, arena bytes it needs)."""
        lay = self.layout
        g = FactGroup(pbs)
        if len(g.pbs) > _lib.MAX_CALLS:
            raise ValueError(f"at most {_lib.MAX_CALLS} calls per group")
        c = g.c
        c.n_calls, c.H, c.key_bits, c.n_mm = len(g.pbs), lay.H, lay.key_bits, len(self._mm_slots)
        if all(pb.n_cap is not None for pb in g.pbs):
            # fixed-shape calls: nothing the launch sequence depends on may come from the buffers' CONTENT
            g.n = c.n = sum(pb.n_cap for pb in g.pbs)
            c.n_is_capacity = 1
        else:
            g.n = c.n = sum(pb.n_valid for pb in g.pbs)
            c.n_is_capacity = 0
        x_dt = None
        for f, s in enumerate(self._mm_slots):
            c.mm_dim[f] = s.mm_dim
        for i, pb in enumerate(g.pbs):
            cl = lay.calls[pb.include_user]
            if pb.ids.dtype != torch.int32 or not pb.ids.is_contiguous() or tuple(pb.ids.shape) != (pb.T, cl.n_single):
                raise TypeError("PackedBatch.ids must be contiguous int32 [T, n_single]")
            self._structs.call(pb, pb.ids, None, _lib.DTYPE_F32, None, out=c.calls[i])
            c.calls[i].item_cat = None
            for f, s in enumerate(self._mm_slots):
                x = pb.mm_x[s.src]
                if not x.is_contiguous() or tuple(x.shape) != (pb.T, s.mm_dim):
                    raise TypeError(f"mm input {s.name} must be contiguous [T, {s.mm_dim}]")
                if x_dt is None:
                    x_dt = x.dtype
                elif x.dtype != x_dt:
                    raise TypeError("all mm inputs of a group must share one dtype")
                c.mm_x[i][f] = x.data_ptr()
        c.mm_x_dtype = _lib.DTYPE_F32 if x_dt is None else self._dt_code(x_dt)
        nbytes = self.lib.tgr_fact_group_bytes(C.byref(c), len(self.tables))
        if nbytes == 0:
            check(-1, "tgr_fact_group_bytes")
        return g, int(nbytes)

    def prepare(self, pbs: Sequence[PackedBatch], arena: Optional[torch.Tensor] = None) -> FactGroup:
        """keys -> sort -> dedup -> id remap for the calls of one group (ONE C call). Independent of table VALUES.
        ``arena``: carve the group from this uint8 buffer (kept by the caller) instead of the engine's recycling pool."""
        self._require_cuda()
        self._tab_memo.clear()       # re-read the parameter pointers for this group
        self._prm_fresh = False
        dev = self._device()
        g, nbytes = self._describe(pbs)
        c, nt = g.c, len(self.tables)
        if arena is not None:
            if arena.numel() < nbytes or arena.dtype != torch.uint8 or arena.device != dev:
                raise ValueError(f"prepare: the arena must be a uint8 tensor of >= {nbytes} bytes on {dev}")
            g.arena, g.pool = arena, None
        else:
            g.arena = self._take_arena(nbytes, dev)
            g.pool = self._arena_pool
        e0 = self._t0()
        check(self.lib.tgr_fact_prepare(self._table_array(), nt, C.byref(c), g.arena.data_ptr(), nbytes, _stream()),
              "tgr_fact_prepare")
        self._t1("fact_prepare", e0)
        self.launches += 13 + 2 * len(g.pbs) + sum(1 for pb in g.pbs for z in pb.arr_nnz if z)
        if _DEBUG and not c.n_is_capacity:
            got = int(g._view(c.n_valid, (1,), torch.int32).item())
            if got != g.n:
                raise _lib.TgrError(f"PackedBatch.n_valid mismatch: host {g.n}, device {got}")
        return g

    def _take_arena(self, nbytes: int, dev) -> torch.Tensor:
        """Arenas (~1.7 GB at C2) are recycled instead of going through the allocator every step."""
        pool = self._arena_pool
        best = None
        for i, a in enumerate(pool):
            if a.numel() >= nbytes and a.device == dev and (best is None or a.numel() < pool[best].numel()):
                best = i
        if best is not None:
            return pool.pop(best)
        if len(pool) >= 4:          # too-small leftovers: let the allocator have them back
            pool.clear()
        return torch.empty(int(nbytes * 1.15) + 4096, dtype=torch.uint8, device=dev)

    @staticmethod
    def _dt_code(dt: torch.dtype) -> int:
        from .engine import _dtype_code
        return _dtype_code(dt)

    def group_bytes(self, pbs: Sequence[PackedBatch]) -> int:
        """Arena bytes a group of these calls needs (for callers that keep their own arenas)."""
        return self._describe(pbs)[1]

    def stage(self, g: FactGroup):
        """Hand over a group prepared ahead of time (its key processing already enqueued, e.g. next to the previous step's
        backward): the ``prefetch`` of the same calls picks it up instead of preparing them again."""
        self._staged[tuple(id(pb) for pb in g.pbs)] = g

    def prefetch(self, pbs: Sequence[PackedBatch]) -> FactGroup:
        g = self._staged.pop(tuple(id(pb) for pb in pbs), None)
        early = g is not None
        if g is None:
            g = self.prepare(pbs)
        else:
            self._tab_memo.clear()       # parameter pointers are re-read for the step that consumes the group
            self._prm_fresh = False
        self.current = g
        if self._mm_slots and self.overlap_mm:
            # the forwards follow at once: fold + mm projection of every call on a side stream next to the key processing
            check(self.lib.tgr_fact_mm_branch(C.byref(self._params()), C.byref(g.c), 1 if early else 0, _stream()),
                  "tgr_fact_mm_branch")
            self.launches += len(self._mm_slots) * (1 + len(g.pbs))
        return g

    def _group_of(self, pb: PackedBatch) -> FactGroup:
        g = self.current
        if g is not None and id(pb) in g.index:
            return g
        return self.prepare([pb])

    # ------------------------------------------------------------------ forward of one call
    def fact_forward(self, g: FactGroup, pb: PackedBatch) -> torch.Tensor:
        """-> out [T, H] fp32 (the ReLU masks stay in the group's arena for the backward)."""
        H = self.layout.H
        out = torch.empty((pb.T, H), dtype=torch.float32, device=self._device())
        n_kern = 1 + len(self._mm_slots) + (0 if g.c.projected else 1 + len(self._mm_slots))
        e0 = self._t0()
        check(self.lib.tgr_fact_call_forward(self._table_array(), len(self.tables), C.byref(self._params()), C.byref(g.c),
                                             g.index[id(pb)], out.data_ptr(), _stream()), "tgr_fact_call_forward")
        self._t1("fact_call_forward", e0)
        self.launches += n_kern
        return out

    # ------------------------------------------------------------------ backward of one call
    def _acc(self, g: FactGroup) -> dict:
        """Zero-initialised gradient accumulators of the group's dense parameters: ONE flat buffer, viewed."""
        if g.acc is None:
            dev = self._device()
            shapes = [("dW_item", self.dnn["item"].weight.shape), ("db_item", self.dnn["item"].bias.shape)]
            if any(pb.include_user for pb in g.pbs):
                shapes += [("dW_user", self.dnn["user"].weight.shape), ("db_user", self.dnn["user"].bias.shape)]
            for s in self._mm_slots:
                lin = self.mm[s.name]
                shapes.append((f"dWmm/{s.name}", lin.weight.shape))
                if lin.bias is not None:
                    shapes.append((f"dbmm/{s.name}", lin.bias.shape))
            sizes = [(k, sh, (sh.numel() + 63) // 64 * 64) for k, sh in shapes]
            flat = torch.zeros(sum(z for _, _, z in sizes), dtype=torch.float32, device=dev)
            a, o = {"dW_user": None, "db_user": None}, 0
            for k, sh, z in sizes:
                a[k] = flat[o:o + sh.numel()].view(sh)
                o += z
            cg = _lib.FactGrads()
            cg.dW_item, cg.db_item = a["dW_item"].data_ptr(), a["db_item"].data_ptr()
            if a["dW_user"] is not None:
                cg.dW_user, cg.db_user = a["dW_user"].data_ptr(), a["db_user"].data_ptr()
            for f, s in enumerate(self._mm_slots):
                cg.dW_mm[f] = a[f"dWmm/{s.name}"].data_ptr()
                b = a.get(f"dbmm/{s.name}")
                cg.db_mm[f] = None if b is None else b.data_ptr()
            a["c"] = cg
            g.acc = a
        return g.acc

    def fact_backward(self, g: FactGroup, pb: PackedBatch, d_out: torch.Tensor) -> bool:
        """dZ + bias / mm gradients of one call; with the group's last gradient also the segmented reduce, the row
        gradients and the DNN weight gradients (ONE C call). True when the group is complete."""
        H = self.layout.H
        a = self._acc(g)
        d_out = d_out.reshape(pb.T, H)
        if d_out.dtype != torch.float32:
            d_out = d_out.float()
        if not d_out.is_contiguous():
            d_out = d_out.contiguous()
        g.n_bwd += 1
        finish = g.n_bwd == g.n_fwd
        if finish and g.n_bwd != len(g.pbs):
            raise RuntimeError("a prefetched group needs the gradient of every one of its calls "
                               f"({g.n_bwd} of {len(g.pbs)} arrived); prefetch only the calls that reach the loss")
        n_mm = len(self._mm_slots)
        ev = self.reduce_done_event
        g.c.reduce_done_event = ev.cuda_event if (finish and ev is not None) else None
        e0 = self._t0()
        check(self.lib.tgr_fact_call_backward(self._table_array(), len(self.tables), C.byref(self._params()), C.byref(g.c),
                                              g.index[id(pb)], d_out.data_ptr(), C.byref(a["c"]), 1 if finish else 0,
                                              _stream()), "tgr_fact_call_backward")
        self._t1("fact_call_backward", e0)
        self.launches += 2 + 3 * n_mm + (4 if finish and g.n else 0)
        if finish:
            g.done = True
            if self.current is g:
                self.current = None
        return finish

    def dense_from_rows(self, g: FactGroup) -> List[Optional[torch.Tensor]]:
        """Parity mode: dense [rows, H] gradients of the tables the group's calls index."""
        lay = self.layout
        touched = sorted({s.table for pb in g.pbs for s in lay.calls[pb.include_user].slots if s.table >= 0})
        grads: List[Optional[torch.Tensor]] = [None] * len(self.tables)
        for t in touched:
            grads[t] = torch.zeros_like(self.tables[t].data)
        if g.n:
            tabs = self._table_array(grads=grads)
            check(self.lib.tgr_scatter_rows(tabs, len(self.tables), lay.H, g.c.uniq, g.c.G, g.c.n_unique, g.c.cap, _stream()),
                  "tgr_scatter_rows")
            self.launches += 1
        return grads

    def discard_pending(self) -> int:
        """Drop the groups whose row gradients wait for a row update that will not happen (skipped optimizer step)."""
        groups, self.ready = self.ready, []
        for g in groups:
            g.release()
        return len(groups) + super().discard_pending()

    # ------------------------------------------------------------------ row update
    def _dense_update(self, g: FactGroup, adam):
        """AdamW on itemdnn / userdnn / emb_transform from the group's gradient accumulators (tgr_adam_dense)."""
        a = g.acc
        pairs = [(self.dnn["item"].weight, a["dW_item"]), (self.dnn["item"].bias, a["db_item"])]
        if a.get("dW_user") is not None:
            pairs += [(self.dnn["user"].weight, a["dW_user"]), (self.dnn["user"].bias, a["db_user"])]
        for s in self._mm_slots:
            lin = self.mm[s.name]
            pairs.append((lin.weight, a[f"dWmm/{s.name}"]))
            if lin.bias is not None:
                pairs.append((lin.bias, a[f"dbmm/{s.name}"]))
        if len(pairs) > _lib.MAX_DENSE:
            raise ValueError(f"more than {_lib.MAX_DENSE} dense tensors")
        dl = _lib.DenseList()
        dl.n = len(pairs)
        for i, (p, gr) in enumerate(pairs):
            st = self._dense_state.get(id(p))
            if st is None or st[0].shape != p.shape or st[0].device != p.device:
                st = self._dense_state[id(p)] = (torch.zeros_like(p.data), torch.zeros_like(p.data))
            if p.dtype != torch.float32 or not p.data.is_contiguous():
                raise TypeError("dense parameters must be contiguous float32")
            dl.w[i], dl.g[i], dl.m[i], dl.v[i], dl.numel[i] = p.data.data_ptr(), gr.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), p.numel()
        e0 = self._t0()
        if self.adam_dev is not None:
            check(self.lib.tgr_adam_dense(C.byref(dl), None, self.adam_dev.data_ptr(), _stream()), "tgr_adam_dense")
        else:
            check(self.lib.tgr_adam_dense(C.byref(dl), C.addressof(adam), None, _stream()), "tgr_adam_dense")
        self._t1("adam_dense", e0)
        self.launches += 1

    def fused_step(self, lr: float = 1e-3, betas=(0.9, 0.98), eps: float = 1e-8, weight_decay: float = 1e-2,
                   grad_scale: float = 1.0, dense: bool = False):
        """``dense=True`` (needs ``own_dense``): also AdamW-update the path's Linear layers with the same hyper-parameters."""
        groups, self.ready = self.ready, []
        if not groups:
            return 0
        if dense and (not self.own_dense or len(groups) != 1):
            raise RuntimeError("fused_step(dense=True) needs engine.own_dense and the prefetch protocol (one group per step)")
        self._require_cuda()
        self.ensure_state()
        self._tab_memo.pop(True, None)
        self.step += 1
        H, dev = self.layout.H, self._device()
        adam = make_adam(lr, betas[0], betas[1], eps, weight_decay, self.step, grad_scale)
        tabs = self._table_array(state=True)
        if len(groups) == 1:
            g = groups[0]
            if g.n:
                e0 = self._t0()
                if self.adam_dev is not None:
                    # hyper-parameter block in device memory (graphed.GraphedStep refreshes it before every replay)
                    check(self.lib.tgr_adam_rows_dev(tabs, len(self.tables), H, g.c.uniq, g.c.G, g.c.n_unique, g.c.cap,
                                                     self.adam_dev.data_ptr(), _stream()), "tgr_adam_rows_dev")
                else:
                    check(self.lib.tgr_adam_rows(tabs, len(self.tables), H, g.c.uniq, g.c.G, g.c.n_unique, g.c.cap,
                                                 C.byref(adam), _stream()), "tgr_adam_rows")
                self._t1("adam_rows", e0)
                self.launches += 1
            if dense and g.acc is not None:
                self._dense_update(g, adam)
            g.release()
            return g.n
        # several groups touched the step (per-call protocol): merge their (key, row gradient) lists in group order,
        # then ONE update per row. Slow path: one host read of the unique counts.
        if len(groups) > _lib.MAX_CALLS:
            raise ValueError(f"more than {_lib.MAX_CALLS} feat2emb groups queued for one optimizer step")
        counts = torch.cat([g.n_unique for g in groups]).tolist()
        keys = torch.cat([g.uniq[:c] for g, c in zip(groups, counts)])
        srcs = torch.cat([torch.arange(c, dtype=torch.int32, device=dev) + (i << 29) for i, c in enumerate(counts)])
        R = int(keys.numel())
        if R == 0:
            return 0
        keys_o, srcs_o = torch.empty_like(keys), torch.empty_like(srcs)
        ws = self._buf("keys_ws", self.lib.tgr_sort_workspace_bytes(R), dev)
        check(self.lib.tgr_sort_pairs(keys.data_ptr(), srcs.data_ptr(), keys_o.data_ptr(), srcs_o.data_ptr(), R,
                                      self.layout.key_bits, ws.data_ptr(), ws.numel(), _stream()), "tgr_sort_pairs")
        structs = (Call * len(groups))()
        for i, (g, c) in enumerate(zip(groups, counts)):
            cs = structs[i]
            cs.T, cs.n_slots, cs.n_single = max(c, 1), 1, 1
            cs.slots[0].kind, cs.slots[0].side, cs.slots[0].col, cs.slots[0].table, cs.slots[0].src = 0, 0, 0, 0, 0
            cs.item_cat, cs.item_ld, cs.cat_dtype = g.c.G, H, _lib.DTYPE_F32
        rws = self._buf("reduce_ws", self.lib.tgr_reduce_workspace_bytes(R, H), dev)
        check(self.lib.tgr_bwd_reduce(tabs, len(self.tables), H, structs, len(groups), keys_o.data_ptr(), srcs_o.data_ptr(), R,
                                      1, None, None, C.byref(adam), rws.data_ptr(), rws.numel(), _stream()), "tgr_bwd_reduce")
        self.launches += 8
        for g in groups:
            g.release()
        return R


class FactoredFn(torch.autograd.Function):
    """out = feat2emb(packed call) through the factored kernels. Inputs: tables, emb_transform (W, b) pairs,
    itemdnn (W, b), userdnn (W, b) — so autograd routes gradients where the reference's graph would."""

    @staticmethod
    def forward(ctx, engine: FactoredEngine, pb: PackedBatch, needs_grad: bool, *params):
        g = engine._group_of(pb)
        out = engine.fact_forward(g, pb)
        ctx.engine, ctx.pb, ctx.group = engine, pb, g
        ctx.n_params = len(params)
        if needs_grad:      # (grad mode is always off inside Function.forward: the caller tells)
            g.n_fwd += 1
        # [T, H], NOT a view: the reference scales feat2emb's result in place (`seqs *= ...`, model.py:325) and autograd
        # refuses in-place writes on a view created inside a custom Function; the caller reshapes outside
        return out

    @staticmethod
    def backward(ctx, d_out):
        eng, pb, g = ctx.engine, ctx.pb, ctx.group
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        if not eng.fact_backward(g, pb, d_out):
            return (None, None, None, *grads)
        a = g.acc
        n_t = len(eng.tables)
        names = list(eng.layout.item_emb_feat)
        if eng.mode == "fused":
            eng.ready.append(g)
            if eng.own_dense:       # the Linear gradients stay in g.acc for fused_step(dense=True)
                return (None, None, None, *grads)
        else:
            dense = eng.dense_from_rows(g)
            for i in range(n_t):
                if ctx.needs_input_grad[3 + i]:
                    grads[i] = dense[i]
            g.release()
        for j, k in enumerate(names):
            grads[n_t + 2 * j] = a[f"dWmm/{k}"]
            grads[n_t + 2 * j + 1] = a.get(f"dbmm/{k}")
        o = n_t + 2 * len(names)
        grads[o], grads[o + 1] = a["dW_item"], a["db_item"]
        grads[o + 2], grads[o + 3] = a["dW_user"], a["db_user"]
        return (None, None, None, *grads)
