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
from .engine import EmbeddingEngine, _stream
from .layout import KIND_ARRAY, KIND_MM, KIND_SINGLE, SIDE_ITEM, FeatureLayout
from .packed import PackedBatch

SUPPORTED_H = (32, 64, 128)


class FactGroup:
    """Everything the calls prepared together share; owns its buffers until the row update has run."""

    def __init__(self, pbs: Sequence[PackedBatch]):
        self.pbs = list(pbs)
        self.index = {id(pb): i for i, pb in enumerate(self.pbs)}
        self.n = 0
        self.pairs = None          # keeps the sorted (key, src) storage alive
        self.keys = self.srcs = 0  # device pointers into ``pairs``
        self.uniq = self.seg_of = self.n_unique = None
        self.cap = 1
        self.ids_u: List[torch.Tensor] = []
        self.arr_u: List[torch.Tensor] = []
        self.P: Optional[torch.Tensor] = None
        self.fold: Dict[str, tuple] = {}
        self.n_fwd = 0             # forwards that will get a gradient
        self.n_bwd = 0
        self.dz: Dict[int, tuple] = {}
        self.acc: Optional[dict] = None   # dW_item, dW_user, db_item, db_user, dWmm{}, dbmm{}
        self.g_rows: Optional[torch.Tensor] = None
        self.done = False


class FactoredEngine(EmbeddingEngine):
    """EmbeddingEngine whose forward/backward run the factored kernels. ``dnn``: {'item': nn.Linear, 'user': nn.Linear}
    (itemdnn / userdnn, model.py:150-151) — borrowed like the tables."""

    def __init__(self, layout: FeatureLayout, tables, mm, dnn: Dict[str, torch.nn.Linear], mode: str = "fused"):
        super().__init__(layout, tables, mm, mode)
        if layout.H not in SUPPORTED_H:
            raise ValueError(f"the factored path supports hidden_units in {SUPPORTED_H}, got {layout.H}")
        self.dnn = dnn
        self.path = "factored"
        self.current: Optional[FactGroup] = None   # group made by prefetch(), consumed by the next forwards
        self.ready: List[FactGroup] = []           # groups whose row gradients wait for fused_step
        full = layout.calls[True]
        self._dnn_struct = Dnn()
        seen = set()
        for s in full.slots:
            if s.kind == KIND_MM:
                continue
            if s.table in seen:
                raise ValueError("a table feeding two slots cannot be factored")
            seen.add(s.table)
            self._dnn_struct.table_side[s.table] = s.side
            self._dnn_struct.table_col[s.table] = s.col
        self._mm_slots = [s for s in full.slots if s.kind == KIND_MM]
        # call templates with every slot at column 0 of its side: the "concat gradient" of the reduction is dZ [T, H]
        self._call0: Dict[bool, Call] = {}
        for inc in (True, False):
            c = Call()
            C.memmove(C.addressof(c), C.addressof(self._structs._call_tmpl[inc]), C.sizeof(Call))
            for i in range(c.n_slots):
                c.slots[i].col = 0
            self._call0[inc] = c

    # ------------------------------------------------------------------ structs
    def _dnn(self) -> Dnn:
        d = self._dnn_struct
        wi = self.dnn["item"].weight.data
        wu = self.dnn["user"].weight.data
        for w in (wi, wu):
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise TypeError("itemdnn / userdnn weights must be contiguous float32")
        d.w_item, d.item_ld = wi.data_ptr(), wi.stride(0)
        d.w_user, d.user_ld = wu.data_ptr(), wu.stride(0)
        return d

    def _dz_call(self, pb: PackedBatch, dz_item: torch.Tensor, dz_user: Optional[torch.Tensor], out: Call) -> Call:
        C.memmove(C.addressof(out), C.addressof(self._call0[pb.include_user]), C.sizeof(Call))
        out.T = pb.T
        out.item_cat, out.item_ld = dz_item.data_ptr(), self.layout.H
        if dz_user is not None:
            out.user_cat, out.user_ld = dz_user.data_ptr(), self.layout.H
        out.cat_dtype = _lib.DTYPE_F32
        return out

    # ------------------------------------------------------------------ group preparation (value independent)
    def prepare(self, pbs: Sequence[PackedBatch]) -> FactGroup:
        """keys -> sort -> dedup -> id remap for the calls of one group. Independent of table VALUES."""
        self._require_cuda()
        lay, dev = self.layout, self._device()
        g = FactGroup(pbs)
        if len(g.pbs) > _lib.MAX_CALLS:
            raise ValueError(f"at most {_lib.MAX_CALLS} calls per group")
        calls = []
        for pb in g.pbs:
            cl = lay.calls[pb.include_user]
            di = torch.empty((0, cl.item_dim), device=dev)
            du = torch.empty((0, max(cl.user_dim, 1)), device=dev) if pb.include_user else None
            calls.append((pb, di, du))
        structs, g.keys, g.srcs, g.n, _ = self._sorted_pairs(calls)
        g.pairs = self._ws.pop("pairs")
        n = g.n
        g.cap = cap = max(n, 1)
        g.uniq = torch.empty(cap, dtype=torch.int32, device=dev)
        seg_off = torch.empty(cap + 1, dtype=torch.int32, device=dev)
        g.seg_of = torch.empty(cap, dtype=torch.int32, device=dev)
        g.n_unique = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._buf("dedup_ws", self.lib.tgr_dedup_workspace_bytes(n), dev)
        e0 = self._t0()
        check(self.lib.tgr_dedup(g.keys, n, g.uniq.data_ptr(), seg_off.data_ptr(), g.seg_of.data_ptr(),
                                 g.n_unique.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "tgr_dedup")
        self._t1("dedup", e0)
        self.launches += 4
        # ids -> 1 + unique index: one scatter over the sorted pairs for every SINGLE slot of every call
        e0 = self._t0()
        ptrs = (C.c_void_p * len(g.pbs))()
        for i, pb in enumerate(g.pbs):
            o = torch.zeros_like(pb.ids)
            g.ids_u.append(o)
            ptrs[i] = o.data_ptr()
        if n:
            check(self.lib.tgr_remap_scatter(g.srcs, g.seg_of.data_ptr(), n, None, structs, len(g.pbs), ptrs, _stream()),
                  "tgr_remap_scatter")
            self.launches += 1
        for pb in g.pbs:   # array values (a token may hold several): searching remap, they are few
            cl = lay.calls[pb.include_user]
            arr_r = torch.zeros_like(pb.arr_val)
            for s in cl.slots:
                if s.kind != KIND_ARRAY or pb.arr_nnz[s.src] == 0:
                    continue
                kb1 = (C.c_uint32 * 1)(lay.tables[s.table].key_base)
                rw1 = (C.c_int32 * 1)(lay.tables[s.table].rows)
                off = 4 * pb.arr_begin[s.src]
                check(self.lib.tgr_remap_ids(pb.arr_val.data_ptr() + off, pb.arr_nnz[s.src], 1, kb1, rw1, g.uniq.data_ptr(),
                                             g.n_unique.data_ptr(), None, arr_r.data_ptr() + off, _stream()),
                      "tgr_remap_ids(array)")
                self.launches += 1
            g.arr_u.append(arr_r)
        self._t1("remap", e0)
        return g

    def prefetch(self, pbs: Sequence[PackedBatch]) -> FactGroup:
        self.current = self.prepare(pbs)
        return self.current

    def _group_of(self, pb: PackedBatch) -> FactGroup:
        g = self.current
        if g is not None and id(pb) in g.index:
            return g
        return self.prepare([pb])

    # ------------------------------------------------------------------ value dependent: projection + mm fold
    def _project(self, g: FactGroup):
        lay, dev, H = self.layout, self._device(), self.layout.H
        g.P = torch.empty((g.cap, H), dtype=torch.float32, device=dev)
        tabs = self._table_array()
        e0 = self._t0()
        check(self.lib.tgr_fact_project_rows(tabs, len(self.tables), H, C.byref(self._dnn()), g.uniq.data_ptr(),
                                             g.n_unique.data_ptr(), g.cap, g.P.data_ptr(), _stream()),
              "tgr_fact_project_rows")
        self._t1("fact_project_rows", e0)
        self.launches += 1
        wi = self.dnn["item"].weight.data
        for s in self._mm_slots:
            lin = self.mm[s.name]
            w = lin.weight.data
            b = lin.bias.data if lin.bias is not None else None
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise TypeError("emb_transform weight must be contiguous float32")
            M = torch.empty((H, s.mm_dim), dtype=torch.float32, device=dev)
            c = torch.empty((H,), dtype=torch.float32, device=dev)
            check(self.lib.tgr_fact_mm_fold(wi.data_ptr() + 4 * s.col, wi.stride(0), w.data_ptr(),
                                            None if b is None else b.data_ptr(), H, s.mm_dim, M.data_ptr(), c.data_ptr(),
                                            _stream()), "tgr_fact_mm_fold")
            self.launches += 1
            g.fold[s.name] = (M, c)

    # ------------------------------------------------------------------ forward of one call
    def fact_forward(self, g: FactGroup, pb: PackedBatch):
        """-> (out [T, H] fp32, mask [T, H/4] uint8)."""
        lay, dev, H = self.layout, self._device(), self.layout.H
        if g.P is None:
            self._project(g)
        i = g.index[id(pb)]
        T = pb.T
        cl = lay.calls[pb.include_user]
        out = torch.empty((T, H), dtype=torch.float32, device=dev)
        mask = torch.empty((T, H // 4), dtype=torch.uint8, device=dev)
        mmz = []
        for s in cl.slots:
            if s.kind != KIND_MM:
                continue
            x = pb.mm_x[s.src]
            if not x.is_contiguous() or x.shape != (T, s.mm_dim):
                raise TypeError(f"mm input {s.name} must be contiguous [T, {s.mm_dim}]")
            M, c = g.fold[s.name]
            z = torch.empty((T, H), dtype=torch.float32, device=dev)
            e0 = self._t0()
            check(self.lib.tgr_mm_proj_fwd(x.data_ptr(), self._dt(x), T, s.mm_dim, M.data_ptr(), c.data_ptr(), H, z.data_ptr(),
                                           H, _lib.DTYPE_F32, _stream()), "tgr_mm_proj_fwd")
            self._t1("mm_proj_fwd", e0)
            self.launches += 1
            mmz.append(z)
        mm_ptrs = (C.c_void_p * max(len(mmz), 1))(*[z.data_ptr() for z in mmz])
        call = self._structs.call(pb, out, None, _lib.DTYPE_F32, None)
        bi = self.dnn["item"].bias.data
        bu = self.dnn["user"].bias.data if pb.include_user else None
        e0 = self._t0()
        check(self.lib.tgr_fact_forward(C.byref(call), H, g.ids_u[i].data_ptr(),
                                        g.arr_u[i].data_ptr() if g.arr_u[i].numel() else None, g.P.data_ptr(), mm_ptrs,
                                        len(mmz), bi.data_ptr(), None if bu is None else bu.data_ptr(), out.data_ptr(),
                                        mask.data_ptr(), _stream()), "tgr_fact_forward")
        self._t1("fact_forward", e0)
        self.launches += 1
        return out, mask

    @staticmethod
    def _dt(x: torch.Tensor) -> int:
        from .engine import _dtype_code
        return _dtype_code(x.dtype)

    # ------------------------------------------------------------------ backward of one call
    def _acc(self, g: FactGroup) -> dict:
        if g.acc is None:
            dev = self._device()
            a = {"dW_item": torch.zeros_like(self.dnn["item"].weight.data),
                 "db_item": torch.zeros_like(self.dnn["item"].bias.data),
                 "dW_user": None, "db_user": None, "dWmm": {}, "dbmm": {}}
            if any(pb.include_user for pb in g.pbs):
                a["dW_user"] = torch.zeros_like(self.dnn["user"].weight.data)
                a["db_user"] = torch.zeros_like(self.dnn["user"].bias.data)
            for s in self._mm_slots:
                lin = self.mm[s.name]
                a["dWmm"][s.name] = torch.zeros_like(lin.weight.data)
                a["dbmm"][s.name] = torch.zeros_like(lin.bias.data) if lin.bias is not None else None
            g.acc = a
        return g.acc

    def fact_backward(self, g: FactGroup, pb: PackedBatch, mask: torch.Tensor, d_out: torch.Tensor) -> bool:
        """dZ + bias / mm gradients of one call; True when the group is now complete (group_backward may run)."""
        lay, dev, H = self.layout, self._device(), self.layout.H
        i = g.index[id(pb)]
        T = pb.T
        a = self._acc(g)
        d_out = d_out.reshape(T, H)
        if d_out.dtype != torch.float32:
            d_out = d_out.float()
        if not d_out.is_contiguous():
            d_out = d_out.contiguous()
        dz_item = torch.empty((T, H), dtype=torch.float32, device=dev)
        dz_user = torch.empty((T, H), dtype=torch.float32, device=dev) if pb.include_user else None
        ws = self._buf("relu_ws", self.lib.tgr_fact_relu_mask_workspace_bytes(T, H), dev)
        e0 = self._t0()
        check(self.lib.tgr_fact_relu_mask(d_out.data_ptr(), mask.data_ptr(), T, H, dz_item.data_ptr(),
                                          None if dz_user is None else dz_user.data_ptr(), a["db_item"].data_ptr(),
                                          None if dz_user is None else a["db_user"].data_ptr(), ws.data_ptr(), ws.numel(),
                                          _stream()), "tgr_fact_relu_mask")
        self._t1("fact_relu_mask", e0)
        self.launches += 2
        wi = self.dnn["item"].weight.data
        for s in lay.calls[pb.include_user].slots:
            if s.kind != KIND_MM:
                continue
            x = pb.mm_x[s.src]
            lin = self.mm[s.name]
            A = torch.empty((H, s.mm_dim), dtype=torch.float32, device=dev)
            sv = torch.empty((H,), dtype=torch.float32, device=dev)
            mws = self._buf("mm_bwd", self.lib.tgr_mm_proj_bwd_workspace_bytes(T, s.mm_dim, H), dev)
            e0 = self._t0()
            check(self.lib.tgr_mm_proj_bwd(x.data_ptr(), self._dt(x), T, s.mm_dim, dz_item.data_ptr(), H, _lib.DTYPE_F32, H,
                                           A.data_ptr(), sv.data_ptr(), 0, mws.data_ptr(), mws.numel(), _stream()),
                  "tgr_mm_proj_bwd")
            dbm = a["dbmm"][s.name]
            check(self.lib.tgr_fact_mm_chain_bwd(wi.data_ptr() + 4 * s.col, wi.stride(0), lin.weight.data.data_ptr(),
                                                 None if lin.bias is None else lin.bias.data.data_ptr(), A.data_ptr(),
                                                 sv.data_ptr(), H, s.mm_dim, a["dWmm"][s.name].data_ptr(),
                                                 None if dbm is None else dbm.data_ptr(),
                                                 a["dW_item"].data_ptr() + 4 * s.col, a["dW_item"].stride(0), _stream()),
                  "tgr_fact_mm_chain_bwd")
            self._t1("mm_proj_bwd", e0)
            self.launches += 3
        g.dz[i] = (dz_item, dz_user)
        g.n_bwd += 1
        return g.n_bwd == g.n_fwd

    def group_backward(self, g: FactGroup):
        """Segmented sum of dZ rows per unique key, row gradients and DNN weight gradients of the whole group."""
        lay, dev, H = self.layout, self._device(), self.layout.H
        if len(g.dz) != len(g.pbs):
            raise RuntimeError("a prefetched group needs the gradient of every one of its calls "
                               f"({len(g.dz)} of {len(g.pbs)} arrived); prefetch only the calls that reach the loss")
        a = self._acc(g)
        n = g.n
        G = torch.empty((g.cap, H), dtype=torch.float32, device=dev)
        if n:
            structs = (Call * len(g.pbs))()
            for i, pb in enumerate(g.pbs):
                self._dz_call(pb, g.dz[i][0], g.dz[i][1], structs[i])
            rws = self._buf("reduce_ws", self.lib.tgr_reduce_workspace_bytes(n, H), dev)
            tabs = self._table_array()
            e0 = self._t0()
            check(self.lib.tgr_bwd_reduce(tabs, len(self.tables), H, structs, len(g.pbs), g.keys, g.srcs, n, 0,
                                          g.seg_of.data_ptr(), G.data_ptr(), None, rws.data_ptr(), rws.numel(), _stream()),
                  "tgr_bwd_reduce")
            self._t1("bwd_reduce", e0)
            self.launches += 2
            fws = self._buf("fact_bwd_ws", self.lib.tgr_fact_backward_workspace_bytes(len(self.tables), H), dev)
            e0 = self._t0()
            check(self.lib.tgr_fact_unique_backward(tabs, len(self.tables), H, C.byref(self._dnn()), g.uniq.data_ptr(),
                                                    g.n_unique.data_ptr(), g.cap, G.data_ptr(), a["dW_item"].data_ptr(),
                                                    None if a["dW_user"] is None else a["dW_user"].data_ptr(),
                                                    fws.data_ptr(), fws.numel(), _stream()), "tgr_fact_unique_backward")
            self._t1("fact_unique_backward", e0)
            self.launches += 2
        g.g_rows = G
        g.dz.clear()
        g.done = True
        if self.current is g:
            self.current = None

    def dense_from_rows(self, g: FactGroup) -> List[Optional[torch.Tensor]]:
        """Parity mode: dense [rows, H] gradients of the tables the group's calls index."""
        lay = self.layout
        touched = sorted({s.table for pb in g.pbs for s in lay.calls[pb.include_user].slots if s.table >= 0})
        grads: List[Optional[torch.Tensor]] = [None] * len(self.tables)
        for t in touched:
            grads[t] = torch.zeros_like(self.tables[t].data)
        if g.n:
            tabs = self._table_array(grads=grads)
            check(self.lib.tgr_scatter_rows(tabs, len(self.tables), lay.H, g.uniq.data_ptr(), g.g_rows.data_ptr(),
                                            g.n_unique.data_ptr(), g.cap, _stream()), "tgr_scatter_rows")
            self.launches += 1
        return grads

    # ------------------------------------------------------------------ row update
    def fused_step(self, lr: float = 1e-3, betas=(0.9, 0.98), eps: float = 1e-8, weight_decay: float = 1e-2,
                   grad_scale: float = 1.0):
        groups, self.ready = self.ready, []
        if not groups:
            return 0
        self._require_cuda()
        self.ensure_state()
        self.step += 1
        H, dev = self.layout.H, self._device()
        adam = make_adam(lr, betas[0], betas[1], eps, weight_decay, self.step, grad_scale)
        tabs = self._table_array(state=True)
        if len(groups) == 1:
            g = groups[0]
            if g.n:
                e0 = self._t0()
                check(self.lib.tgr_adam_rows(tabs, len(self.tables), H, g.uniq.data_ptr(), g.g_rows.data_ptr(),
                                             g.n_unique.data_ptr(), g.cap, C.byref(adam), _stream()), "tgr_adam_rows")
                self._t1("adam_rows", e0)
                self.launches += 1
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
            cs.item_cat, cs.item_ld, cs.cat_dtype = g.g_rows.data_ptr(), H, _lib.DTYPE_F32
        rws = self._buf("reduce_ws", self.lib.tgr_reduce_workspace_bytes(R, H), dev)
        check(self.lib.tgr_bwd_reduce(tabs, len(self.tables), H, structs, len(groups), keys_o.data_ptr(), srcs_o.data_ptr(), R,
                                      1, None, None, C.byref(adam), rws.data_ptr(), rws.numel(), _stream()), "tgr_bwd_reduce")
        self.launches += 8
        return R


class FactoredFn(torch.autograd.Function):
    """out = feat2emb(packed call) through the factored kernels. Inputs: tables, emb_transform (W, b) pairs,
    itemdnn (W, b), userdnn (W, b) — so autograd routes gradients where the reference's graph would."""

    @staticmethod
    def forward(ctx, engine: FactoredEngine, pb: PackedBatch, needs_grad: bool, *params):
        g = engine._group_of(pb)
        out, mask = engine.fact_forward(g, pb)
        ctx.engine, ctx.pb, ctx.group, ctx.mask = engine, pb, g, mask
        ctx.n_params = len(params)
        if needs_grad:      # (grad mode is always off inside Function.forward: the caller tells)
            g.n_fwd += 1
        return out.view(pb.B, pb.L, -1)

    @staticmethod
    def backward(ctx, d_out):
        eng, pb, g = ctx.engine, ctx.pb, ctx.group
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        if not eng.fact_backward(g, pb, ctx.mask, d_out):
            return (None, None, None, *grads)
        eng.group_backward(g)
        a = g.acc
        n_t = len(eng.tables)
        names = list(eng.layout.item_emb_feat)
        if eng.mode == "fused":
            eng.ready.append(g)
        else:
            dense = eng.dense_from_rows(g)
            for i in range(n_t):
                if ctx.needs_input_grad[3 + i]:
                    grads[i] = dense[i]
        for j, k in enumerate(names):
            grads[n_t + 2 * j] = a["dWmm"][k]
            grads[n_t + 2 * j + 1] = a["dbmm"][k]
        o = n_t + 2 * len(names)
        grads[o], grads[o + 1] = a["dW_item"], a["db_item"]
        grads[o + 2], grads[o + 3] = a["dW_user"], a["db_user"]
        return (None, None, None, *grads)
