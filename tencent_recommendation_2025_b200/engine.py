"""Host-side driver of the CUDA embedding path: builds the C structs, owns workspaces, wires autograd.

Everything here is plumbing (device pointers, streams, torch-allocated buffers); the work is done by
libtgr_embed.so. There is no fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Adam, Call, Table, check, make_adam
from .layout import FeatureLayout, KIND_MM
from .packed import PackedBatch
from ._structs import StructCache

_DEBUG = os.environ.get("TGR_DEBUG", "0") not in ("", "0")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    # raw cudaStream_t of torch's current stream (torch.cuda.current_stream() builds a Stream object: ~20 us per call)
    if _RAW_STREAM is not None:
        return _RAW_STREAM(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _lib.DTYPE_F32
    if dt == torch.bfloat16:
        return _lib.DTYPE_BF16
    raise TypeError(f"unsupported dtype {dt} (float32 / bfloat16 only)")


class EmbeddingEngine:
    """Runs feat2emb's gather/pool/concat/projection and its backward on tables it does not own.

    ``tables``: the nn.Embedding weights in ``layout.tables`` order (borrowed — the engine reads
    ``data_ptr()`` at call time, so ``.to()``, ``load_state_dict`` and in-place init keep working).
    ``mm``: {feature id: nn.Linear} (emb_transform).
    mode 'parity': backward returns real dense gradients (reference optimizer runs unchanged).
    mode 'fused' : backward queues the call; ``fused_step`` reduces all queued calls and applies the
                   AdamW row update in one pass; table ``.grad`` stays None.
    """

    def __init__(self, layout: FeatureLayout, tables: Sequence[torch.nn.Parameter], mm: Dict[str, torch.nn.Linear],
                 mode: str = "parity", check_shapes: bool = True):
        if mode not in ("parity", "fused"):
            raise ValueError("mode must be 'parity' or 'fused'")
        self.lib = _lib.load()
        self.layout = layout
        self.tables = list(tables)
        assert len(self.tables) == len(layout.tables)
        for p, t in zip(self.tables, layout.tables):
            if check_shapes and tuple(p.shape) != (t.rows, layout.H):
                raise ValueError(f"table {t.name}: shape {tuple(p.shape)} != {(t.rows, layout.H)}")
        self.mm = mm
        self.mode = mode
        self.step = 0
        self.exp_avg: List[Optional[torch.Tensor]] = [None] * len(self.tables)
        self.exp_avg_sq: List[Optional[torch.Tensor]] = [None] * len(self.tables)
        self.pending: List[Tuple[PackedBatch, torch.Tensor, Optional[torch.Tensor]]] = []
        self._ws: Dict[str, torch.Tensor] = {}
        self.launches = 0          # host-side estimate kept for debugging; the real count is _lib.launch_count()
        self.check_ids = _DEBUG
        self._err: Optional[torch.Tensor] = None
        self.timing: Optional[Dict[str, list]] = None   # set to {} to time every C-ABI call with CUDA events
        self._structs = StructCache(layout)
        self._validated = False

    # ------------------------------------------------------------------ per-kernel timing (bench / profiling)
    def _t0(self):
        if self.timing is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _t1(self, name: str, e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.timing.setdefault(name, []).append((e0, e1))

    def timing_summary(self) -> Dict[str, Tuple[float, int]]:
        """{call name: (total ms, launches)} over everything recorded since ``timing`` was set."""
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in (self.timing or {}).items()}

    # ------------------------------------------------------------------ buffers
    def _buf(self, name: str, nbytes: int, device) -> torch.Tensor:
        b = self._ws.get(name)
        if b is None or b.numel() < nbytes or b.device != torch.device(device):
            b = torch.empty(max(int(nbytes * 1.25), 256), dtype=torch.uint8, device=device)
            self._ws[name] = b
        return b

    def _device(self):
        return self.tables[0].device

    def _require_cuda(self):
        if not self.tables[0].is_cuda:
            raise _lib.TgrError("the embedding tables are not on a CUDA device; this path has no CPU fallback")

    def ensure_state(self):
        for i, p in enumerate(self.tables):
            if self.exp_avg[i] is None or self.exp_avg[i].device != p.device:
                self.exp_avg[i] = torch.zeros_like(p.data)
                self.exp_avg_sq[i] = torch.zeros_like(p.data)

    # ------------------------------------------------------------------ C structs
    def _table_array(self, state: bool = False, grads: Optional[List[Optional[torch.Tensor]]] = None):
        ws = [p.data for p in self.tables]
        if not self._validated:
            for w, t in zip(ws, self.layout.tables):
                if w.dtype != torch.float32 or not w.is_contiguous():
                    raise TypeError(f"table {t.name} must be contiguous float32")
            self._validated = True
        if state:
            return self._structs.tables(ws, self.exp_avg, self.exp_avg_sq, grads)
        return self._structs.tables(ws, None, None, grads)

    def _call_struct(self, pb: PackedBatch, item_cat: torch.Tensor, user_cat: Optional[torch.Tensor], out=None) -> Call:
        cl = self.layout.calls[pb.include_user]
        if pb.ids.dtype != torch.int32 or not pb.ids.is_contiguous() or tuple(pb.ids.shape) != (pb.T, cl.n_single):
            raise TypeError("PackedBatch.ids must be contiguous int32 [T, n_single]")
        err = None
        if self.check_ids:
            if self._err is None or self._err.device != item_cat.device:
                self._err = torch.zeros(1, dtype=torch.int32, device=item_cat.device)
            err = self._err.data_ptr()
        return self._structs.call(pb, item_cat, user_cat, _dtype_code(item_cat.dtype), err, out)

    # ------------------------------------------------------------------ forward
    def forward(self, pb: PackedBatch, out_dtype: torch.dtype = torch.float32):
        """-> (item_cat [T, item_dim], user_cat [T, user_dim] | None), every slot written by CUDA kernels."""
        self._require_cuda()
        lay = self.layout
        cl = lay.calls[pb.include_user]
        dev = self._device()
        T = pb.T
        item_cat = torch.empty((T, cl.item_dim), dtype=out_dtype, device=dev)
        user_cat = torch.empty((T, cl.user_dim), dtype=out_dtype, device=dev) if pb.include_user else None
        tabs = self._table_array()
        call = self._call_struct(pb, item_cat, user_cat)
        e0 = self._t0()
        check(self.lib.tgr_fwd_gather_pool_concat(tabs, len(self.tables), lay.H, C.byref(call), _stream()),
              "tgr_fwd_gather_pool_concat")
        self._t1("fwd_gather_pool_concat", e0)
        self.launches += 1
        esz = item_cat.element_size()
        for s in cl.slots:
            if s.kind != KIND_MM:
                continue
            lin = self.mm[s.name]
            x = pb.mm_x[s.src]
            if not x.is_contiguous() or x.shape != (T, s.mm_dim):
                raise TypeError(f"mm input {s.name} must be contiguous [T, {s.mm_dim}]")
            w, b = lin.weight.data, lin.bias.data if lin.bias is not None else None
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise TypeError("emb_transform weight must be contiguous float32")
            e0 = self._t0()
            check(self.lib.tgr_mm_proj_fwd(x.data_ptr(), _dtype_code(x.dtype), T, s.mm_dim, w.data_ptr(), _ptr(b), lay.H,
                                           item_cat.data_ptr() + s.col * esz, item_cat.stride(0),
                                           _dtype_code(item_cat.dtype), _stream()), "tgr_mm_proj_fwd")
            self._t1("mm_proj_fwd", e0)
            self.launches += 1
        if self.check_ids:
            bad = int(self._err.item())
            if bad:
                self._err.zero_()
                raise IndexError(f"index out of range in slot {cl.slots[bad - 1].name!r} (id >= table rows)")
        return item_cat, user_cat

    # ------------------------------------------------------------------ backward pieces
    def mm_backward(self, pb: PackedBatch, d_item: torch.Tensor):
        """-> {fid: (dW [H, mm_dim], db [H])} from the concat gradient, read in place."""
        lay = self.layout
        out = {}
        esz = d_item.element_size()
        for s in lay.calls[pb.include_user].slots:
            if s.kind != KIND_MM:
                continue
            x = pb.mm_x[s.src]
            dW = torch.empty((lay.H, s.mm_dim), dtype=torch.float32, device=d_item.device)
            db = torch.empty((lay.H,), dtype=torch.float32, device=d_item.device)
            nbytes = self.lib.tgr_mm_proj_bwd_workspace_bytes(pb.T, s.mm_dim, lay.H)
            ws = self._buf("mm_bwd", nbytes, d_item.device)
            e0 = self._t0()
            check(self.lib.tgr_mm_proj_bwd(x.data_ptr(), _dtype_code(x.dtype), pb.T, s.mm_dim,
                                           d_item.data_ptr() + s.col * esz, d_item.stride(0), _dtype_code(d_item.dtype),
                                           lay.H, dW.data_ptr(), db.data_ptr(), 0, ws.data_ptr(), ws.numel(), _stream()),
                  "tgr_mm_proj_bwd")
            self._t1("mm_proj_bwd", e0)
            self.launches += 2
            out[s.name] = (dW, db)
        return out

    def _sorted_pairs(self, calls: List[Tuple[PackedBatch, torch.Tensor, Optional[torch.Tensor]]]):
        """build_keys + sort for a list of (batch, d_item_cat, d_user_cat). -> (call structs, keys, srcs, n)."""
        dev = calls[0][1].device
        n_calls = len(calls)
        if n_calls > _lib.MAX_CALLS:
            raise ValueError(f"at most {_lib.MAX_CALLS} calls per reduction")
        structs = (Call * n_calls)()
        for i, (pb, di, du) in enumerate(calls):
            if not di.is_contiguous():
                di = di.contiguous()
            if du is not None and not du.is_contiguous():
                du = du.contiguous()
            calls[i] = (pb, di, du)
            self._call_struct(pb, di, du, out=structs[i])
        n = sum(pb.n_valid for pb, _, _ in calls)
        n_max = int(self.lib.tgr_bwd_max_entries(structs, n_calls))
        if n > n_max:
            raise ValueError("PackedBatch.n_valid exceeds the entry bound")
        pairs = self._buf("pairs", 16 * max(n, 1) + 64, dev)
        cnt = self._buf("n_valid", 16, dev)
        q = max(n, 1) * 4
        keys_a = pairs.data_ptr()
        srcs_a = keys_a + q
        keys_b = srcs_a + q
        srcs_b = keys_b + q
        ws_bytes = max(self.lib.tgr_build_keys_workspace_bytes(n_max), self.lib.tgr_sort_workspace_bytes(n))
        ws = self._buf("keys_ws", ws_bytes, dev)
        tabs = self._table_array()
        e0 = self._t0()
        check(self.lib.tgr_bwd_build_keys(tabs, len(self.tables), structs, n_calls, keys_a, srcs_a, cnt.data_ptr(),
                                          ws.data_ptr(), ws.numel(), _stream()), "tgr_bwd_build_keys")
        self._t1("bwd_build_keys", e0)
        self.launches += 3
        if _DEBUG:
            got = int(cnt.view(torch.int32)[0].item())
            if got != n:
                raise _lib.TgrError(f"PackedBatch.n_valid mismatch: host {n}, device {got}")
        e0 = self._t0()
        check(self.lib.tgr_sort_pairs(keys_a, srcs_a, keys_b, srcs_b, n, self.layout.key_bits, ws.data_ptr(), ws.numel(),
                                      _stream()), "tgr_sort_pairs")
        self._t1("sort_pairs", e0)
        self.launches += 5
        return structs, keys_b, srcs_b, n, calls

    def dedup_reduce(self, calls: List[Tuple[PackedBatch, torch.Tensor, Optional[torch.Tensor]]]):
        """Sparse gradient of the calls: (uniq keys [U] int64-able uint32 tensor, rows [U, H] fp32, counts)."""
        dev = calls[0][1].device
        structs, keys, srcs, n, calls = self._sorted_pairs(list(calls))
        H = self.layout.H
        uniq = torch.empty(max(n, 1), dtype=torch.int32, device=dev)       # uint32 payload
        seg_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
        seg_of = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        n_unique = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._buf("dedup_ws", self.lib.tgr_dedup_workspace_bytes(n), dev)
        check(self.lib.tgr_dedup(keys, n, uniq.data_ptr(), seg_off.data_ptr(), seg_of.data_ptr(), n_unique.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream()), "tgr_dedup")
        self.launches += 4
        grads = torch.empty((max(n, 1), H), dtype=torch.float32, device=dev)
        rws = self._buf("reduce_ws", self.lib.tgr_reduce_workspace_bytes(n, H), dev)
        tabs = self._table_array()
        check(self.lib.tgr_bwd_reduce(tabs, len(self.tables), H, structs, len(calls), keys, srcs, n, 0, seg_of.data_ptr(),
                                      grads.data_ptr(), None, rws.data_ptr(), rws.numel(), _stream()), "tgr_bwd_reduce")
        self.launches += 2
        return uniq, seg_off, n_unique, grads, n

    def dense_grads(self, pb: PackedBatch, d_item: torch.Tensor, d_user: Optional[torch.Tensor]):
        """Parity mode: dense [rows, H] gradients of the tables this call touches (None for the others)."""
        lay = self.layout
        touched = sorted({s.table for s in lay.calls[pb.include_user].slots if s.table >= 0})
        grads: List[Optional[torch.Tensor]] = [None] * len(self.tables)
        for t in touched:
            grads[t] = torch.zeros_like(self.tables[t].data)
        uniq, seg_off, n_unique, rows, n = self.dedup_reduce([(pb, d_item, d_user)])
        if n > 0:
            tabs = self._table_array(grads=grads)   # tables this call never indexes keep a NULL target
            check(self.lib.tgr_scatter_rows(tabs, len(self.tables), lay.H, uniq.data_ptr(), rows.data_ptr(),
                                            n_unique.data_ptr(), n, _stream()), "tgr_scatter_rows")
            self.launches += 1
        return grads

    # ------------------------------------------------------------------ fused step
    def queue(self, pb: PackedBatch, d_item: torch.Tensor, d_user: Optional[torch.Tensor]):
        self.pending.append((pb, d_item, d_user))

    def discard_pending(self) -> int:
        """Drop row gradients queued for a row update that will not happen (a GradScaler skipped the step)."""
        n = len(self.pending)
        self.pending = []
        return n

    def fused_step(self, lr: float = 1e-3, betas=(0.9, 0.98), eps: float = 1e-8, weight_decay: float = 1e-2,
                   grad_scale: float = 1.0):
        """Reduce every queued call's concat gradients per touched row and apply ONE AdamW row update
        (dense-Adam formula, torch/optim/adam.py:416-419,457,476,531-547, on touched rows only)."""
        if not self.pending:
            return 0
        self._require_cuda()
        self.ensure_state()
        self.step += 1
        total = 0
        pend, self.pending = self.pending, []
        for i in range(0, len(pend), _lib.MAX_CALLS):
            group = pend[i:i + _lib.MAX_CALLS]
            if i > 0:
                raise ValueError(f"more than {_lib.MAX_CALLS} feat2emb calls queued for one optimizer step")
            structs, keys, srcs, n, group = self._sorted_pairs(group)
            if n == 0:
                continue
            H = self.layout.H
            dev = group[0][1].device
            rws = self._buf("reduce_ws", self.lib.tgr_reduce_workspace_bytes(n, H), dev)
            tabs = self._table_array(state=True)
            adam = make_adam(lr, betas[0], betas[1], eps, weight_decay, self.step, grad_scale)
            e0 = self._t0()
            check(self.lib.tgr_bwd_reduce(tabs, len(self.tables), H, structs, len(group), keys, srcs, n, 1, None, None,
                                          C.byref(adam), rws.data_ptr(), rws.numel(), _stream()), "tgr_bwd_reduce")
            self._t1("bwd_reduce_adam", e0)
            self.launches += 2
            total += n
        return total


class GatherConcatFn(torch.autograd.Function):
    """item_cat, user_cat = gather/pool/project/concat(packed batch); table weights and emb_transform
    parameters are inputs so autograd routes gradients exactly where the reference's graph would."""

    @staticmethod
    def forward(ctx, engine: EmbeddingEngine, pb: PackedBatch, out_dtype, *params):
        item_cat, user_cat = engine.forward(pb, out_dtype)
        ctx.engine, ctx.pb = engine, pb
        ctx.n_params = len(params)
        if user_cat is None:
            user_cat = item_cat.new_empty(0)
            ctx.mark_non_differentiable(user_cat)
        return item_cat, user_cat

    @staticmethod
    def backward(ctx, d_item, d_user):
        eng, pb = ctx.engine, ctx.pb
        lay = eng.layout
        n_t = len(eng.tables)
        if not pb.include_user:
            d_user = None
        d_item = d_item.contiguous()
        if d_user is not None:
            d_user = d_user.contiguous()
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_params
        if eng.mode == "fused":
            eng.queue(pb, d_item, d_user)
        else:
            dense = eng.dense_grads(pb, d_item, d_user)
            for i in range(n_t):
                if ctx.needs_input_grad[3 + i]:
                    grads[i] = dense[i]
        mm_names = list(lay.item_emb_feat)
        if mm_names and any(ctx.needs_input_grad[3 + n_t:]):
            mg = eng.mm_backward(pb, d_item)
            for j, k in enumerate(mm_names):
                grads[n_t + 2 * j] = mg[k][0]
                grads[n_t + 2 * j + 1] = mg[k][1]
        return (None, None, None, *grads)
