"""Device-resident item features (second half; csrc/tgr_resident.cu).

In the reference every item-side sparse feature and every frozen mm vector is a function of the ITEM ID: the dataset looks
them up in ``item_feat_dict[str(id)]`` / the mm store for every token of every batch (model/BaseLine/dataset.py:159,260-263)
and ``feat2emb`` then walks the resulting dicts (model.py:186-224,283-296). With the ``[items + 1, n_item_sparse]`` feature
table and the ``[items + 1, mm_dim]`` mm tables resident in HBM, a call's host side shrinks to

    item ids [T]  +  the user tokens (one per sequence: token index, user id, user-sparse values)  +  the user arrays (CSR)

— ~1 MB per call instead of ~20 MB — and the packed call the kernels consume is rebuilt on the device, bit for bit what the
host tensorizer produces (tests/test_gpu_resident.py).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check
from .layout import FeatureLayout
from .packed import PackedBatch, _arr_tok
from .synth import PackedCall


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class SlimCall:
    """Host side of one call when the item features live on the device (pinned int32 buffer + the part sizes)."""

    B: int
    L: int
    include_user: bool
    ints: torch.Tensor            # [item_ids T | user_tok nU | user_vals nU * n_ucols | arr_off | arr_val | arr_tok], parts 16 B aligned
    offs: List[int]
    sizes: List[int]
    n_user_tok: int
    n_arr: int
    arr_begin: List[int]
    arr_nnz: List[int]
    n_valid: int
    n_cap: Optional[int] = None   # fixed-shape calls: bound of n_valid for any content (see CallShape)

    @property
    def T(self) -> int:
        return self.B * self.L

    @property
    def nbytes(self) -> int:
        return self.ints.numel() * 4


@dataclass(frozen=True)
class CallShape:
    """Everything about a slim call that may size a buffer or a launch: with these fixed, two calls differ in CONTENT only
    (user tokens padded with token index -1, array values padded with id 0 behind each array's last token), which is what a
    captured CUDA graph needs (graphed.GraphedStep)."""

    B: int
    L: int
    include_user: bool
    n_user_cap: int
    arr_caps: tuple

    @staticmethod
    def covering(pcs: Sequence[PackedCall], slack: float = 1.5, round_to: int = 1024) -> "CallShape":
        """Shape that fits every example call with ``slack`` head-room on the array lengths."""
        p0 = pcs[0]
        n_arr = p0.arr_off.shape[0]
        caps = []
        for a in range(n_arr):
            m = max(int(pc.arr_off[a, -1] - pc.arr_off[a, 0]) for pc in pcs)
            caps.append((int(m * slack) + round_to) // round_to * round_to)
        return CallShape(p0.B, p0.L, bool(p0.include_user), p0.B if p0.include_user else 0, tuple(caps))


@dataclass
class SlimStep:
    """The slim calls of one step in ONE pinned buffer (one H2D copy per step); ``calls[i].ints`` are views of ``ints``."""

    ints: torch.Tensor
    calls: List[SlimCall]
    bases: List[int]

    @property
    def n_valid(self) -> int:
        return sum(c.n_valid for c in self.calls)


class ResidentItemFeatures:
    """Item-side feature tables in HBM + the device-side expansion of slim calls into ``PackedBatch``es."""

    def __init__(self, layout: FeatureLayout, item_feat: np.ndarray, mm_tables: Sequence, device, mm_dtype=torch.float32):
        self.lib = _lib.load()
        self.layout = layout
        self.device = torch.device(device)
        n_is = len(layout.item_sparse)
        item_feat = np.ascontiguousarray(item_feat, np.int32)
        if item_feat.ndim != 2 or item_feat.shape[1] != n_is:
            raise ValueError(f"item_feat must be [items + 1, {n_is}]")
        if item_feat[0].any():
            raise ValueError("row 0 of the item feature table (the padding item) must be all zero")
        self.n_items = item_feat.shape[0]
        full = layout.calls[True]
        names = layout.single_slot_names(True)
        self.id_col = {True: names.index("item_id"), False: layout.single_slot_names(False).index("item_id")}
        self.user_col0 = names.index("user_id")
        self.n_ucols = 1 + len(layout.user_sparse)
        # host: entries the key builder will emit per item id (the id itself + every in-range non-padding feature value)
        rows = [layout.tables[s.table].rows for s in full.slots if s.kind == 0 and s.side == 0][1:1 + n_is]
        item_rows = layout.tables[[s for s in full.slots if s.kind == 0 and s.side == 0][0].table].rows
        nnz = np.zeros(self.n_items, np.int32)
        for j, r in enumerate(rows):
            nnz += (item_feat[:, j] > 0) & (item_feat[:, j] < r)
        nnz[1:min(self.n_items, item_rows)] += 1
        self.nnz_item = nnz.astype(np.uint8)
        self.user_rows = [layout.tables[s.table].rows for s in full.slots if s.kind == 0 and s.side == 1]
        self.arr_rows = {s.src: layout.tables[s.table].rows for s in full.slots if s.kind == 1}
        self.feat_dev = torch.from_numpy(item_feat).to(self.device)
        self.mm_dev = []
        for (k, d), tab in zip(layout.item_emb_feat.items(), mm_tables):
            t = tab if isinstance(tab, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(tab))
            if tuple(t.shape) != (self.n_items, d):
                raise ValueError(f"mm table {k} must be [{self.n_items}, {d}]")
            self.mm_dev.append(t.to(self.device, dtype=mm_dtype).contiguous())
        self.mm_dtype = mm_dtype

    # ------------------------------------------------------------------ construction from the synthetic world
    @classmethod
    def from_world(cls, world, device, mm_dtype=torch.float32, chunk: int = 1 << 20) -> "ResidentItemFeatures":
        """The synthetic generator's feature functions evaluated for every item id (bench / tests)."""
        lay, V = world.layout, world.cfg.item_num
        feat = np.zeros((V + 1, len(lay.item_sparse)), np.int32)
        mm = [torch.empty((V + 1, d), dtype=mm_dtype) for d in lay.item_emb_feat.values()]
        for a in range(0, V + 1, chunk):
            ids = np.arange(a, min(V + 1, a + chunk), dtype=np.int64)
            feat[a:a + ids.size] = world.item_sparse_values(ids)
            for j, d in enumerate(lay.item_emb_feat.values()):
                mm[j][a:a + ids.size] = torch.from_numpy(world.mm_vectors(ids, j, d)).to(mm_dtype)
        return cls(lay, feat, mm, device, mm_dtype)

    # ------------------------------------------------------------------ host side
    def _slim_parts(self, pc: PackedCall, shape: Optional[CallShape]):
        lay = self.layout
        item_ids = np.ascontiguousarray(pc.ids[:, self.id_col[pc.include_user]])
        if pc.include_user:
            ucols = pc.ids[:, self.user_col0:self.user_col0 + self.n_ucols]
            tok = np.nonzero(ucols.any(axis=1))[0].astype(np.int32)
            uvals = np.ascontiguousarray(ucols[tok]).astype(np.int32)
        else:
            tok, uvals = np.zeros(0, np.int32), np.zeros((0, self.n_ucols), np.int32)
        arr_tok = _arr_tok(pc)
        n_arr = pc.arr_off.shape[0]
        begins = [int(pc.arr_off[a, 0]) for a in range(n_arr)]
        nnzs = [int(pc.arr_off[a, -1] - pc.arr_off[a, 0]) for a in range(n_arr)]
        arr_off, arr_val = pc.arr_off, pc.arr_val
        n_cap = None
        if shape is not None:
            if (pc.B, pc.L, bool(pc.include_user)) != (shape.B, shape.L, shape.include_user) or n_arr != len(shape.arr_caps):
                raise ValueError("call does not match its CallShape")
            if tok.size > shape.n_user_cap:
                raise ValueError(f"{tok.size} user tokens exceed the shape's capacity {shape.n_user_cap}")
            pad = shape.n_user_cap - tok.size
            tok_p = np.concatenate([tok, np.full(pad, -1, np.int32)])             # -1: skipped by the scatter kernel
            uvals_p = np.concatenate([uvals, np.zeros((pad, self.n_ucols), np.int32)])
            val_p = np.zeros(sum(shape.arr_caps), np.int32)                         # id 0 = padding: emits no key
            tok_a = np.zeros(sum(shape.arr_caps), np.int32)
            off_p = np.empty_like(pc.arr_off)
            b = 0
            src = 0
            for a, cap in enumerate(shape.arr_caps):
                if nnzs[a] > cap:
                    raise ValueError(f"array {a}: {nnzs[a]} values exceed the shape's capacity {cap}")
                val_p[b:b + nnzs[a]] = pc.arr_val[begins[a]:begins[a] + nnzs[a]]
                tok_a[b:b + nnzs[a]] = arr_tok[src:src + nnzs[a]]
                off_p[a] = pc.arr_off[a] - begins[a] + b
                src += nnzs[a]
                b += cap
            cl = lay.calls[pc.include_user]
            n_cap = pc.T * cl.n_single + sum(shape.arr_caps)
            begins_s = [int(x) for x in np.concatenate([[0], np.cumsum(shape.arr_caps)[:-1]])] if n_arr else []
            parts = [item_ids, tok_p, uvals_p.reshape(-1), off_p.reshape(-1), val_p, tok_a]
            meta = (shape.n_user_cap, n_arr, begins_s, [int(c) for c in shape.arr_caps])
        else:
            parts = [item_ids, tok, uvals.reshape(-1), pc.arr_off.reshape(-1), pc.arr_val, arr_tok]
            meta = (int(tok.size), n_arr, begins, nnzs)
        sizes = [int(p.size) for p in parts]
        offs, tot = [], 0
        for sz in sizes:
            offs.append(tot)
            tot += (sz + 3) // 4 * 4
        # entries the key builder emits: host-known without touching the feature values of the items
        in_range = (item_ids >= 0) & (item_ids < self.n_items)
        n = int(self.nnz_item[np.where(in_range, item_ids, 0)].sum(dtype=np.int64))
        for j, r in enumerate(self.user_rows):
            if pc.include_user:
                n += int(np.count_nonzero((uvals[:, j] > 0) & (uvals[:, j] < r)))
        for a in range(n_arr):
            vv = arr_val[begins[a]:begins[a] + nnzs[a]]
            n += int(np.count_nonzero((vv > 0) & (vv < self.arr_rows[a])))
        return parts, sizes, offs, max(tot, 4), meta, n, n_cap

    def slim(self, pc: PackedCall, pin: Optional[bool] = None, shape: Optional[CallShape] = None,
             out: Optional[torch.Tensor] = None) -> SlimCall:
        """What the data pipeline hands over for one call: ids + user tokens + user arrays (here derived from a packed call).
        ``shape``: pad to a fixed CallShape. ``out``: write into this (pinned) int32 buffer instead of allocating one."""
        parts, sizes, offs, tot, meta, n, n_cap = self._slim_parts(pc, shape)
        if out is None:
            if pin is None:
                pin = torch.cuda.is_available()
            out = torch.empty(tot, dtype=torch.int32, pin_memory=pin)
        elif out.numel() != tot:
            raise ValueError(f"slim: the output buffer holds {out.numel()} ints, the call needs {tot}")
        v = out.numpy()
        for p, o, sz in zip(parts, offs, sizes):
            v[o:o + sz] = p
        return SlimCall(pc.B, pc.L, pc.include_user, out, offs, sizes, meta[0], meta[1], meta[2], meta[3], n, n_cap)

    def slim_step(self, pcs: Sequence[PackedCall], shapes: Optional[Sequence[CallShape]] = None,
                  pin: Optional[bool] = None) -> SlimStep:
        """All calls of a step in one pinned buffer."""
        shapes = list(shapes) if shapes is not None else [None] * len(pcs)
        tots = [self._slim_parts(pc, sh)[3] for pc, sh in zip(pcs, shapes)]
        if pin is None:
            pin = torch.cuda.is_available()
        ints = torch.empty(sum(tots), dtype=torch.int32, pin_memory=pin)
        calls, bases, b = [], [], 0
        for pc, sh, t in zip(pcs, shapes, tots):
            calls.append(self.slim(pc, shape=sh, out=ints[b:b + t]))
            bases.append(b)
            b += t
        return SlimStep(ints, calls, bases)

    def slim_items(self, item_ids, pin: Optional[bool] = None) -> SlimCall:
        """Slim call of the candidate sweep's shape (model.py:418-425): one 'sequence' of n item ids, no user side. Only
        valid for layouts without item-side array features (true of both shipped variants, dataset.py:191-212)."""
        if self.layout.calls[False].n_array:
            raise ValueError("slim_items: the layout has item-side array features")
        ids = np.ascontiguousarray(item_ids, np.int32).reshape(-1)
        n = ids.size
        if pin is None:
            pin = torch.cuda.is_available()
        ints = torch.empty(max((n + 3) // 4 * 4, 4), dtype=torch.int32, pin_memory=pin)
        ints.numpy()[:n] = ids
        in_range = (ids >= 0) & (ids < self.n_items)
        nv = int(self.nnz_item[np.where(in_range, ids, 0)].sum(dtype=np.int64))
        o = (n + 3) // 4 * 4
        return SlimCall(1, n, False, ints, [0, o, o, o, o, o], [n, 0, 0, 0, 0, 0], 0, 0, [], [], nv)

    # ------------------------------------------------------------------ device side
    def expand(self, sc: SlimCall, dev_ints: torch.Tensor, ids_out: Optional[torch.Tensor] = None,
               mm_out: Optional[List[torch.Tensor]] = None, mm_stream: Optional["torch.cuda.Stream"] = None) -> PackedBatch:
        """Slim call (its int buffer already on the device) -> the packed call, on the current stream. ``mm_stream``: the mm
        row gathers go to that stream instead (the caller forks it from / joins it into the current stream): they only feed
        the mm projection, so they can run next to the id expansion and the key processing."""
        lay = self.layout
        cl = lay.calls[sc.include_user]
        T, o, s = sc.T, sc.offs, sc.sizes
        n_single = cl.n_single
        ids = ids_out if ids_out is not None else torch.empty((T, n_single), dtype=torch.int32, device=self.device)
        item_ids = dev_ints[o[0]:o[0] + s[0]]
        check(self.lib.tgr_expand_item_features(item_ids.data_ptr(), T, n_single, self.id_col[sc.include_user],
                                                self.id_col[sc.include_user] + 1, self.feat_dev.data_ptr(),
                                                self.feat_dev.shape[1], self.n_items, ids.data_ptr(), _stream()),
              "tgr_expand_item_features")
        if sc.include_user and sc.n_user_tok:
            check(self.lib.tgr_scatter_user_tokens(dev_ints[o[1]:].data_ptr(), dev_ints[o[2]:].data_ptr(), sc.n_user_tok,
                                                   self.n_ucols, n_single, self.user_col0, ids.data_ptr(), _stream()),
                  "tgr_scatter_user_tokens")
        mm = []
        for j, tab in enumerate(self.mm_dev):
            out = mm_out[j] if mm_out is not None else torch.empty((T, tab.shape[1]), dtype=tab.dtype, device=self.device)
            check(self.lib.tgr_gather_mm_rows(item_ids.data_ptr(), T, tab.data_ptr(), _lib.DTYPE_BF16 if tab.dtype == torch.bfloat16
                                              else _lib.DTYPE_F32, tab.shape[1], self.n_items, out.data_ptr(),
                                              mm_stream.cuda_stream if mm_stream is not None else _stream()),
                  "tgr_gather_mm_rows")
            mm.append(out)
        return PackedBatch(sc.B, sc.L, sc.include_user, ids, dev_ints[o[3]:o[3] + s[3]].view(sc.n_arr, T + 1),
                           dev_ints[o[4]:o[4] + s[4]], dev_ints[o[5]:o[5] + s[5]], sc.arr_begin, sc.arr_nnz, mm, sc.n_valid,
                           sc.nbytes, sc.n_cap)

    def to_device(self, sc: SlimCall) -> PackedBatch:
        return self.expand(sc, sc.ints.to(self.device, non_blocking=True))


class ResidentFeeder:
    """Double-buffered host -> device feed of slim calls: the H2D copies of step k+1 run on a copy stream while step k
    computes; ``take()`` makes the compute stream wait for the copies and expands the calls there into persistent slots."""

    def __init__(self, store: ResidentItemFeatures, slots: int = 4):
        self.store = store
        self.device = store.device
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [{"ints": [], "ids": [], "mm": [], "done": None} for _ in range(slots)]
        self.next = 0
        self.queue = []
        self.taken = []

    @staticmethod
    def _fit(lst, i, shape, dtype, device):
        n = 1
        for d in shape:
            n *= d
        while len(lst) <= i:
            lst.append(None)
        if lst[i] is None or lst[i].numel() < n or lst[i].dtype != dtype:
            lst[i] = torch.empty(int(n * 1.1) + 64, dtype=dtype, device=device)
        return lst[i][:n].view(shape)

    def submit(self, calls: Sequence[SlimCall]):
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if any(slot is t for t in self.taken):
            raise RuntimeError("ResidentFeeder: every slot is still in use (retire() consumed batches, or add slots)")
        if slot["done"] is not None:
            slot["done"].synchronize()
        devs = []
        with torch.cuda.stream(self.stream):
            for i, sc in enumerate(calls):
                d = self._fit(slot["ints"], i, (sc.ints.numel(),), torch.int32, self.device)
                d.copy_(sc.ints, non_blocking=True)
                devs.append(d)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((list(calls), devs, ev, slot))

    def take(self) -> List[PackedBatch]:
        calls, devs, ev, slot = self.queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        slot["done"] = None
        self.taken.append(slot)
        st, lay = self.store, self.store.layout
        pbs, j = [], 0
        for i, (sc, d) in enumerate(zip(calls, devs)):
            cl = lay.calls[sc.include_user]
            ids = self._fit(slot["ids"], i, (sc.T, cl.n_single), torch.int32, self.device)
            mm = []
            for tab in st.mm_dev:
                mm.append(self._fit(slot["mm"], j, (sc.T, tab.shape[1]), tab.dtype, self.device))
                j += 1
            pbs.append(st.expand(sc, d, ids, mm))
        return pbs

    def retire(self):
        slot = self.taken.pop(0)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        slot["done"] = ev
