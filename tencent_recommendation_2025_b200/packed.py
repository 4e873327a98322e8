"""Packed batches: the kernel-facing replacement of the reference's per-feature ``feat2tensor`` tensors.

The reference walks the B x L feature dicts once PER FEATURE (22 / 14 times per call) and issues one
synchronous host->device copy per feature (model/BaseLine/model.py:186-224,272). Here one
pass over the dicts fills ONE token-major int32 id matrix (+ CSR arrays, + dense mm inputs), staged in a
single pinned buffer and uploaded with one async copy per dtype.
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import chain
from operator import itemgetter
from typing import List, Optional, Sequence

import numpy as np
import torch

from .layout import FeatureLayout
from .synth import PackedCall


_PACK_LIB = None


def _pack_lib():
    """libtgr_pack.so (csrc/tgr_pack.c): the dict walk in C on the CPython API. ctypes.PyDLL keeps the GIL and turns
    a Python exception set by the C side into a raise."""
    global _PACK_LIB
    if _PACK_LIB is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtgr_pack.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found: build it with `python -m tencent_recommendation_2025_b200.build`")
        lib = C.PyDLL(path)
        lib.tgr_pack_single.restype = C.c_int
        lib.tgr_pack_single.argtypes = [C.py_object, C.c_long, C.c_long, C.py_object, C.c_void_p, C.c_long, C.c_long, C.c_void_p]
        lib.tgr_pack_array.restype = C.c_longlong
        lib.tgr_pack_array.argtypes = [C.py_object, C.c_long, C.c_long, C.py_object, C.c_void_p, C.c_void_p, C.c_longlong]
        lib.tgr_pack_mm.restype = C.c_int
        lib.tgr_pack_mm.argtypes = [C.py_object, C.c_long, C.c_long, C.py_object, C.c_long, C.c_void_p]
        _PACK_LIB = lib
    return _PACK_LIB


def pack_from_dicts(layout: FeatureLayout, seq, feature_array, mask=None, include_user: bool = False) -> PackedCall:
    """list[B] of indexable[L] of dict  ->  PackedCall (host), ONE walk over the token dicts in C (tgr_pack.c; SURVEY.md
    §8(f) N1). Same result as ``pack_from_dicts_py`` (the Python restatement the tests compare it with), ~5x faster."""
    call = layout.calls[include_user]
    seq_np = seq.detach().cpu().numpy() if isinstance(seq, torch.Tensor) else np.asarray(seq)
    if seq_np.ndim != 2:
        raise ValueError("seq must be [B, L]")
    B, L = seq_np.shape
    if len(feature_array) != B:
        raise ValueError(f"feature_array has {len(feature_array)} sequences, seq has {B}")
    for row in feature_array:
        if len(row) != L:
            # the reference's numpy row assignment fails on ragged input (model.py:217-222)
            raise ValueError("setting an array element with a sequence: ragged feature sequences")
    lib = _pack_lib()
    fa = feature_array if isinstance(feature_array, (list, tuple)) else list(feature_array)
    T = B * L
    flat = seq_np.reshape(-1).astype(np.int64)
    ids = np.zeros((T, call.n_single), np.int32)
    names = layout.single_slot_names(include_user)
    m = None
    if include_user:
        if mask is None:
            raise ValueError("include_user=True needs the token-type mask")
        m = (mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)).reshape(-1)
        ids[:, names.index("item_id")] = np.where(m == 1, flat, 0)
        ids[:, names.index("user_id")] = np.where(m == 2, flat, 0)
    else:
        ids[:, names.index("item_id")] = flat
    feat_cols = [(c, k) for c, k in enumerate(names) if k not in ("item_id", "user_id")]
    if feat_cols and T:
        keys = tuple(k for _, k in feat_cols)
        cols = np.ascontiguousarray([c for c, _ in feat_cols], np.int32)
        lib.tgr_pack_single(fa, B, L, keys, cols.ctypes.data, len(keys), call.n_single, ids.ctypes.data)
    arr_names = layout.array_slot_names(include_user)
    arr_off = np.zeros((len(arr_names), T + 1), np.int32)
    arr_vals: List[np.ndarray] = []
    base = 0
    for j, k in enumerate(arr_names):
        cnt = np.zeros(T, np.int32)
        cap = 4 * T + 1024                                   # one walk unless the lists average more than 4 values
        vals = np.empty(cap, np.int32)
        total = int(lib.tgr_pack_array(fa, B, L, k, cnt.ctypes.data, vals.ctypes.data, cap)) if T else 0
        if total > cap:
            vals = np.empty(total, np.int32)
            lib.tgr_pack_array(fa, B, L, k, cnt.ctypes.data, vals.ctypes.data, total)
        vals = vals[:total].copy()
        arr_off[j, 0] = base
        arr_off[j, 1:] = base + np.cumsum(cnt, dtype=np.int64)
        arr_vals.append(vals)
        base = int(arr_off[j, -1])
    arr_val = np.concatenate(arr_vals) if arr_vals else np.zeros((0,), np.int32)
    mm_x = []
    for k, d in layout.item_emb_feat.items():
        x = np.zeros((T, d), np.float32)
        if T:
            lib.tgr_pack_mm(fa, B, L, k, d, x.ctypes.data)
        mm_x.append(x)
    return PackedCall(B, L, include_user, ids, arr_off, arr_val, mm_x,
                      seq=seq_np.astype(np.int32), mask=None if m is None else np.asarray(m).reshape(B, L).astype(np.int32))


def pack_from_dicts_py(layout: FeatureLayout, seq, feature_array, mask=None, include_user: bool = False) -> PackedCall:
    """The tensorizer in Python / numpy (the restatement ``pack_from_dicts`` is tested against).
    list[B] of indexable[L] of dict  ->  PackedCall (host).  Semantics of model.py:237-247 + feat2tensor:

    * item/user id columns: ``(mask == 1) * seq`` / ``(mask == 2) * seq`` when include_user, else ``seq``;
    * sparse features: ``item[k]`` for every token (KeyError if a dict lacks k, as the reference);
      ragged sequences raise ValueError (model.py:222);
    * array features: the token's list with padding ids (0) dropped — summing row 0 is a no-op because
      row 0 is the all-zero padding row (nn.Embedding(padding_idx=0), main.py:106-111);
    * mm features: ``item[k]`` if present else zeros (model.py:288-293).
    """
    call = layout.calls[include_user]
    seq_np = seq.detach().cpu().numpy() if isinstance(seq, torch.Tensor) else np.asarray(seq)
    if seq_np.ndim != 2:
        raise ValueError("seq must be [B, L]")
    B, L = seq_np.shape
    if len(feature_array) != B:
        raise ValueError(f"feature_array has {len(feature_array)} sequences, seq has {B}")
    for row in feature_array:
        if len(row) != L:
            # the reference's numpy row assignment fails on ragged input (model.py:217-222)
            raise ValueError("setting an array element with a sequence: ragged feature sequences")
    T = B * L
    flat = seq_np.reshape(-1).astype(np.int64)
    ids = np.zeros((T, call.n_single), np.int32)
    names = layout.single_slot_names(include_user)
    if include_user:
        if mask is None:
            raise ValueError("include_user=True needs the token-type mask")
        m = (mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)).reshape(-1)
        ids[:, names.index("item_id")] = np.where(m == 1, flat, 0)
        ids[:, names.index("user_id")] = np.where(m == 2, flat, 0)
    else:
        ids[:, names.index("item_id")] = flat
    feat_cols = [(c, k) for c, k in enumerate(names) if k not in ("item_id", "user_id")]
    tokens = list(chain.from_iterable(feature_array))
    if feat_cols:
        keys = [k for _, k in feat_cols]
        if len(keys) == 1:
            vals = np.fromiter((tok[keys[0]] for tok in tokens), dtype=np.int64, count=T).reshape(T, 1)
        else:
            vals = np.array(list(map(itemgetter(*keys), tokens)), dtype=np.int64).reshape(T, len(keys))
        ids[:, [c for c, _ in feat_cols]] = vals
    arr_names = layout.array_slot_names(include_user)
    arr_off = np.zeros((len(arr_names), T + 1), np.int32)
    arr_vals: List[np.ndarray] = []
    base = 0
    for j, k in enumerate(arr_names):
        lists = [tok[k] for tok in tokens]
        lens = np.fromiter((len(v) for v in lists), dtype=np.int64, count=T)
        flatv = np.fromiter(chain.from_iterable(lists), dtype=np.int64, count=int(lens.sum()))
        keep = flatv != 0
        owner = np.repeat(np.arange(T), lens)[keep]
        cnt = np.bincount(owner, minlength=T)
        arr_off[j, 0] = base
        arr_off[j, 1:] = base + np.cumsum(cnt)
        arr_vals.append(flatv[keep].astype(np.int32))
        base = int(arr_off[j, -1])
    arr_val = np.concatenate(arr_vals) if arr_vals else np.zeros((0,), np.int32)
    mm_x = []
    for k, d in layout.item_emb_feat.items():
        x = np.zeros((T, d), np.float32)
        for t, tok in enumerate(tokens):
            v = tok.get(k) if isinstance(tok, dict) else (tok[k] if k in tok else None)
            if v is not None:
                x[t] = v
        mm_x.append(x)
    return PackedCall(B, L, include_user, ids, arr_off, arr_val, mm_x,
                      seq=seq_np.astype(np.int32), mask=None if mask is None else np.asarray(m).reshape(B, L).astype(np.int32))


@dataclass
class PackedBatch:
    """Device-resident packed call. ``n_valid`` (non-padding ids) is known on the host at pack time and sizes
    the backward's sort without a device->host sync."""

    B: int
    L: int
    include_user: bool
    ids: torch.Tensor                 # int32 [T, n_single]
    arr_off: torch.Tensor             # int32 [n_array, T+1]
    arr_val: torch.Tensor             # int32 [nnz]
    arr_tok: torch.Tensor             # int32 [nnz] token of each array value
    arr_begin: List[int]
    arr_nnz: List[int]
    mm_x: List[torch.Tensor]          # [T, mm_dim] float32 / bfloat16
    n_valid: int
    h2d_bytes: int = 0
    n_cap: Optional[int] = None       # fixed-shape calls (resident.StepShape): upper bound of n_valid for ANY content of the
                                      # buffers; the kernels then take the count from device memory (CUDA-graph replay)

    @property
    def T(self) -> int:
        return self.B * self.L


def _arr_tok(pc: PackedCall) -> np.ndarray:
    toks = []
    for j in range(pc.arr_off.shape[0]):
        lens = np.diff(pc.arr_off[j].astype(np.int64))
        toks.append(np.repeat(np.arange(pc.T, dtype=np.int32), lens))
    return np.concatenate(toks).astype(np.int32) if toks else np.zeros((0,), np.int32)


def count_valid(layout: FeatureLayout, pc: PackedCall) -> int:
    """Non-padding, in-range ids of a call == entries tgr_bwd_build_keys will emit."""
    call = layout.calls[pc.include_user]
    n = 0
    for s in call.slots:
        rows = layout.tables[s.table].rows if s.table >= 0 else 0
        if s.kind == 0:
            col = pc.ids[:, s.src]
            n += int(np.count_nonzero((col > 0) & (col < rows)))
        elif s.kind == 1:
            lo, hi = int(pc.arr_off[s.src, 0]), int(pc.arr_off[s.src, -1])
            v = pc.arr_val[lo:hi]
            n += int(np.count_nonzero((v > 0) & (v < rows)))
    return n


@dataclass
class HostPacked:
    """A packed call staged ONCE in pinned host memory (what a pin_memory DataLoader hands over):
    ``upload`` is then nothing but 1 + n_mm async H2D copies."""

    B: int
    L: int
    include_user: bool
    ints: torch.Tensor                # pinned int32: [ids | arr_off | arr_val | arr_tok], each part 16 B aligned
    offs: List[int]
    sizes: List[int]
    n_single: int
    n_arr: int
    arr_begin: List[int]
    arr_nnz: List[int]
    mm_x: List[torch.Tensor]          # pinned [T, mm_dim]
    n_valid: int
    group: Optional["StepGroup"] = None   # the step's other calls (PackingCollate)

    @property
    def T(self) -> int:
        return self.B * self.L

    @property
    def nbytes(self) -> int:
        return self.ints.numel() * 4 + sum(x.numel() * x.element_size() for x in self.mm_x)

    def pin_memory(self, device=None) -> "HostPacked":
        """DataLoader(pin_memory=True) hook (torch/utils/data/_utils/pin_memory.py calls ``.pin_memory()`` on custom
        batch members): page-lock the staging buffers in the MAIN process, in place."""
        if not self.ints.is_pinned():
            self.ints = self.ints.pin_memory()
            self.mm_x = [x.pin_memory() for x in self.mm_x]
        return self

    def upload(self, device, non_blocking: bool = True) -> "PackedBatch":
        dev = self.ints.to(device, non_blocking=non_blocking)
        o, s, T = self.offs, self.sizes, self.T
        ids = dev[o[0]:o[0] + s[0]].view(T, self.n_single)
        arr_off = dev[o[1]:o[1] + s[1]].view(self.n_arr, T + 1)
        mm = [x.to(device, non_blocking=non_blocking) for x in self.mm_x]
        return PackedBatch(self.B, self.L, self.include_user, ids, arr_off, dev[o[2]:o[2] + s[2]], dev[o[3]:o[3] + s[3]],
                           self.arr_begin, self.arr_nnz, mm, self.n_valid, self.nbytes)


def _may_pin() -> bool:
    """Pinned allocations need a CUDA context: never inside a DataLoader worker (fork after CUDA init cannot
    re-initialise it, and tensors crossing the worker queue lose their pinned status anyway)."""
    if not torch.cuda.is_available():
        return False
    try:
        from torch.utils.data import get_worker_info
        return get_worker_info() is None
    except Exception:
        return True


def stage_pinned(layout: FeatureLayout, pc: PackedCall, mm_dtype: torch.dtype = torch.float32,
                 pin: Optional[bool] = None) -> HostPacked:
    """Host staging of one packed call: pinned in the main process, plain CPU memory inside a DataLoader worker
    (``HostPacked.pin_memory`` lets ``DataLoader(pin_memory=True)`` pin it after it crossed the worker queue)."""
    if pin is None:
        pin = _may_pin()
    if pc.n_valid is None:
        pc.n_valid = count_valid(layout, pc)
    arr_tok = _arr_tok(pc)
    n_arr = pc.arr_off.shape[0]
    parts = [pc.ids.reshape(-1), pc.arr_off.reshape(-1), pc.arr_val, arr_tok]
    sizes = [p.size for p in parts]
    offs, tot = [], 0
    for sz in sizes:
        offs.append(tot)
        tot += (sz + 3) // 4 * 4
    ints = torch.empty(max(tot, 4), dtype=torch.int32, pin_memory=pin)
    v = ints.numpy()
    for p, o, sz in zip(parts, offs, sizes):
        v[o:o + sz] = p
    mm = []
    for x in pc.mm_x:
        t = torch.empty(x.shape, dtype=mm_dtype, pin_memory=pin)
        t.copy_(torch.from_numpy(np.ascontiguousarray(x)))
        mm.append(t)
    begins = [int(pc.arr_off[j, 0]) for j in range(n_arr)]
    nnz = [int(pc.arr_off[j, -1] - pc.arr_off[j, 0]) for j in range(n_arr)]
    return HostPacked(pc.B, pc.L, pc.include_user, ints, offs, sizes, pc.ids.shape[1], n_arr, begins, nnz, mm, pc.n_valid)


class PackingCollate:
    """``collate_fn`` wrapper for the reference's DataLoader (model/BaseLine/dataset.py:268-293): runs
    the dataset's own collate, then tensorizes the three feature lists of the step IN THE WORKER (one C walk each) and
    hands them over as ``HostPacked`` calls (plain CPU tensors inside a worker; pinned by ``DataLoader(pin_memory=True)``
    through ``HostPacked.pin_memory`` or directly when the collate runs in the main process) in place of ``seq_feat / pos_feat /
    neg_feat``. The model code stays as it is — ``feat2emb(seq, seq_feature, mask, include_user)`` recognises a packed
    call in the ``feature_array`` position — and the three calls of a step carry a reference to each other, so the
    factored engine prepares them as ONE group on the first ``feat2emb`` of the step.

        loader = DataLoader(dataset, batch_size=B, collate_fn=PackingCollate(model.layout, dataset.collate_fn))
    """

    def __init__(self, layout: FeatureLayout, base_collate, mm_dtype: torch.dtype = torch.float32):
        self.layout, self.base, self.mm_dtype = layout, base_collate, mm_dtype

    def __call__(self, batch):
        out = list(self.base(batch))
        seq, pos, neg, token_type = out[0], out[1], out[2], out[3]
        calls = [pack_from_dicts(self.layout, seq, out[6], token_type, True),
                 pack_from_dicts(self.layout, pos, out[7], None, False),
                 pack_from_dicts(self.layout, neg, out[8], None, False)]
        hps = [stage_pinned(self.layout, pc, self.mm_dtype) for pc in calls]
        group = StepGroup(hps)
        for hp in hps:
            hp.group = group
        out[6], out[7], out[8] = hps
        return tuple(out)


class StepGroup:
    """The packed calls of one training step (seq / pos / neg), uploaded together on first use."""

    def __init__(self, host_calls):
        self.host = list(host_calls)
        self.device: Optional[List[PackedBatch]] = None

    def upload(self, device) -> List[PackedBatch]:
        if self.device is None:
            self.device = [hp.upload(device) for hp in self.host]
        return self.device

    def batch_of(self, hp, device) -> PackedBatch:
        pbs = self.upload(device)
        for h, pb in zip(self.host, pbs):
            if h is hp:
                return pb
        raise KeyError("packed call is not part of this step group")


class HostPrefetcher:
    """Double-buffered host -> device feed (what a pin_memory DataLoader + a copy stream give a training loop):
    ``submit(host_calls)`` enqueues the H2D copies of one step's pinned ``HostPacked`` calls on a side stream into
    PERSISTENT device staging slots (no allocator traffic in the loop), ``take()`` makes the compute stream wait for
    the oldest submitted step and returns its ``PackedBatch`` list (views of the slot); ``retire()`` — called once
    the step that consumed the oldest taken batch is fully enqueued — records the event after which its slot may be
    rewritten. Submitting step k+1 before running step k overlaps its copies with step k's kernels."""

    def __init__(self, device, slots: int = 4):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [{"ints": [], "mm": [], "done": None} for _ in range(slots)]
        self.next = 0
        self.queue = []
        self.taken = []

    @staticmethod
    def _fit(lst, i, n, dtype, device, reader=None):
        while len(lst) <= i:
            lst.append(None)
        if lst[i] is None or lst[i].numel() < n or lst[i].dtype != dtype:
            if lst[i] is not None and reader is not None:
                lst[i].record_stream(reader)    # a step enqueued on the compute stream may still read the old buffer
            lst[i] = torch.empty(int(n * 1.1) + 64, dtype=dtype, device=device)
        return lst[i]

    def submit(self, host_calls: Sequence["HostPacked"]):
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if any(slot is t for t in self.taken):
            raise RuntimeError("HostPrefetcher: every slot is still in use (retire() consumed batches, or add slots)")
        if slot["done"] is not None:
            slot["done"].synchronize()          # the step that read this slot finished slots - 1 steps ago
        pbs = []
        reader = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            j = 0
            for i, hp in enumerate(host_calls):
                n = hp.ints.numel()
                dev = self._fit(slot["ints"], i, n, torch.int32, self.device, reader)[:n]
                dev.copy_(hp.ints, non_blocking=True)
                o, s, T = hp.offs, hp.sizes, hp.T
                mm = []
                for x in hp.mm_x:
                    d = self._fit(slot["mm"], j, x.numel(), x.dtype, self.device, reader)[:x.numel()].view(x.shape)
                    d.copy_(x, non_blocking=True)
                    mm.append(d)
                    j += 1
                pbs.append(PackedBatch(hp.B, hp.L, hp.include_user, dev[o[0]:o[0] + s[0]].view(T, hp.n_single),
                                       dev[o[1]:o[1] + s[1]].view(hp.n_arr, T + 1), dev[o[2]:o[2] + s[2]],
                                       dev[o[3]:o[3] + s[3]], hp.arr_begin, hp.arr_nnz, mm, hp.n_valid, hp.nbytes))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((pbs, ev, slot))

    def take(self) -> List[PackedBatch]:
        pbs, ev, slot = self.queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        slot["done"] = None
        self.taken.append(slot)
        return pbs

    def retire(self):
        """Everything that reads the oldest taken batch has been enqueued on the current stream."""
        slot = self.taken.pop(0)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        slot["done"] = done


class _PinnedPool:
    """Reusable pinned staging buffers. cudaHostAlloc costs milliseconds, so buffers are recycled; a buffer is
    handed out again only after the event recorded behind its last H2D copy has completed."""

    def __init__(self):
        self.free = {}   # (dtype, bucket) -> list of (tensor, event)

    @staticmethod
    def _bucket(n: int) -> int:
        b = 1024
        while b < n:
            b *= 2
        return b

    def get(self, n: int, dtype: torch.dtype) -> torch.Tensor:
        key = (dtype, self._bucket(max(n, 1)))
        lst = self.free.setdefault(key, [])
        for i, (t, ev) in enumerate(lst):
            if ev is None or ev.query():
                lst.pop(i)
                return t
        return torch.empty(key[1], dtype=dtype, pin_memory=True)

    def put(self, t: torch.Tensor, dtype: torch.dtype):
        ev = torch.cuda.Event()
        ev.record()
        self.free.setdefault((dtype, t.numel()), []).append((t, ev))


_POOL = _PinnedPool()


def to_device(layout: FeatureLayout, pc: PackedCall, device, mm_dtype: torch.dtype = torch.float32,
              pin: bool = True, non_blocking: bool = True) -> PackedBatch:
    """One pinned staging buffer + one async H2D copy for all integer data (and one per mm feature)."""
    use_pool = pin and torch.cuda.is_available() and torch.device(device).type == "cuda"
    if pc.n_valid is None:
        pc.n_valid = count_valid(layout, pc)   # host-known entry count of the backward (cached on the call)
    T = pc.T
    arr_tok = _arr_tok(pc)
    n_arr = pc.arr_off.shape[0]
    parts = [pc.ids.reshape(-1), pc.arr_off.reshape(-1), pc.arr_val, arr_tok]
    sizes = [p.size for p in parts]
    # 16-byte align every part (ids needs it for the 128-bit id-block loads)
    offs, tot = [], 0
    for s in sizes:
        offs.append(tot)
        tot += (s + 3) // 4 * 4
    n_int = max(tot, 4)
    stage = _POOL.get(n_int, torch.int32) if use_pool else torch.empty(n_int, dtype=torch.int32)
    st_np = stage.numpy()
    for p, o, s in zip(parts, offs, sizes):
        st_np[o:o + s] = p
    dev = stage[:n_int].to(device, non_blocking=non_blocking)
    if use_pool:
        _POOL.put(stage, torch.int32)
    ids = dev[offs[0]:offs[0] + sizes[0]].view(T, pc.ids.shape[1])
    arr_off = dev[offs[1]:offs[1] + sizes[1]].view(n_arr, T + 1)
    arr_val = dev[offs[2]:offs[2] + sizes[2]]
    arr_tok_d = dev[offs[3]:offs[3] + sizes[3]]
    mm = []
    h2d = tot * 4
    for x in pc.mm_x:
        n = x.size
        if use_pool:
            st = _POOL.get(n, mm_dtype)
            view = st[:n].view(x.shape)
            view.copy_(torch.from_numpy(np.ascontiguousarray(x)))     # converts to mm_dtype on the host
            mm.append(view.to(device, non_blocking=non_blocking))
            _POOL.put(st, mm_dtype)
        else:
            xt = torch.from_numpy(np.ascontiguousarray(x)).to(mm_dtype)
            mm.append(xt.to(device, non_blocking=non_blocking))
        h2d += n * torch.empty((), dtype=mm_dtype).element_size()
    begins = [int(pc.arr_off[j, 0]) for j in range(n_arr)]
    nnz = [int(pc.arr_off[j, -1] - pc.arr_off[j, 0]) for j in range(n_arr)]
    pb = PackedBatch(pc.B, pc.L, pc.include_user, ids, arr_off, arr_val, arr_tok_d, begins, nnz, mm,
                     pc.n_valid, h2d)
    return pb
