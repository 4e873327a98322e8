"""Cached construction of the C-ABI structs (include/tgr_embed.h). Filling ~40 ctypes fields per call costs tens
of microseconds in Python; the static part (slot table, table rows / key bases) is built once and memcpy'd."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import Call, Table


class StructCache:
    def __init__(self, layout):
        self.layout = layout
        self._call_tmpl: Dict[bool, Call] = {}
        self._tab_cache: Dict[bool, Tuple[tuple, C.Array]] = {}
        for inc in (True, False):
            cl = layout.calls[inc]
            c = Call()
            c.n_slots = len(cl.slots)
            c.n_single = cl.n_single
            c.n_arrays = cl.n_array
            for i, s in enumerate(cl.slots):
                c.slots[i].kind, c.slots[i].side, c.slots[i].col = s.kind, s.side, s.col
                c.slots[i].table, c.slots[i].src = s.table, s.src
            self._call_tmpl[inc] = c

    def call(self, pb, item_cat: torch.Tensor, user_cat: Optional[torch.Tensor], cat_dtype: int, err_ptr: Optional[int],
             out: Optional[Call] = None) -> Call:
        c = out if out is not None else Call()
        C.memmove(C.addressof(c), C.addressof(self._call_tmpl[pb.include_user]), C.sizeof(Call))
        c.T = pb.T
        c.ids = pb.ids.data_ptr()
        n_arr = c.n_arrays
        if n_arr:
            off_base = pb.arr_off.data_ptr()
            off_stride = pb.arr_off.stride(0) * 4
            tok_base = pb.arr_tok.data_ptr()
            for a in range(n_arr):
                c.arr_off[a] = off_base + a * off_stride
                c.arr_tok[a] = tok_base + 4 * pb.arr_begin[a]
                c.arr_begin[a] = pb.arr_begin[a]
                c.arr_nnz[a] = pb.arr_nnz[a]
            c.arr_val = pb.arr_val.data_ptr() if pb.arr_val.numel() else None
        c.item_cat = item_cat.data_ptr()
        c.item_ld = item_cat.stride(0)
        if user_cat is not None:
            c.user_cat = user_cat.data_ptr()
            c.user_ld = user_cat.stride(0)
        c.cat_dtype = cat_dtype
        c.err_flag = err_ptr
        return c

    def tables(self, weights: List[torch.Tensor], exp_avg=None, exp_avg_sq=None, grads=None) -> C.Array:
        """Table array; cached while every pointer is unchanged (the common case: parameters do not move)."""
        state = exp_avg is not None
        ptrs = tuple(w.data_ptr() for w in weights)
        if state:
            ptrs += tuple(t.data_ptr() for t in exp_avg) + tuple(t.data_ptr() for t in exp_avg_sq)
        if grads is None:
            hit = self._tab_cache.get(state)
            if hit is not None and hit[0] == ptrs:
                return hit[1]
        n = len(weights)
        arr = (Table * n)()
        for i, t in enumerate(self.layout.tables):
            arr[i].weight = ptrs[i]
            arr[i].exp_avg = ptrs[n + i] if state else None
            arr[i].exp_avg_sq = ptrs[2 * n + i] if state else None
            arr[i].grad = grads[i].data_ptr() if grads is not None and grads[i] is not None else None
            arr[i].rows = t.rows
            arr[i].key_base = t.key_base
        if grads is None:
            self._tab_cache[state] = (ptrs, arr)
        return arr
