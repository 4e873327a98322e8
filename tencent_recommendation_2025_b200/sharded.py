"""Row-sharded embedding tables across W GPUs of one box; the reference is single-GPU.

Partition: global key = table.key_base + id over the flat mega-table; owner = key mod W,
local_row = key div W. Each rank keeps its slice of the flat table (+ AdamW m, v) and a data-parallel
share of the batch. One exchange each way per call, always on DEDUPLICATED keys:

  forward : local unique keys -> bucket by owner -> all-to-all counts, then local-row ids (4 B each)
            -> owner gathers rows -> all-to-all rows back -> ids remapped to the received buffer ->
            the same fused gather/pool/concat kernel writes the concat buffers.
  backward: per-unique-key gradients reduced locally first -> all-to-all (local row, grad row) to the
            owner -> owner stable-sorts by row (source-rank order inside a row) -> fixed-order
            segmented sum + fused AdamW row update on its slice.

The per-rank logic is written as generators that *yield* collective requests, so the same code runs
(a) under torch.distributed (NCCL on GPUs, gloo in the CPU tests) and (b) with W emulated ranks inside
one process (``run_emulated``) — how the 1-GPU test box checks the W>1 routing bit-exactly.
Local compute is behind ``ShardOps``; the product implementation is ``CudaShardOps`` (C-ABI kernels).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Generator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import Call, check, make_adam
from .engine import EmbeddingEngine, _dtype_code, _stream
from .layout import FeatureLayout, KIND_ARRAY, KIND_MM, KIND_SINGLE
from .packed import PackedBatch


# ----------------------------------------------------------------------------------------------------
# key <-> (owner, local row) and table <-> shard layout  (pure index arithmetic, shared by every backend)
# ----------------------------------------------------------------------------------------------------
def shard_rows(total_rows: int, W: int) -> int:
    return (total_rows + W - 1) // W


def shard_of_tables(tables: Sequence[torch.Tensor], rank: int, W: int) -> torch.Tensor:
    """Slice of the flat mega-table owned by ``rank``: rows with global key % W == rank, at key // W."""
    flat = torch.cat([t.detach() for t in tables], dim=0)
    n = shard_rows(flat.shape[0], W)
    out = torch.zeros((n, flat.shape[1]), dtype=flat.dtype, device=flat.device)
    part = flat[rank::W]
    out[: part.shape[0]] = part
    return out


def tables_from_shards(layout: FeatureLayout, shards: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Inverse of shard_of_tables over all ranks -> per-table tensors in layout order (checkpoint keys)."""
    W = len(shards)
    H = shards[0].shape[1]
    flat = torch.zeros((layout.total_rows, H), dtype=shards[0].dtype, device=shards[0].device)
    for r, s in enumerate(shards):
        n = len(range(r, layout.total_rows, W))
        flat[r::W] = s[:n]
    return [flat[t.key_base:t.key_base + t.rows].clone() for t in layout.tables]


# ----------------------------------------------------------------------------------------------------
# peer-memory transport of the small messages (counts, bucketed local-row ids, dense gradients)
# ----------------------------------------------------------------------------------------------------
class SymmIO:
    """Symmetric-memory mailboxes of one rank (csrc/tgr_symm.cu). ``gather[q]`` [W*W] int32 receives every rank's per-owner
    counts (rank s stores its row into slot s of every peer), ``rows[q]`` holds this rank's bucketed local-row ids where the
    owners pull their bucket from; q alternates per prepared step (the look-ahead prepares step k+1 while step k's buffers
    are still being read). ``*_peers[q][r]`` = device pointer of rank r's buffer in this process' address space."""

    def __init__(self, W: int, rank: int, gather, rows, gather_peers, rows_peers, barrier=None):
        self.W, self.rank = W, rank
        self.gather, self.rows = gather, rows
        self.gather_peers, self.rows_peers = gather_peers, rows_peers
        self.barrier = barrier          # callable or None (emulated ranks: stream order is the barrier)
        self.parity = 0


def emulate_io(ranks, rows_cap: int):
    """W emulated ranks in one process: plain tensors stand in for the symmetric buffers, every rank sees the others'."""
    W = len(ranks)
    dev = ranks[0].ops.local.device
    gather = [[torch.zeros(W * W, dtype=torch.int32, device=dev) for _ in range(2)] for _ in range(W)]
    rows = [[torch.zeros(rows_cap, dtype=torch.int32, device=dev) for _ in range(2)] for _ in range(W)]
    for r, rk in enumerate(ranks):
        rk.ops.io = SymmIO(W, r, gather[r], rows[r], [[gather[s][q].data_ptr() for s in range(W)] for q in range(2)],
                           [[rows[s][q].data_ptr() for s in range(W)] for q in range(2)])


# ----------------------------------------------------------------------------------------------------
# local compute (product = CUDA kernels through the C ABI)
# ----------------------------------------------------------------------------------------------------
class CudaShardOps:
    """Per-rank device work of the sharded path; every method is a handful of C-ABI kernel launches."""

    def __init__(self, layout: FeatureLayout, local_table: torch.Tensor, mm: Dict[str, torch.nn.Linear], W: int):
        self.lib = _lib.load()
        self.layout = layout
        self.W = W
        self.local = local_table
        self.exp_avg = torch.zeros_like(local_table)
        self.exp_avg_sq = torch.zeros_like(local_table)
        dev = local_table.device
        # a helper engine over a dummy table set only provides key building / sorting / reduction plumbing
        self._dummy = [torch.nn.Parameter(torch.empty((0, layout.H), device=dev), requires_grad=False) for _ in layout.tables]
        self.mm = mm
        self._ws: Dict[str, torch.Tensor] = {}
        self.launches = 0
        self.step = 0
        self.io: Optional[SymmIO] = None      # peer-memory mailboxes (set by ShardedBaselineEmbedding / emulate_io)
        self._eng = _KeyEngine(layout, dev)

    def _buf(self, name, nbytes, dev):
        b = self._ws.get(name)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes * 1.25), 256), dtype=torch.uint8, device=dev)
            self._ws[name] = b
        return b

    # -- forward -------------------------------------------------------------------------------------
    def unique_keys(self, pbs: List[PackedBatch]):
        """Sorted unique global keys of the calls' non-padding ids -> (uniq int32[cap], n_unique int32[1], cap)."""
        dev = self.local.device
        n = sum(pb.n_valid for pb in pbs)
        keys, srcs = self._eng.sorted_pairs(pbs, None)
        cap = max(n, 1)
        uniq = torch.empty(cap, dtype=torch.int32, device=dev)
        seg_off = torch.empty(cap + 1, dtype=torch.int32, device=dev)
        n_unique = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._buf("dedup", self.lib.tgr_dedup_workspace_bytes(n), dev)
        check(self.lib.tgr_dedup(keys, n, uniq.data_ptr(), seg_off.data_ptr(), None, n_unique.data_ptr(), ws.data_ptr(),
                                 ws.numel(), _stream()), "tgr_dedup")
        self.launches += 12
        return uniq, n_unique, cap

    def route(self, uniq, n_unique, cap, out_rows: Optional[torch.Tensor] = None):
        dev = uniq.device
        if out_rows is not None:
            if out_rows.numel() < cap:
                raise ValueError(f"the step has up to {cap} unique rows, the symmetric id buffer holds {out_rows.numel()}: "
                                 "raise max_step_entries")
            rows = out_rows[:cap]
        else:
            rows = torch.empty(cap, dtype=torch.int32, device=dev)
        perm = torch.empty(cap, dtype=torch.int32, device=dev)
        counts = torch.zeros(self.W, dtype=torch.int32, device=dev)
        ws = self._buf("route", self.lib.tgr_route_workspace_bytes(cap, self.W), dev)
        check(self.lib.tgr_route_bucket(uniq.data_ptr(), n_unique.data_ptr(), cap, self.W, rows.data_ptr(), perm.data_ptr(),
                                        counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "tgr_route_bucket")
        self.launches += 3
        return rows, perm, counts

    # -- peer-memory transport (csrc/tgr_symm.cu) ------------------------------------------------------
    def io_put_counts(self, counts: torch.Tensor, q: int):
        """This rank's per-owner counts into slot `rank` of every peer's gather buffer (all-gather by stores)."""
        io = self.io
        ptrs = (C.c_void_p * io.W)(*io.gather_peers[q])
        check(self.lib.tgr_peer_put(ptrs, io.W, io.rank, counts.data_ptr(), io.W, _stream()), "tgr_peer_put")
        self.launches += 1

    def io_pull_ids(self, M: List[List[int]], q: int) -> torch.Tensor:
        """This owner's bucket out of every source's bucketed id list (M[s][o] = rows source s sends to owner o)."""
        io = self.io
        W, me = io.W, io.rank
        cnt = [M[s][me] for s in range(W)]
        R = sum(cnt)
        out = torch.empty(max(R, 1), dtype=torch.int32, device=self.local.device)
        if R:
            srcs = (C.c_void_p * W)(*[io.rows_peers[q][s] + 4 * sum(M[s][:me]) for s in range(W)])
            cnts = (C.c_int64 * W)(*cnt)
            check(self.lib.tgr_peer_pull(srcs, cnts, W, out.data_ptr(), _stream()), "tgr_peer_pull")
            self.launches += 1
        return out[:R]

    def merge_owner(self, recv_rows: torch.Tensor, recv_counts: Sequence[int], with_code: bool):
        """Owner-side order of the received ids: a stable W-way merge of the per-source buckets (each is sorted) — the same
        result as the stable radix sort it replaces, in one launch and without the torch index arithmetic."""
        R = int(sum(recv_counts))
        if R == 0:
            return None
        if max(recv_counts) >= (1 << 24):
            raise ValueError("more than 2^24 received contributions from one source in one step")
        dev = self.local.device
        keys_out = torch.empty(R, dtype=torch.int32, device=dev)
        code_out = torch.empty(R, dtype=torch.int32, device=dev)
        cnts = (C.c_int64 * len(recv_counts))(*[int(c) for c in recv_counts])
        check(self.lib.tgr_merge_buckets(recv_rows.data_ptr(), cnts, len(recv_counts), 1 if with_code else 0, keys_out.data_ptr(),
                                         code_out.data_ptr(), _stream()), "tgr_merge_buckets")
        self.launches += 1
        return keys_out, code_out

    def gather(self, rows: torch.Tensor, n: int) -> torch.Tensor:
        dev = self.local.device
        out = torch.empty((n, self.layout.H), dtype=torch.float32, device=dev)
        if n:
            n_dev = torch.full((1,), n, dtype=torch.int32, device=dev)   # a fill kernel: no pageable H2D copy / sync
            check(self.lib.tgr_gather_rows(self.local.data_ptr(), self.layout.H, rows.data_ptr(), n_dev.data_ptr(), n,
                                           out.data_ptr(), _stream()), "tgr_gather_rows")
            self.launches += 1
        return out

    def forward_from_rows(self, pb: PackedBatch, uniq, n_unique, perm, rows_buf: torch.Tensor, out_dtype,
                          remapped=None):
        """rows_buf [U+1, H] (row 0 zero) holds this call's unique rows in bucketed order; remap ids, then run
        the fused gather/pool/concat kernel on it as if it were one table."""
        lay = self.layout
        dev = rows_buf.device
        cl = lay.calls[pb.include_user]
        n_rows = rows_buf.shape[0]
        if remapped is not None:
            return self._fwd_remapped(pb, remapped[0], remapped[1], rows_buf, out_dtype)
        names = [s for s in cl.slots if s.kind == KIND_SINGLE]
        kb = (C.c_uint32 * _lib.MAX_SLOTS)()
        rw = (C.c_int32 * _lib.MAX_SLOTS)()
        for s in names:
            kb[s.src] = lay.tables[s.table].key_base
            rw[s.src] = lay.tables[s.table].rows
        ids_r = torch.empty_like(pb.ids)
        check(self.lib.tgr_remap_ids(pb.ids.data_ptr(), pb.ids.numel(), cl.n_single, kb, rw, uniq.data_ptr(),
                                     n_unique.data_ptr(), perm.data_ptr(), ids_r.data_ptr(), _stream()), "tgr_remap_ids")
        arr_r = torch.empty_like(pb.arr_val)
        for s in cl.slots:
            if s.kind != KIND_ARRAY or pb.arr_nnz[s.src] == 0:
                continue
            kb1 = (C.c_uint32 * 1)(lay.tables[s.table].key_base)
            rw1 = (C.c_int32 * 1)(lay.tables[s.table].rows)
            off = 4 * pb.arr_begin[s.src]
            check(self.lib.tgr_remap_ids(pb.arr_val.data_ptr() + off, pb.arr_nnz[s.src], 1, kb1, rw1, uniq.data_ptr(),
                                         n_unique.data_ptr(), perm.data_ptr(), arr_r.data_ptr() + off, _stream()),
                  "tgr_remap_ids(array)")
            self.launches += 1
        return self._fwd_remapped(pb, ids_r, arr_r, rows_buf, out_dtype)

    def _fwd_remapped(self, pb, ids_r, arr_r, rows_buf, out_dtype):
        lay = self.layout
        dev = rows_buf.device
        cl = lay.calls[pb.include_user]
        n_rows = rows_buf.shape[0]
        pb_r = PackedBatch(pb.B, pb.L, pb.include_user, ids_r, pb.arr_off, arr_r, pb.arr_tok, pb.arr_begin, pb.arr_nnz,
                           pb.mm_x, pb.n_valid)
        T = pb.T
        item_cat = torch.empty((T, cl.item_dim), dtype=out_dtype, device=dev)
        user_cat = torch.empty((T, cl.user_dim), dtype=out_dtype, device=dev) if pb.include_user else None
        tabs = (_lib.Table * len(lay.tables))()
        base = 0
        for i in range(len(lay.tables)):
            tabs[i].weight = rows_buf.data_ptr()
            tabs[i].rows = n_rows
            tabs[i].key_base = base
            base += n_rows
        call = self._eng.structs.call(pb_r, item_cat, user_cat, _dtype_code(out_dtype), None)
        e0 = self._eng._t0()
        check(self.lib.tgr_fwd_gather_pool_concat(tabs, len(lay.tables), lay.H, C.byref(call), _stream()),
              "tgr_fwd_gather_pool_concat")
        self._eng._t1("fwd_gather_pool_concat", e0)
        self.launches += 2
        esz = item_cat.element_size()
        for s in cl.slots:
            if s.kind != KIND_MM:
                continue
            lin = self.mm[s.name]
            x = pb.mm_x[s.src]
            check(self.lib.tgr_mm_proj_fwd(x.data_ptr(), _dtype_code(x.dtype), T, s.mm_dim, lin.weight.data.data_ptr(),
                                           lin.bias.data.data_ptr(), lay.H, item_cat.data_ptr() + s.col * esz,
                                           item_cat.stride(0), _dtype_code(out_dtype), _stream()), "tgr_mm_proj_fwd")
            self.launches += 1
        return item_cat, user_cat

    # -- step-level prefetch: one sort/dedup shared by the forward exchange and the backward reduction ----
    def prepare(self, pbs: List[PackedBatch]):
        """Sorted (key, src) pairs + dedup of ALL calls of a step; buffers are owned by the returned dict."""
        dev = self.local.device
        n = sum(pb.n_valid for pb in pbs)
        keys, srcs = self._eng.sorted_pairs(pbs, None)
        pairs = self._eng._ws.pop("pairs")          # keep this step's pairs alive until its backward
        cap = max(n, 1)
        uniq = torch.empty(cap, dtype=torch.int32, device=dev)
        seg_off = torch.empty(cap + 1, dtype=torch.int32, device=dev)
        seg_of = torch.empty(cap, dtype=torch.int32, device=dev)
        n_unique = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._buf("dedup", self.lib.tgr_dedup_workspace_bytes(n), dev)
        check(self.lib.tgr_dedup(keys, n, uniq.data_ptr(), seg_off.data_ptr(), seg_of.data_ptr(), n_unique.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream()), "tgr_dedup")
        self.launches += 12
        return {"pairs": pairs, "keys": keys, "srcs": srcs, "n": n, "uniq": uniq, "seg_of": seg_of,
                "n_unique": n_unique, "cap": cap, "pbs": list(pbs)}

    def remap_all(self, pf, perm):
        """ids of every prefetched call -> row numbers of the received buffer, by ONE scatter over the sorted
        pairs (no search); the few array values go through the searching remap."""
        lay = self.layout
        pbs = pf["pbs"]
        structs = (Call * len(pbs))()
        outs = []
        ptrs = (C.c_void_p * len(pbs))()
        for i, pb in enumerate(pbs):
            structs[i] = self._eng.structs._call_tmpl[pb.include_user]
            o = torch.zeros_like(pb.ids)
            outs.append(o)
            ptrs[i] = o.data_ptr()
        check(self.lib.tgr_remap_scatter(pf["srcs"], pf["seg_of"].data_ptr(), pf["n"], perm.data_ptr(), structs, len(pbs),
                                         ptrs, _stream()), "tgr_remap_scatter")
        self.launches += 1 + len(pbs)
        arrs = []
        for pb in pbs:
            cl = lay.calls[pb.include_user]
            arr_r = torch.empty_like(pb.arr_val)
            for s in cl.slots:
                if s.kind != KIND_ARRAY or pb.arr_nnz[s.src] == 0:
                    continue
                kb1 = (C.c_uint32 * 1)(lay.tables[s.table].key_base)
                rw1 = (C.c_int32 * 1)(lay.tables[s.table].rows)
                off = 4 * pb.arr_begin[s.src]
                check(self.lib.tgr_remap_ids(pb.arr_val.data_ptr() + off, pb.arr_nnz[s.src], 1, kb1, rw1, pf["uniq"].data_ptr(),
                                             pf["n_unique"].data_ptr(), perm.data_ptr(), arr_r.data_ptr() + off, _stream()),
                      "tgr_remap_ids(array)")
                self.launches += 1
            arrs.append(arr_r)
        pf["remapped"] = {id(pb): (o, a) for pb, o, a in zip(pbs, outs, arrs)}

    def reduce_cached(self, pf, calls):
        """Per-unique-key gradient rows from the step's cached sorted pairs (no second sort)."""
        dev = self.local.device
        H = self.layout.H
        n = pf["n"]
        structs = (Call * len(calls))()
        for i, (pb, di, du) in enumerate(calls):
            self._eng._call_struct(pb, di, du, out=structs[i])
        grads = torch.empty((pf["cap"], H), dtype=torch.float32, device=dev)
        if n:
            rws = self._buf("reduce", self.lib.tgr_reduce_workspace_bytes(n, H), dev)
            tabs = self._eng._table_array()
            check(self.lib.tgr_bwd_reduce(tabs, len(self.layout.tables), H, structs, len(calls), pf["keys"], pf["srcs"], n, 0,
                                          pf["seg_of"].data_ptr(), grads.data_ptr(), None, rws.data_ptr(), rws.numel(),
                                          _stream()), "tgr_bwd_reduce")
            self.launches += 2
        return grads

    def prepare_owner(self, recv_rows: torch.Tensor, R: int, recv_counts: Optional[Sequence[int]] = None):
        """Owner side, done during the forward: stable order of the requested local rows (source-rank order kept) — by
        merging the per-source buckets when their sizes are known, else by a stable sort."""
        if R == 0:
            return None
        if recv_counts is not None and R <= (1 << 24):
            return self.merge_owner(recv_rows, recv_counts, with_code=False)
        if R > (1 << 24):
            raise ValueError("more than 2^24 received contributions in one step")
        dev = self.local.device
        pos = torch.arange(R, dtype=torch.int32, device=dev)
        keys_out = torch.empty(R, dtype=torch.int32, device=dev)
        pos_out = torch.empty(R, dtype=torch.int32, device=dev)
        ws = self._buf("sort", self.lib.tgr_sort_workspace_bytes(R), dev)
        bits = max(1, int(self.local.shape[0] - 1).bit_length())
        check(self.lib.tgr_sort_pairs(recv_rows.data_ptr(), pos.data_ptr(), keys_out.data_ptr(), pos_out.data_ptr(), R, bits,
                                      ws.data_ptr(), ws.numel(), _stream()), "tgr_sort_pairs")
        self.launches += 5
        return keys_out, pos_out

    def apply_cached(self, owner_state, recv_grads: torch.Tensor, R: int, hyper: dict):
        self.step += 1
        if R == 0:
            return
        keys_out, pos_out = owner_state
        self._apply_sorted(keys_out, pos_out, recv_grads, R, hyper)

    # -- backward ------------------------------------------------------------------------------------
    def reduce(self, calls):
        """Per-unique-key gradient rows of the queued calls -> (uniq, n_unique, grads [cap, H], cap)."""
        uniq, seg_off, n_unique, grads, n = self._eng.dedup_reduce(calls)
        self.launches += 14
        return uniq, n_unique, grads, max(n, 1)

    def permute(self, grads, perm, n_unique, cap):
        out = torch.empty_like(grads)
        check(self.lib.tgr_permute_rows(grads.data_ptr(), self.layout.H, perm.data_ptr(), n_unique.data_ptr(), cap, 0,
                                        out.data_ptr(), _stream()), "tgr_permute_rows")
        self.launches += 1
        return out

    def mm_backward(self, pb, d_item):
        return self._eng.mm_backward_with(self.mm, pb, d_item)

    def apply(self, recv_rows: torch.Tensor, recv_grads: torch.Tensor, R: int, hyper: dict):
        """Owner side: contributions arrive ordered by (source rank, key); a stable sort by local row keeps the
        source-rank order inside each row, then the fixed-tile reduction + AdamW updates the slice in place."""
        self.step += 1
        if R == 0:
            return
        dev = self.local.device
        H = self.layout.H
        if R > (1 << 24):
            raise ValueError("more than 2^24 received contributions in one step")
        pos = torch.arange(R, dtype=torch.int32, device=dev)
        keys_out = torch.empty(R, dtype=torch.int32, device=dev)
        pos_out = torch.empty(R, dtype=torch.int32, device=dev)
        ws = self._buf("sort", self.lib.tgr_sort_workspace_bytes(R), dev)
        bits = max(1, int(self.local.shape[0] - 1).bit_length())
        check(self.lib.tgr_sort_pairs(recv_rows.data_ptr(), pos.data_ptr(), keys_out.data_ptr(), pos_out.data_ptr(), R, bits,
                                      ws.data_ptr(), ws.numel(), _stream()), "tgr_sort_pairs")
        self._apply_sorted(keys_out, pos_out, recv_grads, R, hyper)
        self.launches += 5

    def _apply_sorted(self, keys_out, pos_out, recv_grads, R, hyper):
        dev = self.local.device
        H = self.layout.H
        tab = (_lib.Table * 1)()
        tab[0].weight = self.local.data_ptr()
        tab[0].exp_avg = self.exp_avg.data_ptr()
        tab[0].exp_avg_sq = self.exp_avg_sq.data_ptr()
        tab[0].rows = self.local.shape[0]
        tab[0].key_base = 0
        call = Call()
        call.T = R
        call.n_slots = 1
        call.n_single = 1
        call.slots[0].kind, call.slots[0].side, call.slots[0].col, call.slots[0].table, call.slots[0].src = 0, 0, 0, 0, 0
        call.item_cat = recv_grads.data_ptr()
        call.item_ld = H
        call.cat_dtype = _lib.DTYPE_F32
        adam = make_adam(hyper["lr"], hyper["betas"][0], hyper["betas"][1], hyper["eps"], hyper["weight_decay"], self.step,
                         hyper.get("grad_scale", 1.0))
        rws = self._buf("reduce", self.lib.tgr_reduce_workspace_bytes(R, H), dev)
        check(self.lib.tgr_bwd_reduce(tab, 1, H, C.byref(call), 1, keys_out.data_ptr(), pos_out.data_ptr(), R, 1, None, None,
                                      C.byref(adam), rws.data_ptr(), rws.numel(), _stream()), "tgr_bwd_reduce")
        self.launches += 2


class FactShardOps(CudaShardOps):
    """Sharded path with the FACTORED kernels (factored.py): the exchange protocol, routing, owner-side gather and
    AdamW are CudaShardOps'; the per-rank forward / backward run on the rows fetched from their owners — projected
    once per unique row through the (replicated) itemdnn / userdnn blocks, gather-summed per token — so no rank ever
    builds a concat buffer. Needs the step-level prefetch protocol (one group per step)."""

    fetched_rows_direct = True

    def __init__(self, layout: FeatureLayout, local_table: torch.Tensor, mm, dnn, W: int):
        super().__init__(layout, local_table, mm, W)
        from .factored import FactoredEngine
        # device pointers of EVERY rank's shard in this process' address space (NVLink peer memory / the other emulated
        # ranks' tensors), own shard included. When set, the projection kernel reads the rows in place from their owners
        # — no owner-side gather, no row all-to-all; the exchange protocol only synchronises (one barrier per step).
        self.peers: Optional[List[int]] = None
        # gradient window: this rank's bucketed row gradients live in a buffer every owner can read over NVLink, so the
        # owner-side reduction PULLS its contributions in place — no gradient all-to-all, one barrier instead.
        self.grad_win: Optional[torch.Tensor] = None     # [rows, H] (symmetric memory)
        self.grad_peers: Optional[List[int]] = None      # base pointer of every rank's window
        self._bar = torch.zeros(1, device=local_table.device)
        dev = local_table.device
        dummy = [torch.nn.Parameter(torch.zeros((1, layout.H), device=dev), requires_grad=False) for _ in layout.tables]
        self.feng = FactoredEngine(layout, dummy, mm, dnn, mode="fused", check_shapes=False)

    def prepare(self, pbs: List[PackedBatch]):
        g = self.feng.prepare(pbs)
        return {"group": g, "n": g.n, "uniq": g.uniq, "n_unique": g.n_unique, "cap": int(g.c.cap), "pbs": list(pbs)}

    def remap_all(self, pf, perm):
        return   # ids already address the sorted unique list; the permutation is applied when the rows are projected

    def fetch_rows_async(self, pf):
        """Copy this step's unique rows out of their owners' shards (NVLink peer memory) into the group's arena on a
        SIDE stream: remote reads are latency-bound (~2 us, no local L2), so they run next to whatever value-independent
        work the caller enqueues now (the next step's key processing) instead of stalling the projection kernel."""
        g = pf["group"]
        if len(self.peers) != self.W or self.W > _lib.MAX_PEERS:
            raise ValueError("peer shard pointers do not match the world size")
        if g.n == 0:
            g.fetch_done = None
            return
        dev = self.local.device
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)                      # behind the id all-to-all, i.e. behind every owner's previous update
        self._side.wait_event(ready)
        ptrs = (C.c_void_p * self.W)(*self.peers)
        check(self.lib.tgr_fetch_peer_rows(ptrs, self.W, self.layout.H, g.c.uniq, g.c.n_unique, g.c.cap, g.c.rows_local,
                                           self._side.cuda_stream), "tgr_fetch_peer_rows")
        g.arena.record_stream(self._side)
        g.fetch_done = torch.cuda.Event()
        g.fetch_done.record(self._side)
        self.launches += 1

    def forward_prefetched(self, pb: PackedBatch, st, out_dtype=None):
        g = st["pf"]["group"]
        if not g.c.projected:
            if self.peers is not None:
                done = getattr(g, "fetch_done", None)
                if done is not None:
                    torch.cuda.current_stream(self.local.device).wait_event(done)
                g.c.src.fetched_rows = g.c.rows_local       # sorted unique order: no permutation
            else:
                g.c.src.fetched_rows = st["rows_buf"].data_ptr()
                g.c.src.fetched_perm = st["perm"].data_ptr()
                g.keep = (st["rows_buf"], st["perm"])
        return self.feng.fact_forward(g, pb)

    def prepare_owner(self, recv_rows: torch.Tensor, R: int, counts_dev: Optional[torch.Tensor] = None,
                      recv_counts: Optional[Sequence[int]] = None):
        """Owner side, done ahead: stable order of the requested local rows. With the gradient window the payload of
        an entry is `source rank << 24 | index inside that source's bucket` (where the owner will pull the row from)."""
        if self.grad_peers is None or counts_dev is None:
            return super().prepare_owner(recv_rows, R, recv_counts)
        if R == 0:
            return None
        if recv_counts is not None:
            return self.merge_owner(recv_rows, recv_counts, with_code=True)
        dev = self.local.device
        cnt = counts_dev.to(torch.int64)
        src_rank = torch.repeat_interleave(torch.arange(self.W, device=dev), cnt, output_size=R)
        first = torch.cumsum(cnt, 0) - cnt
        code = ((src_rank << 24) | (torch.arange(R, device=dev) - first[src_rank])).to(torch.int32)
        keys_out = torch.empty(R, dtype=torch.int32, device=dev)
        code_out = torch.empty(R, dtype=torch.int32, device=dev)
        ws = self._buf("sort", self.lib.tgr_sort_workspace_bytes(R), dev)
        bits = max(1, int(self.local.shape[0] - 1).bit_length())
        check(self.lib.tgr_sort_pairs(recv_rows.data_ptr(), code.data_ptr(), keys_out.data_ptr(), code_out.data_ptr(), R, bits,
                                      ws.data_ptr(), ws.numel(), _stream()), "tgr_sort_pairs")
        self.launches += 5
        return keys_out, code_out

    def permute_to_window(self, grads, perm, n_unique, cap):
        check(self.lib.tgr_permute_rows(grads.data_ptr(), self.layout.H, perm.data_ptr(), n_unique.data_ptr(), cap, 0,
                                        self.grad_win.data_ptr(), _stream()), "tgr_permute_rows")
        self.launches += 1

    def apply_from_peers(self, owner_state, starts: List[int], R: int, hyper: dict):
        """Fixed-order segmented sum of the contributions pulled from every source's window + AdamW on the slice."""
        self.step += 1
        if R == 0 or owner_state is None:
            return
        keys_out, code_out = owner_state
        H, dev = self.layout.H, self.local.device
        bases = (C.c_void_p * self.W)(*[self.grad_peers[s] + 4 * H * starts[s] for s in range(self.W)])
        tab = (_lib.Table * 1)()
        tab[0].weight, tab[0].exp_avg, tab[0].exp_avg_sq = self.local.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        tab[0].rows, tab[0].key_base = self.local.shape[0], 0
        adam = make_adam(hyper["lr"], hyper["betas"][0], hyper["betas"][1], hyper["eps"], hyper["weight_decay"], self.step,
                         hyper.get("grad_scale", 1.0))
        rws = self._buf("reduce", self.lib.tgr_reduce_workspace_bytes(R, H), dev)
        check(self.lib.tgr_bwd_reduce_rows(tab, 1, H, bases, self.W, keys_out.data_ptr(), code_out.data_ptr(), R, C.byref(adam),
                                           rws.data_ptr(), rws.numel(), _stream()), "tgr_bwd_reduce_rows")
        self.launches += 3

    def reduce_cached(self, pf, calls):
        g = pf["group"]
        if not g.done:
            raise RuntimeError("fused_step before the backward of every prefetched call")
        return g.rows("G")

    def release(self, pf):
        pf["group"].release()

    @property
    def total_launches(self):
        return self.launches + self.feng.launches


class _KeyEngine(EmbeddingEngine):
    """EmbeddingEngine plumbing (key build / sort / dedup / reduce) without owning real tables: the global
    key space comes from the layout, table pointers are never dereferenced by those kernels."""

    def __init__(self, layout: FeatureLayout, device):
        self.lib = _lib.load()
        self.layout = layout
        self.tables = [torch.nn.Parameter(torch.zeros((1, layout.H), device=device), requires_grad=False)
                       for _ in layout.tables]
        self.mm = {}
        self.mode = "fused"
        self.step = 0
        self.exp_avg = [None] * len(self.tables)
        self.exp_avg_sq = [None] * len(self.tables)
        self.pending = []
        self._ws = {}
        self.launches = 0
        self.check_ids = False
        self._err = None
        self.timing = None
        from ._structs import StructCache
        self._structs = StructCache(layout)
        self.structs = self._structs
        self._validated = True
        self._dev = torch.device(device)

    def _device(self):
        return self._dev

    def sorted_pairs(self, pbs: List[PackedBatch], dcats):
        """(keys, srcs) device pointers of the stably sorted pairs of the calls' non-padding ids."""
        calls = []
        for i, pb in enumerate(pbs):
            cl = self.layout.calls[pb.include_user]
            if dcats is None:
                # forward: no gradient buffers yet; the key builder never touches them
                di = torch.empty((0, cl.item_dim), device=self._dev)
                du = torch.empty((0, max(cl.user_dim, 1)), device=self._dev) if pb.include_user else None
            else:
                di, du = dcats[i]
            calls.append((pb, di, du))
        structs, keys, srcs, n, calls = self._sorted_pairs(calls)
        return keys, srcs

    def mm_backward_with(self, mm, pb, d_item):
        self.mm = mm
        return self.mm_backward(pb, d_item)


# ----------------------------------------------------------------------------------------------------
# per-rank state + generator orchestration
# ----------------------------------------------------------------------------------------------------
class ShardedRank:
    """One rank of the row-sharded path. ``ops`` does the local compute (CudaShardOps in product)."""

    def __init__(self, layout: FeatureLayout, ops, rank: int, W: int):
        self.layout, self.ops, self.rank, self.W = layout, ops, rank, W
        self.pending: list = []

    # Each generator yields ("a2a_equal", tensor[W]) or ("a2a_v", tensor, send_splits, recv_splits) and is sent
    # back the received tensor. Host split sizes come from ONE device->host read per exchange.
    # ---- step-level protocol, in three phases so that everything that does not depend on table VALUES can be
    # issued one step ahead (software pipelining): only phase C sits on the step's critical path.
    #   A  prepare_gen        : keys -> sort -> dedup -> route -> all-to-all of the per-owner counts; the split
    #                           sizes travel to the host by an ASYNC pinned copy (no blocking sync)
    #   B  finish_prepare_gen : (host sizes now known) all-to-all of the local-row ids, owner-side stable sort,
    #                           id remap — still independent of table values
    #   C  prefetch_gen       : owner gathers rows -> all-to-all rows back -> rows_buf
    # The backward then sends gradient rows only (the owner already knows which rows each source asked for,
    # in which order): no ids, no counts, no host sync.
    def _to_host_async(self, t: torch.Tensor):
        if t.is_cuda:
            ring = self.__dict__.setdefault("_pin_ring", [])
            if len(ring) < 4 or ring[0].shape != t.shape or ring[0].dtype != t.dtype:
                ring.insert(0, torch.empty(t.shape, dtype=t.dtype, pin_memory=True))   # pinned buffers are recycled
                del ring[4:]
            else:
                ring.insert(0, ring.pop())
            h = ring[0]
            h.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return h, ev
        return t.clone(), None

    def prepare_gen(self, pbs: List[PackedBatch]) -> Generator:
        ops = self.ops
        pf = ops.prepare(list(pbs))
        io = getattr(ops, "io", None) if getattr(ops, "grad_peers", None) is not None else None
        if io is not None:
            # peer-memory transport: ids go into this rank's symmetric buffer (owners pull their bucket later), the counts
            # are stored into every peer's gather buffer; one device-side barrier, no NCCL
            q = io.parity
            io.parity ^= 1
            rows_b, perm, counts = ops.route(pf["uniq"], pf["n_unique"], pf["cap"], out_rows=io.rows[q])
            ops.io_put_counts(counts, q)
            yield ("symm_barrier",)
            M = io.gather[q].view(self.W, self.W)
            host = self._to_host_async(M)
            self.prep = {"pbs": list(pbs), "pf": pf, "rows_b": rows_b, "perm": perm, "host": host, "stage": 1, "M": M, "io_q": q}
            return None
        rows_b, perm, counts = ops.route(pf["uniq"], pf["n_unique"], pf["cap"])
        if getattr(ops, "grad_peers", None) is not None:
            # the owner will PULL gradient rows out of every source's window: it needs the whole W x W count matrix
            # (where its bucket starts inside each source's buffer), so the counts are all-gathered
            M = yield ("allgather", counts)                       # [W, W], M[s, o] = rows rank s sends to owner o
            host = self._to_host_async(M)
            self.prep = {"pbs": list(pbs), "pf": pf, "rows_b": rows_b, "perm": perm, "host": host, "stage": 1, "M": M}
            return None
        recv_counts_t = yield ("a2a_equal", counts)
        host = self._to_host_async(torch.stack([counts, recv_counts_t]))
        self.prep = {"pbs": list(pbs), "pf": pf, "rows_b": rows_b, "perm": perm, "host": host, "stage": 1}
        return None

    def finish_prepare_gen(self) -> Generator:
        ops = self.ops
        p = self.prep
        if p is None or p["stage"] != 1:
            raise RuntimeError("finish_prepare without a pending prepare")
        h, ev = p["host"]
        if ev is not None:
            ev.synchronize()          # normally long complete: the copy was issued a phase earlier
        if getattr(ops, "peers", None) is not None and getattr(self, "pf", None) is not None:
            # peers read rows / write gradient windows as soon as this step's id all-to-all completes: that is only
            # safe behind every owner's previous row update, i.e. after fused_step() of the step in flight
            raise RuntimeError("finish_prepare() before fused_step() of the current step: the id all-to-all is the "
                               "barrier behind the owners' row update and must be enqueued after it")
        starts, window = None, False
        if "M" in p:
            M = h.tolist()
            p["M_host"] = M
            send_counts, recv_counts = M[self.rank], [M[s][self.rank] for s in range(self.W)]
            starts = [sum(M[s][:self.rank]) for s in range(self.W)]      # my bucket's first row in source s' window
            # every rank sees the same matrix, so all of them take the same branch
            window = max(sum(row) for row in M) <= ops.grad_win.shape[0] and max(max(row) for row in M) < (1 << 24)
        else:
            send_counts, recv_counts = h[0].tolist(), h[1].tolist()
        U, R = sum(send_counts), sum(recv_counts)
        if "io_q" in p:
            # behind this barrier every owner's previous row update is complete (it is enqueued after fused_step on every
            # rank) and every source's bucketed ids are in place: the owners pull their buckets, nobody waits on the host
            yield ("symm_barrier",)
            recv_rows = ops.io_pull_ids(p["M_host"], p["io_q"])
            p["synced"] = True
        else:
            recv_rows = yield ("a2a_v", p["rows_b"][:U], send_counts, recv_counts)
        if window:
            owner_state = ops.prepare_owner(recv_rows, R, counts_dev=p["M"][:, self.rank], recv_counts=recv_counts)
        else:
            owner_state = ops.prepare_owner(recv_rows, R, recv_counts=recv_counts)
        if hasattr(ops, "remap_all"):
            ops.remap_all(p["pf"], p["perm"])
        p.update(send_counts=send_counts, recv_counts=recv_counts, U=U, R=R, recv_rows=recv_rows, owner=owner_state, stage=2,
                 starts=starts, window=window)
        return None

    def prefetch_gen(self, pbs: List[PackedBatch]) -> Generator:
        """Fetch the unique rows of ALL calls of the step with one exchange. Uses the state prepared ahead by
        prepare_gen/finish_prepare_gen when it is for these very batches, else runs those phases inline."""
        ops = self.ops
        p = getattr(self, "prep", None)
        if p is None or len(p["pbs"]) != len(pbs) or any(a is not b for a, b in zip(p["pbs"], pbs)):
            yield from self.prepare_gen(pbs)
            p = self.prep
        if p["stage"] == 1:
            yield from self.finish_prepare_gen()
        self.prep = None
        if getattr(ops, "peers", None) is not None:
            # rows are read in place from the owners' shards by the projection kernel: all that is left of the forward
            # exchange is ordering — every owner's previous row update must be complete before anybody reads. The
            # local-row id all-to-all of this step was enqueued AFTER that update on every rank and completes here
            # only once every peer's part has arrived, so it already is that barrier — unless some pair of ranks
            # exchanged nothing (then no message orders them), in which case an explicit one is issued.
            # The decision must be the same on EVERY rank (a barrier is a collective): it is taken from the all-gathered
            # W x W count matrix when the step has one, else the barrier is always issued.
            M = p.get("M_host")
            if p.get("synced"):
                pass                                      # the peer-memory id exchange ended with a device-side barrier
            elif M is None or any(c == 0 for row in M for c in row):
                yield ("barrier", ops._bar)
            ops.fetch_rows_async(p["pf"])
            back = None
        else:
            served = ops.gather(p["recv_rows"], p["R"])
            back = yield ("a2a_v", served, p["recv_counts"], p["send_counts"])
        if back is None or getattr(ops, "fetched_rows_direct", False):
            rows_buf = back          # the factored kernels index the received rows through the permutation
        else:
            rows_buf = torch.cat([torch.zeros((1, self.layout.H), dtype=back.dtype, device=back.device), back], dim=0)
        self.pf = {"pbs": list(pbs), "pf": p["pf"], "perm": p["perm"], "rows_buf": rows_buf, "send_counts": p["send_counts"],
                   "recv_counts": p["recv_counts"], "U": p["U"], "R": p["R"], "owner": p["owner"],
                   "starts": p.get("starts"), "window": p.get("window", False)}
        self.last_fwd = {"U": p["U"], "R": p["R"], "send_counts": p["send_counts"], "recv_counts": p["recv_counts"]}
        return p["U"]

    def forward_gen(self, pb: PackedBatch, out_dtype=torch.float32) -> Generator:
        ops, W = self.ops, self.W
        st = getattr(self, "pf", None)
        if st is not None and any(pb is q for q in st["pbs"]):
            if hasattr(ops, "forward_prefetched"):
                return ops.forward_prefetched(pb, st, out_dtype)
            rm = st["pf"].get("remapped")
            if rm is not None:
                return ops.forward_from_rows(pb, st["pf"]["uniq"], st["pf"]["n_unique"], st["perm"], st["rows_buf"],
                                             out_dtype, remapped=rm[id(pb)])
            return ops.forward_from_rows(pb, st["pf"]["uniq"], st["pf"]["n_unique"], st["perm"], st["rows_buf"], out_dtype)
        uniq, n_unique, cap = ops.unique_keys([pb])
        rows_b, perm, counts = ops.route(uniq, n_unique, cap)
        recv_counts_t = yield ("a2a_equal", counts)
        both = torch.stack([counts, recv_counts_t]).cpu()          # the exchange's only host sync
        send_counts, recv_counts = both[0].tolist(), both[1].tolist()
        U, R = sum(send_counts), sum(recv_counts)
        recv_rows = yield ("a2a_v", rows_b[:U], send_counts, recv_counts)
        served = ops.gather(recv_rows, R)
        back = yield ("a2a_v", served, recv_counts, send_counts)
        rows_buf = torch.cat([torch.zeros((1, self.layout.H), dtype=back.dtype, device=back.device), back], dim=0)
        self.last_fwd = {"U": U, "R": R, "send_counts": send_counts, "recv_counts": recv_counts}
        return ops.forward_from_rows(pb, uniq, n_unique, perm, rows_buf, out_dtype)

    def queue(self, pb, d_item, d_user):
        self.pending.append((pb, d_item, d_user))

    def step_gen(self, hyper: dict) -> Generator:
        ops = self.ops
        pend, self.pending = self.pending, []
        H = self.layout.H
        st = getattr(self, "pf", None)
        if st is not None:
            self.pf = None
            by_id = {id(pb): (pb, di, du) for pb, di, du in pend}
            if len(by_id) != len(st["pbs"]) or any(id(pb) not in by_id for pb in st["pbs"]):
                raise RuntimeError("fused_step after prefetch needs the gradient of every prefetched call")
            calls = [by_id[id(pb)] for pb in st["pbs"]]            # source codes carry the prefetch call order
            grads = ops.reduce_cached(st["pf"], calls)
            if st.get("window"):
                # bucketed gradient rows go into this rank's window; once every rank has written (barrier) each owner's
                # reduction pulls its contributions over NVLink. The window is next written a whole step later, after
                # the following id all-to-all — which every owner enqueues behind this update.
                ops.permute_to_window(grads, st["perm"], st["pf"]["n_unique"], st["pf"]["cap"])
                yield (("symm_barrier",) if getattr(ops, "io", None) is not None else ("barrier", ops._bar))
                ops.apply_from_peers(st["owner"], st["starts"], st["R"], hyper)
            else:
                gb = ops.permute(grads, st["perm"], st["pf"]["n_unique"], st["pf"]["cap"])
                recv_grads = yield ("a2a_v", gb[:st["U"]], st["send_counts"], st["recv_counts"])
                ops.apply_cached(st["owner"], recv_grads, st["R"], hyper)
            if hasattr(ops, "release"):
                ops.release(st["pf"])
            self.last_step = {"U": st["U"], "R": st["R"], "send_counts": st["send_counts"], "recv_counts": st["recv_counts"]}
            return st["U"]
        if pend:
            uniq, n_unique, grads, cap = ops.reduce(pend)
            rows_b, perm, counts = ops.route(uniq, n_unique, cap)
            gb = ops.permute(grads, perm, n_unique, cap)
        else:   # a rank without work still takes part in the collectives
            dev = ops.local.device
            counts = torch.zeros(self.W, dtype=torch.int32, device=dev)
            rows_b = torch.zeros(1, dtype=torch.int32, device=dev)
            gb = torch.zeros((1, H), device=dev)
        recv_counts_t = yield ("a2a_equal", counts)
        both = torch.stack([counts, recv_counts_t]).cpu()
        send_counts, recv_counts = both[0].tolist(), both[1].tolist()
        U, R = sum(send_counts), sum(recv_counts)
        recv_rows = yield ("a2a_v", rows_b[:U], send_counts, recv_counts)
        recv_grads = yield ("a2a_v", gb[:U], send_counts, recv_counts)
        ops.apply(recv_rows, recv_grads, R, hyper)
        self.last_step = {"U": U, "R": R, "send_counts": send_counts, "recv_counts": recv_counts}
        return U


def run_distributed(gen: Generator, group=None, symm_barrier=None):
    """Drive one rank's generator with torch.distributed collectives (NCCL on GPUs, gloo on CPU). ``symm_barrier``: the
    device-side barrier of the peer-memory transport (signal pads of the symmetric allocation)."""
    import torch.distributed as dist
    pg = group if group is not None else dist.group.WORLD
    try:
        req = next(gen)
        while True:
            # straight to the ProcessGroup (the torch.distributed wrappers cost 50-75 us of host time per call);
            # wait() only orders the current stream behind the collective, the host does not block
            if req[0] == "symm_barrier":
                symm_barrier()
                out = None
            elif req[0] == "barrier":
                pg.allreduce([req[1]]).wait()
                out = None
            elif req[0] == "allgather":
                t = req[1].contiguous()
                flat = torch.empty(pg.size() * t.numel(), dtype=t.dtype, device=t.device)
                if hasattr(pg, "_allgather_base"):
                    pg._allgather_base(flat, t.reshape(-1)).wait()
                else:
                    dist.all_gather_into_tensor(flat, t.reshape(-1), group=group)
                out = flat.view((pg.size(),) + tuple(t.shape))
            elif req[0] == "a2a_equal":
                out = torch.empty_like(req[1])
                pg.alltoall_base(out, req[1].contiguous(), [], []).wait()
            else:
                _, t, ss, rs = req
                t = t.contiguous()
                out = torch.empty((sum(rs),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                pg.alltoall_base(out, t, list(rs), list(ss)).wait()
            req = gen.send(out)
    except StopIteration as e:
        return e.value


def run_emulated(gens: List[Generator]):
    """Drive W ranks' generators in lockstep inside one process, doing the all-to-alls by slicing."""
    W = len(gens)
    results = [None] * W
    live = [True] * W
    reqs = [None] * W
    for r, g in enumerate(gens):
        try:
            reqs[r] = next(g)
        except StopIteration as e:          # purely local work (e.g. forward after a prefetch)
            results[r] = e.value
            live[r] = False
    assert all(live) or not any(live), "ranks finished at different points"
    while any(live):
        kind = reqs[0][0]
        assert all(r[0] == kind for r in reqs), "ranks diverged"
        outs = []
        if kind in ("barrier", "symm_barrier"):
            outs = [None] * W
        elif kind == "allgather":
            full = torch.stack([reqs[s][1] for s in range(W)])
            outs = [full.clone() for _ in range(W)]
        elif kind == "a2a_equal":
            for r in range(W):
                outs.append(torch.stack([reqs[s][1][r] for s in range(W)]))
        else:
            for r in range(W):
                parts = []
                for s in range(W):
                    _, t, ss, rs = reqs[s]
                    o = sum(ss[:r])
                    parts.append(t[o:o + ss[r]])
                    assert ss[r] == reqs[r][3][s], "split sizes disagree"
                outs.append(torch.cat(parts, dim=0))
        for r in range(W):
            try:
                reqs[r] = gens[r].send(outs[r])
            except StopIteration as e:
                results[r] = e.value
                live[r] = False
        assert all(live) or not any(live), "ranks finished at different points"
    return results


# ----------------------------------------------------------------------------------------------------
# nn.Module face of one rank
# ----------------------------------------------------------------------------------------------------
class ShardedGatherConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, pb, out_dtype, *params):
        item_cat, user_cat = mod._run(mod.rank_state.forward_gen(pb, out_dtype))
        ctx.mod, ctx.pb = mod, pb
        ctx.n_params = len(params)
        if user_cat is None:
            user_cat = item_cat.new_empty(0)
            ctx.mark_non_differentiable(user_cat)
        return item_cat, user_cat

    @staticmethod
    def backward(ctx, d_item, d_user):
        mod, pb = ctx.mod, ctx.pb
        d_item = d_item.contiguous()
        d_user = d_user.contiguous() if pb.include_user else None
        mod.rank_state.queue(pb, d_item, d_user)
        grads = [None] * ctx.n_params
        names = list(mod.layout.item_emb_feat)
        if names and any(ctx.needs_input_grad[3:]):
            mg = mod.ops.mm_backward(pb, d_item)
            for j, k in enumerate(names):
                grads[2 * j], grads[2 * j + 1] = mg[k]
        return (None, None, None, *grads)


class ShardedFactoredFn(torch.autograd.Function):
    """Factored flavour: forward returns feat2emb's output itself; the replicated DNN / emb_transform gradients come
    back from the backward of the step's last call (they are the caller's to all-reduce), row gradients stay queued
    for fused_step's exchange with the owners."""

    @staticmethod
    def forward(ctx, mod, pb, needs_grad, *params):
        st = getattr(mod.rank_state, "pf", None)
        if st is None or not any(pb is q for q in st["pbs"]):
            raise RuntimeError("the factored sharded path needs prefetch(all calls of the step) before feat2emb_packed")
        out = mod._run(mod.rank_state.forward_gen(pb))
        g = st["pf"]["group"]
        ctx.mod, ctx.pb, ctx.group = mod, pb, g
        ctx.n_params = len(params)
        if needs_grad:
            g.n_fwd += 1
        return out          # [T, H], not a view (see FactoredFn.forward); reshaped by the caller

    @staticmethod
    def backward(ctx, d_out):
        mod, pb, g = ctx.mod, ctx.pb, ctx.group
        grads = [None] * ctx.n_params
        mod.rank_state.queue(pb, None, None)
        if not mod.ops.feng.fact_backward(g, pb, d_out):
            return (None, None, None, *grads)
        a = g.acc
        names = list(mod.layout.item_emb_feat)
        for j, k in enumerate(names):
            grads[2 * j] = a[f"dWmm/{k}"]
            grads[2 * j + 1] = a.get(f"dbmm/{k}")
        o = 2 * len(names)
        grads[o], grads[o + 1] = a["dW_item"], a["db_item"]
        grads[o + 2], grads[o + 3] = a["dW_user"], a["db_user"]
        return (None, None, None, *grads)


class ShardedBaselineEmbedding(torch.nn.Module):
    """Row-sharded variant of ``BaselineEmbedding``: this rank's slice of the flat table lives in
    ``local_table``; emb_transform / itemdnn / userdnn are replicated (their gradients are the caller's
    to all-reduce, as in any data-parallel run). Row updates are always fused (``fused_step``)."""

    def __init__(self, user_num, item_num, feat_statistics, feat_types, args, rank: int, world_size: int, group=None,
                 path: str = "concat", p2p: bool = True, grad_window_rows: int = 1 << 20, max_step_entries: int = 1 << 22,
                 symm_io: bool = True):
        super().__init__()
        if path not in ("concat", "factored"):
            raise ValueError("path must be 'concat' or 'factored'")
        self.path = path
        H = args.hidden_units
        lay = FeatureLayout(user_num, item_num, feat_statistics, feat_types, H)
        self.layout = lay
        self.rank, self.W, self.group = rank, world_size, group
        self.dev = args.device
        n_local = shard_rows(lay.total_rows, world_size)
        local, self._symm = self._alloc_shard(n_local, H, args.device, world_size, group, path == "factored" and p2p)
        self.local_table = torch.nn.Parameter(local, requires_grad=False)
        self.emb_transform = torch.nn.ModuleDict({k: torch.nn.Linear(d, H) for k, d in lay.item_emb_feat.items()})
        self.userdnn = torch.nn.Linear(lay.user_dim, H)
        self.itemdnn = torch.nn.Linear(lay.item_dim, H)
        self.to(args.device)
        if path == "factored":
            self.ops = FactShardOps(lay, self.local_table.data, dict(self.emb_transform.items()),
                                    {"item": self.itemdnn, "user": self.userdnn}, world_size)
        else:
            self.ops = CudaShardOps(lay, self.local_table.data, dict(self.emb_transform.items()), world_size)
        if self._symm is not None:
            # every rank's shard mapped into this process (CUDA VMM handles exchanged at the rendezvous)
            self._peer_views = [self._symm.get_buffer(r, (n_local, H), torch.float32) for r in range(world_size)]
            self.ops.peers = [t.data_ptr() for t in self._peer_views]
            # gradient window (same size on every rank): steps whose unique rows exceed it use the NCCL all-to-all
            win, self._symm_win = self._alloc_shard(int(grad_window_rows), H, args.device, world_size, group, True)
            if self._symm_win is not None:
                self._win_views = [self._symm_win.get_buffer(r, tuple(win.shape), torch.float32) for r in range(world_size)]
                self.ops.grad_win = win
                self.ops.grad_peers = [t.data_ptr() for t in self._win_views]
                if symm_io:
                    self._setup_symm_io(args.device, world_size, group, int(max_step_entries))
        self.rank_state = ShardedRank(lay, self.ops, rank, world_size)
        self._run = lambda gen: run_distributed(gen, self.group, self._symm_barrier)

    def _setup_symm_io(self, device, W: int, group, rows_cap: int):
        """Mailboxes of the peer-memory transport: ONE symmetric allocation [2 x (W*W counts) | 2 x rows_cap ids] (int32)."""
        n_g = (W * W + 63) // 64 * 64
        n_r = (rows_cap + 63) // 64 * 64
        buf, hdl = self._alloc_symm((2 * n_g + 2 * n_r,), torch.int32, device, W, group)
        if hdl is None:
            return
        self._io_buf, self._io_hdl = buf, hdl
        views = [hdl.get_buffer(r, (2 * n_g + 2 * n_r,), torch.int32) for r in range(W)]
        self._io_views = views
        base = [v.data_ptr() for v in views]
        gather = [buf[q * n_g: q * n_g + W * W] for q in range(2)]
        rows = [buf[2 * n_g + q * n_r: 2 * n_g + (q + 1) * n_r] for q in range(2)]
        self.ops.io = SymmIO(W, self.rank, gather, rows,
                             [[b + 4 * q * n_g for b in base] for q in range(2)],
                             [[b + 4 * (2 * n_g + q * n_r) for b in base] for q in range(2)])

    def _symm_barrier(self):
        """Device-side barrier between the ranks (blocks the current stream, not the host): tgr_peer_barrier over a symmetric
        flag array; falls back to the signal pads of the mailbox allocation."""
        b = getattr(self, "_bar_state", None)
        if b is None:
            buf, hdl = self._alloc_symm((64,), torch.int32, self.dev, self.W, self.group)
            if hdl is None or self.W > 32:
                b = self._bar_state = False
            else:
                views = [hdl.get_buffer(r, (64,), torch.int32) for r in range(self.W)]
                hdl.barrier(channel=0)          # every rank has zeroed its flags before anybody signals
                b = self._bar_state = {"buf": buf, "hdl": hdl, "views": views, "epoch": 0,
                                       "ptrs": (C.c_void_p * self.W)(*[v.data_ptr() for v in views])}
        if b is False:
            self._io_hdl.barrier(channel=0)
            return
        b["epoch"] += 1
        check(self.ops.lib.tgr_peer_barrier(b["ptrs"], self.rank, self.W, b["epoch"], _stream()), "tgr_peer_barrier")

    @staticmethod
    def _alloc_symm(shape, dtype, device, world_size: int, group):
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(shape, dtype=dtype, device=torch.device(device))
            hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
            t.zero_()
            return t, hdl
        except Exception as e:
            import warnings
            warnings.warn(f"symmetric memory unavailable ({e}); falling back to NCCL collectives")
            return None, None

    # -- replicated dense parameters: one-shot pull all-reduce over peer memory --------------------------------------
    def symm_empty(self, n: int) -> torch.Tensor:
        """A zeroed fp32 buffer of n elements (padded to a multiple of 4) every rank can read over NVLink — allocate the
        flat gradient buffer of the replicated parameters here and ``allreduce_dense_`` needs no NCCL collective."""
        n4 = (int(n) + 3) // 4 * 4
        if getattr(self.ops, "io", None) is not None:
            t, hdl = self._alloc_symm((n4,), torch.float32, self.dev, self.W, self.group)
            if hdl is not None:
                self._dense = (t, hdl, [hdl.get_buffer(r, (n4,), torch.float32) for r in range(self.W)],
                               torch.empty(n4, dtype=torch.float32, device=self.dev))
                return t
        self._dense = None
        return torch.zeros(n4, dtype=torch.float32, device=self.dev)

    def allreduce_dense_(self, flat: torch.Tensor, average: bool = True, pre_barrier: bool = True) -> torch.Tensor:
        """flat <- (mean | sum) over ranks, in place: one kernel in which every rank adds the W buffers in rank order
        (csrc/tgr_symm.cu; identical result everywhere) between two device-side barriers — every rank's backward is complete
        before anybody reads, every rank has read before anybody overwrites its own buffer. ``pre_barrier=False`` when the call
        follows ``fused_step`` of the same step (the barrier that orders the gradient windows there already orders the
        finished backward). Without symmetric memory: NCCL all-reduce."""
        import torch.distributed as dist
        d = getattr(self, "_dense", None)
        if d is None or d[0].data_ptr() != flat.data_ptr():
            dist.all_reduce(flat, group=self.group)
            if average:
                flat.div_(self.W)
            return flat
        t, hdl, views, tmp = d
        if pre_barrier:
            self._symm_barrier()
        ptrs = (C.c_void_p * self.W)(*[v.data_ptr() for v in views])
        check(self.ops.lib.tgr_allreduce_peers(ptrs, self.W, t.numel(), 1.0 / self.W if average else 1.0, tmp.data_ptr(), _stream()),
              "tgr_allreduce_peers")
        self._symm_barrier()            # every rank has read every buffer before anybody overwrites its own
        flat.copy_(tmp)
        return flat

    def dense_adam_(self, params, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2):
        """AdamW on the replicated dense parameters (itemdnn / userdnn / emb_transform) from their (already all-reduced)
        ``.grad`` tensors in ONE launch (tgr_adam_dense) — instead of torch's multi-tensor AdamW (37 us of device time and
        0.14 ms of host time per step for ~0.1 M parameters). State is kept here; same arithmetic as the row update."""
        from ._lib import DenseList, MAX_DENSE, make_adam
        params = [p for p in params if p.grad is not None]
        if not params:
            return
        if len(params) > MAX_DENSE:
            raise ValueError(f"more than {MAX_DENSE} dense tensors")
        st = getattr(self, "_dense_adam", None)
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        if st is None or st["key"] != key:
            dl = DenseList()
            dl.n = len(params)
            keep = []
            old = st["state"] if st is not None else {}
            state = {}
            for i, p in enumerate(params):
                if p.dtype != torch.float32 or not p.data.is_contiguous() or not p.grad.is_contiguous():
                    raise TypeError("dense parameters and gradients must be contiguous float32")
                mv = old.get(id(p)) or (torch.zeros_like(p.data), torch.zeros_like(p.data))
                state[id(p)] = mv
                dl.w[i], dl.g[i], dl.m[i], dl.v[i], dl.numel[i] = p.data.data_ptr(), p.grad.data_ptr(), mv[0].data_ptr(), mv[1].data_ptr(), p.numel()
                keep.append(p)
            st = self._dense_adam = {"key": key, "dl": dl, "state": state, "step": st["step"] if st is not None else 0}
        st["step"] += 1
        adam = make_adam(lr, betas[0], betas[1], eps, weight_decay, st["step"], 1.0)
        check(self.ops.lib.tgr_adam_dense(C.byref(st["dl"]), C.addressof(adam), None, _stream()), "tgr_adam_dense")

    @staticmethod
    def _alloc_shard(n_local: int, H: int, device, world_size: int, group, want_p2p: bool):
        """The rank's slice of the flat table. With p2p (factored path, W > 1, NCCL group up) it is allocated as torch
        symmetric memory so the other ranks' kernels can read its rows over NVLink; (tensor, handle | None)."""
        import torch.distributed as dist
        dev = torch.device(device)
        if want_p2p and world_size > 1 and dev.type == "cuda" and dist.is_available() and dist.is_initialized():
            try:
                import torch.distributed._symmetric_memory as symm
                t = symm.empty((n_local, H), dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
                t.zero_()
                return t, hdl
            except Exception as e:     # no VMM / fabric support on this box: the NCCL exchange still works
                import warnings
                warnings.warn(f"symmetric memory unavailable ({e}); rows go through the NCCL all-to-all")
        return torch.zeros((n_local, H), device=device), None

    def load_full_tables(self, tables: Sequence[torch.Tensor]):
        """Take this rank's slice out of full per-table tensors (layout order) — e.g. a reference checkpoint."""
        with torch.no_grad():
            self.local_table.copy_(shard_of_tables([t.to(self.local_table.device) for t in tables], self.rank, self.W))

    def feat2emb_packed(self, pb: PackedBatch):
        from .module import _concat_dtype
        params = []
        for k in self.layout.item_emb_feat:
            params += [self.emb_transform[k].weight, self.emb_transform[k].bias]
        if self.path == "factored":
            params += [self.itemdnn.weight, self.itemdnn.bias, self.userdnn.weight, self.userdnn.bias]
            needs = torch.is_grad_enabled() and any(p.requires_grad for p in params)
            out = ShardedFactoredFn.apply(self, pb, needs, *params).view(pb.B, pb.L, -1)
            return out.to(torch.bfloat16) if _concat_dtype() == torch.bfloat16 else out
        item_cat, user_cat = ShardedGatherConcatFn.apply(self, pb, _concat_dtype(), *params)
        B, L = pb.B, pb.L
        out = torch.relu(self.itemdnn(item_cat.view(B, L, -1)))
        if pb.include_user:
            out = out + torch.relu(self.userdnn(user_cat.view(B, L, -1)))
        return out

    def prefetch(self, pbs: Sequence[PackedBatch]):
        """Optional step-level fast path: exchange the unique rows of ALL the step's calls at once."""
        return self._run(self.rank_state.prefetch_gen(list(pbs)))

    def prepare_next(self, pbs: Sequence[PackedBatch]):
        """Look-ahead: start the key processing of the NEXT step's batches (independent of table values) so it
        overlaps this step; call ``finish_prepare`` later in the step, ``prefetch`` at the start of the next."""
        return self._run(self.rank_state.prepare_gen(list(pbs)))

    def finish_prepare(self):
        return self._run(self.rank_state.finish_prepare_gen())

    def fused_step(self, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        hyper = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_scale=grad_scale)
        return self._run(self.rank_state.step_gen(hyper))
