"""Synthetic TencentGR-shaped batches; numpy only.

The reference ships no data (``.gitignore:1``), so every test and the benchmark draw from this
generator. One draw yields BOTH forms of the same batch:

  * the packed form the CUDA path consumes (``PackedCall``: token-major int32 id matrix, CSR
    arrays, dense mm inputs), and
  * the reference's list-of-dicts ``feature_array`` (model/BaseLine/dataset.py:268-293 collate
    output: list[B] of object-array[L] of dict) — small configs only, it is Python-object heavy.

Shape rules mirrored from the reference's dataset (model/BaseLine/dataset.py):
  * sequences are LEFT padded, padding tokens have id 0 / token_type 0 / all-default features (:123-167)
  * the user token is inserted at the front of the sequence (:119), token_type 2; items are 1
  * item-side features are a function of the item id (read from item_feat_dict, :159), mm vectors
    a function of the item id too (:260-263); 10 % of items have no mm vector -> zeros (:231-233)
  * defaults: sparse 0, array [0] (:214-225)
Everything is a pure function of (seed, id) through a splitmix64 hash, so no V-sized side tables
are needed even for 50M-row configs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .layout import FeatureLayout, KIND_ARRAY, KIND_MM, KIND_SINGLE, DEFAULT_FEAT_TYPES, default_feat_statistics

_U64 = np.uint64


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + _U64(0x9E3779B97F4A7C15)).astype(_U64)
    z = x
    z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
    return z ^ (z >> _U64(31))


def _hash_uniform(ids: np.ndarray, stream: int, seed: int) -> np.ndarray:
    """float64 uniform in [0,1) as a pure function of (seed, stream, id)."""
    with np.errstate(over="ignore"):
        h = _splitmix64(ids.astype(_U64) * _U64(0x2545F4914F6CDD1D) + _U64((seed * 1000003 + stream * 7919 + 12345) & 0xFFFFFFFFFFFFFFFF))
    return (h >> _U64(11)).astype(np.float64) * (1.0 / (1 << 53))


class BoundedZipf:
    """P(rank r) ∝ r^-alpha on [1, V]; exact inverse-CDF sampling; rank -> id via an affine bijection."""

    _cache: Dict[Tuple[int, float], np.ndarray] = {}

    def __init__(self, V: int, alpha: float, salt: int = 0):
        self.V = int(V)
        self.alpha = float(alpha)
        key = (self.V, self.alpha)
        if key not in BoundedZipf._cache:
            w = np.arange(1, self.V + 1, dtype=np.float64) ** (-self.alpha)
            c = np.cumsum(w)
            c /= c[-1]
            BoundedZipf._cache[key] = c
        self.cdf = BoundedZipf._cache[key]
        # affine permutation of ranks: id = 1 + (a*rank + b) mod V, gcd(a, V) = 1
        a = (0x9E3779B1 + 2 * salt) % self.V or 1
        while np.gcd(a, self.V) != 1:
            a += 1
        self.a, self.b = int(a), int((0x7F4A7C15 + salt) % self.V)

    def from_uniform(self, u: np.ndarray) -> np.ndarray:
        rank = np.searchsorted(self.cdf, u, side="left").astype(np.int64)  # 0-based rank
        np.minimum(rank, self.V - 1, out=rank)
        return (1 + (rank * self.a + self.b) % self.V).astype(np.int64)

    def sample(self, rng: np.random.Generator, size) -> np.ndarray:
        return self.from_uniform(rng.random(size))


@dataclass
class SynthConfig:
    B: int = 128
    L: int = 101
    H: int = 32
    item_num: int = 100_000
    user_num: int = 20_000
    alpha: float = 1.05
    feat_alpha: float = 1.05
    mm_ids: Tuple[str, ...] = ("81",)
    min_len: int = 20
    feat_statistics: Optional[Dict[str, int]] = None
    mm_missing: float = 0.10
    array_max_len: int = 10

    def feat_types(self) -> Dict[str, List[str]]:
        ft = {k: list(v) for k, v in DEFAULT_FEAT_TYPES.items()}
        ft["item_emb"] = list(self.mm_ids)
        return ft

    def statistics(self) -> Dict[str, int]:
        return dict(self.feat_statistics) if self.feat_statistics is not None else default_feat_statistics()

    def layout(self) -> FeatureLayout:
        return FeatureLayout(self.user_num, self.item_num, self.statistics(), self.feat_types(), self.H)


@dataclass
class PackedCall:
    """One feat2emb call in the packed, kernel-facing form (host numpy; see module.PackedBatch for device)."""

    B: int
    L: int
    include_user: bool
    ids: np.ndarray                      # int32 [T, n_single], token-major; column order = layout.single_slot_names
    arr_off: np.ndarray                  # int32 [n_array, T+1], absolute offsets into arr_val
    arr_val: np.ndarray                  # int32 [nnz]; padding id 0 already dropped
    mm_x: List[np.ndarray]               # float32 [T, mm_dim] per mm feature, zeros where the item has none
    seq: Optional[np.ndarray] = None     # int32 [B, L] raw ids (what the reference call receives)
    mask: Optional[np.ndarray] = None    # int32 [B, L] token types (include_user calls)
    n_valid: Optional[int] = None        # non-padding in-range ids (cached by packed.to_device)

    @property
    def T(self) -> int:
        return self.B * self.L

    def n_lookups(self) -> int:
        """Non-padding table-row lookups (the metric's 'rows')."""
        return int(np.count_nonzero(self.ids)) + int(np.count_nonzero(self.arr_val))


@dataclass
class SynthStep:
    """One training step's worth of calls: seq (include_user), pos, neg (model.py:324,376-377)."""

    cfg: SynthConfig
    layout: FeatureLayout
    calls: List[PackedCall]
    dicts: Optional[List[list]] = None          # per call: reference-form feature_array, or None
    upstream: Optional[List[np.ndarray]] = None  # per call: float32 [B, L, H] injected dOut (seed+1)

    def n_lookups(self) -> int:
        return sum(c.n_lookups() for c in self.calls)


class SynthWorld:
    """Deterministic feature functions of (seed, id) + batch sampler."""

    def __init__(self, cfg: SynthConfig, seed: int = 0):
        self.cfg = cfg
        self.seed = int(seed)
        self.layout = cfg.layout()
        self.item_zipf = BoundedZipf(cfg.item_num, cfg.alpha, salt=1)
        self.user_zipf = BoundedZipf(cfg.user_num, cfg.alpha, salt=2)
        st = cfg.statistics()
        self._feat_zipf = {k: BoundedZipf(st[k], cfg.feat_alpha, salt=100 + i) for i, k in enumerate(st)}

    # ---- pure functions of the id ------------------------------------------------
    def item_sparse_values(self, item_ids: np.ndarray) -> np.ndarray:
        """int32 [n, n_item_sparse]; row of zeros for id 0."""
        cols = []
        for j, k in enumerate(self.layout.item_sparse):
            v = self._feat_zipf[k].from_uniform(_hash_uniform(item_ids, 10 + j, self.seed))
            cols.append(np.where(item_ids != 0, v, 0))
        return np.stack(cols, axis=1).astype(np.int32) if cols else np.zeros((item_ids.size, 0), np.int32)

    def user_sparse_values(self, user_ids: np.ndarray) -> np.ndarray:
        cols = []
        for j, k in enumerate(self.layout.user_sparse):
            v = self._feat_zipf[k].from_uniform(_hash_uniform(user_ids, 50 + j, self.seed))
            cols.append(np.where(user_ids != 0, v, 0))
        return np.stack(cols, axis=1).astype(np.int32) if cols else np.zeros((user_ids.size, 0), np.int32)

    def user_array_values(self, user_ids: np.ndarray, j: int, k: str) -> Tuple[np.ndarray, np.ndarray]:
        """(lengths int32 [n], values int32 [n, Amax]) for user-array feature #j; length 0 for id 0."""
        A = self.cfg.array_max_len
        u = _hash_uniform(user_ids, 70 + j, self.seed)
        # 1 + Poisson(3) clipped to A, through the Poisson inverse CDF
        pm = np.exp(-3.0) * np.cumprod(np.concatenate([[1.0], 3.0 / np.arange(1, 40)]))
        lens = 1 + np.searchsorted(np.cumsum(pm), u)
        lens = np.minimum(lens, A).astype(np.int32)
        lens = np.where(user_ids != 0, lens, 0).astype(np.int32)
        vals = np.zeros((user_ids.size, A), np.int32)
        for a in range(A):
            ua = _hash_uniform(user_ids * 16 + a, 90 + j, self.seed)
            vals[:, a] = self._feat_zipf[k].from_uniform(ua)
        vals *= (np.arange(A)[None, :] < lens[:, None])
        return lens, vals

    def mm_vectors(self, item_ids: np.ndarray, j: int, dim: int) -> np.ndarray:
        """float32 [n, dim] ~ N(0,1) by Box-Muller on hashed uniforms; zeros for id 0 and for 'missing' items."""
        n = item_ids.size
        idx = item_ids.astype(np.int64)[:, None] * dim + np.arange(dim, dtype=np.int64)[None, :]
        u1 = _hash_uniform(idx.ravel(), 200 + 2 * j, self.seed)
        u2 = _hash_uniform(idx.ravel(), 201 + 2 * j, self.seed)
        z = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
        x = z.reshape(n, dim).astype(np.float32)
        present = (item_ids != 0) & (_hash_uniform(item_ids, 300 + j, self.seed) >= self.cfg.mm_missing)
        x[~present] = 0.0
        return x

    def mm_present(self, item_ids: np.ndarray, j: int) -> np.ndarray:
        return (item_ids != 0) & (_hash_uniform(item_ids, 300 + j, self.seed) >= self.cfg.mm_missing)

    # ---- batch sampler ----------------------------------------------------------
    def sample_sequences(self, rng: np.random.Generator):
        cfg = self.cfg
        B, L = cfg.B, cfg.L
        lo = min(cfg.min_len, L)
        n = rng.integers(lo, L + 1, size=B)
        p = np.arange(L)[None, :]
        start = (L - n)[:, None]
        real = p >= start
        is_user = p == start
        mask = np.where(is_user, 2, np.where(real, 1, 0)).astype(np.int32)
        items = self.item_zipf.sample(rng, (B, L))
        users = self.user_zipf.sample(rng, (B, 1))
        seq = np.where(is_user, users, np.where(real, items, 0)).astype(np.int32)
        # pos[t] = next token when it is an item (dataset.py:149-153); the last position's next is the held-out item
        nxt = np.concatenate([seq[:, 1:], self.item_zipf.sample(rng, (B, 1)).astype(np.int32)], axis=1)
        pos = np.where(real, nxt, 0).astype(np.int32)
        neg = np.where(pos != 0, rng.integers(1, cfg.item_num + 1, size=(B, L)), 0).astype(np.int32)
        return seq, mask, pos, neg

    def pack_call(self, seq: np.ndarray, mask: Optional[np.ndarray], include_user: bool,
                  with_mm: bool = True) -> PackedCall:
        lay = self.layout
        call = lay.calls[include_user]
        B, L = seq.shape
        T = B * L
        flat = seq.reshape(-1).astype(np.int64)
        if include_user:
            m = mask.reshape(-1)
            item_ids = np.where(m == 1, flat, 0)   # model.py:241,243
            user_ids = np.where(m == 2, flat, 0)   # model.py:240,242
        else:
            item_ids, user_ids = flat, None
        ids = np.zeros((T, call.n_single), np.int32)
        ids[:, 0] = item_ids
        ns = len(lay.item_sparse)
        ids[:, 1:1 + ns] = self.item_sparse_values(item_ids)
        arr_off = np.zeros((call.n_array, T + 1), np.int32)
        arr_vals: List[np.ndarray] = []
        if include_user:
            ids[:, 1 + ns] = user_ids
            ids[:, 2 + ns:2 + ns + len(lay.user_sparse)] = self.user_sparse_values(user_ids)
            base = 0
            for j, k in enumerate(lay.user_array):
                lens, vals = self.user_array_values(user_ids, j, k)
                off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64) + base
                arr_off[j] = off
                keep = np.arange(vals.shape[1])[None, :] < lens[:, None]
                arr_vals.append(vals[keep])
                base = int(off[-1])
        arr_val = np.concatenate(arr_vals).astype(np.int32) if arr_vals else np.zeros((0,), np.int32)
        mm_x = []
        if with_mm:
            for j, (k, d) in enumerate(lay.item_emb_feat.items()):
                mm_x.append(self.mm_vectors(item_ids, j, d))
        return PackedCall(B, L, include_user, ids, arr_off, arr_val, mm_x,
                          seq=seq.astype(np.int32), mask=None if mask is None else mask.astype(np.int32))

    def to_dicts(self, pc: PackedCall) -> list:
        return packed_to_dicts(self.layout, pc)

    def make_step(self, step_seed: int = 0, with_dicts: bool = False, with_mm: bool = True,
                  with_upstream: bool = True) -> SynthStep:
        rng = np.random.Generator(np.random.PCG64(self.seed * 7919 + step_seed))
        seq, mask, pos, neg = self.sample_sequences(rng)
        calls = [self.pack_call(seq, mask, True, with_mm), self.pack_call(pos, None, False, with_mm),
                 self.pack_call(neg, None, False, with_mm)]
        dicts = [self.to_dicts(c) for c in calls] if with_dicts else None
        upstream = None
        if with_upstream:
            r2 = np.random.Generator(np.random.PCG64(self.seed * 7919 + step_seed + 1))
            upstream = [r2.standard_normal((self.cfg.B, self.cfg.L, self.cfg.H)).astype(np.float32) for _ in calls]
        return SynthStep(self.cfg, self.layout, calls, dicts, upstream)


def packed_to_dicts(lay: FeatureLayout, pc: PackedCall) -> list:
    """Reference-form ``feature_array`` (list[B] of object-array[L] of dict) for a packed call.

    Every dict carries every feature id, as ``fill_missing_feat`` guarantees (dataset.py:235-265):
    sparse -> int, array -> list[int] (default ``[0]``), mm -> float32 vector (default zeros).
    """
    B, L = pc.B, pc.L
    ns = len(lay.item_sparse)
    item_keys, user_keys = list(lay.item_sparse), list(lay.user_sparse)
    arr_keys, mm_keys = list(lay.user_array), list(lay.item_emb_feat)
    out = []
    for b in range(B):
        row = np.empty([L], dtype=object)
        for l in range(L):
            t = b * L + l
            d = {}
            for j, k in enumerate(item_keys):
                d[k] = int(pc.ids[t, 1 + j])
            if pc.include_user:
                for j, k in enumerate(user_keys):
                    d[k] = int(pc.ids[t, 2 + ns + j])
                for j, k in enumerate(arr_keys):
                    lo, hi = int(pc.arr_off[j, t]), int(pc.arr_off[j, t + 1])
                    d[k] = [int(x) for x in pc.arr_val[lo:hi]] if hi > lo else [0]
            else:
                for k in user_keys:
                    d[k] = 0
                for k in arr_keys:
                    d[k] = [0]
            for j, k in enumerate(mm_keys):
                d[k] = pc.mm_x[j][t].copy() if pc.mm_x else np.zeros(lay.item_emb_feat[k], np.float32)
            row[l] = d
        out.append(row)
    return out


# BASELINE.json configs (SURVEY.md §8 sizes)
def config_c1() -> SynthConfig:
    return SynthConfig(B=128, L=101, H=32, item_num=100_000, user_num=20_000, alpha=1.05)


def config_c2(B: int = 1024) -> SynthConfig:
    return SynthConfig(B=B, L=101, H=64, item_num=5_000_000, user_num=1_000_000, alpha=1.05)


def config_c3(B: int = 1024) -> SynthConfig:
    return SynthConfig(B=B, L=101, H=64, item_num=5_000_000, user_num=1_000_000, alpha=1.05, mm_ids=("81", "82"))


def config_c4(B: int = 1024, alpha: float = 1.05) -> SynthConfig:
    return SynthConfig(B=B, L=101, H=64, item_num=50_000_000, user_num=50_000_000, alpha=alpha)


def config_c5(B: int = 1024) -> SynthConfig:
    return config_c4(B, alpha=1.2)
