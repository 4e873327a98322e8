"""CPU ORACLE — numpy restatement of the reference's sparse-feature embedding path.

TEST INFRASTRUCTURE ONLY. Nothing under ``tencent_recommendation_2025_b200/`` imports this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may. It is the checker, never the product path.

What it restates (paths relative to /root/reference; BaseLineO1 is identical):
  feat2tensor ............ model/BaseLine/model.py:186-224
  feat2emb forward ....... model/BaseLine/model.py:226-310
  backward ............... autograd of the ops above. The arithmetic lives in PyTorch (third-party,
                           unpinned by the reference — README.md:13; this image has torch 2.11.0):
                           embedding_dense_backward = per-row fp32 sum in ascending flat-position
                           order, padding row skipped (torch/_decomp/decompositions.py:1278-1306);
                           Linear/ReLU/cat/sum backward are the textbook formulas.
  AdamW .................. model/BaseLine/main.py:131 -> torch/optim/adam.py:416-419,457,476,531-547
                           (_single_tensor_adam, decoupled decay, eps outside the bias-corrected sqrt)

PINNING: the reference has no tests or golden vectors, so this oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by importing
/root/reference/model/{BaseLine,BaseLineO1}/model.py unmodified (tests/golden/make_golden.py,
fixtures tests/golden/*.npz); tests/test_oracle_golden.py checks every fixture.

No part of this file is copied from the reference: it is written from the op semantics.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------------
# feat2tensor (model.py:186-224)
# ---------------------------------------------------------------------------------------------
def feat2tensor(seq_feature, k: str, is_array: bool) -> np.ndarray:
    """list[B] of indexable[L] of dict -> int64 [B, L] (sparse) or [B, L, A_max] (array, zero padded).

    A_max is the longest list in THIS batch (model.py:199-207); sparse features require every
    sequence to have the same length (the numpy row assignment at :222 raises otherwise).
    """
    B = len(seq_feature)
    if is_array:
        max_a, max_l = 0, 0
        for i in range(B):
            vals = [tok[k] for tok in seq_feature[i]]
            max_l = max(max_l, len(vals))
            max_a = max(max_a, max(len(v) for v in vals))
        out = np.zeros((B, max_l, max_a), np.int64)
        for i in range(B):
            for j, tok in enumerate(seq_feature[i]):
                v = tok[k]
                n = min(len(v), max_a)
                out[i, j, :n] = v[:n]
        return out
    max_l = max(len(seq_feature[i]) for i in range(B))
    out = np.zeros((B, max_l), np.int64)
    for i in range(B):
        out[i] = [tok[k] for tok in seq_feature[i]]   # ValueError on ragged input, as the reference
    return out


def mm_fill(seq_feature, k: str, dim: int) -> np.ndarray:
    """model.py:283-293: float32 [B, L, dim], zeros where the dict has no key k."""
    B, L = len(seq_feature), len(seq_feature[0])
    out = np.zeros((B, L, dim), F32)
    for i, row in enumerate(seq_feature):
        for j, tok in enumerate(row):
            if k in tok:
                out[i, j] = tok[k]
    return out


# ---------------------------------------------------------------------------------------------
# tensor-form inputs
# ---------------------------------------------------------------------------------------------
def tensors_from_dicts(layout, feature_array, include_user: bool) -> Dict[str, np.ndarray]:
    t: Dict[str, np.ndarray] = {}
    for k in layout.item_sparse:
        t[k] = feat2tensor(feature_array, k, False)
    for k in layout.item_array:
        t[k] = feat2tensor(feature_array, k, True)
    if include_user:
        for k in layout.user_sparse:
            t[k] = feat2tensor(feature_array, k, False)
        for k in layout.user_array:
            t[k] = feat2tensor(feature_array, k, True)
    for k, d in layout.item_emb_feat.items():
        t[k] = mm_fill(feature_array, k, d)
    return t


def tensors_from_packed(layout, pc) -> Dict[str, np.ndarray]:
    """Same tensor form from a PackedCall (arrays re-padded to the batch's A_max, min 1: default [0])."""
    B, L, T = pc.B, pc.L, pc.T
    t: Dict[str, np.ndarray] = {}
    names = layout.single_slot_names(pc.include_user)
    for c, k in enumerate(names):
        if k in ("item_id", "user_id"):
            continue
        t[k] = pc.ids[:, c].astype(np.int64).reshape(B, L)
    for j, k in enumerate(layout.array_slot_names(pc.include_user)):
        off = pc.arr_off[j].astype(np.int64)
        lens = off[1:] - off[:-1]
        A = max(1, int(lens.max()) if lens.size else 1)
        a = np.zeros((T, A), np.int64)
        for tok in np.nonzero(lens)[0]:
            a[tok, :lens[tok]] = pc.arr_val[off[tok]:off[tok + 1]]
        t[k] = a.reshape(B, L, A)
    for j, k in enumerate(layout.item_emb_feat):
        t[k] = pc.mm_x[j].reshape(B, L, -1).astype(F32)
    return t


# ---------------------------------------------------------------------------------------------
# forward (model.py:226-310)
# ---------------------------------------------------------------------------------------------
def _pool_left_to_right(rows: np.ndarray) -> np.ndarray:
    """[B,L,A,H] -> [B,L,H]; fp32 adds in array order == torch ``.sum(2)`` on CPU."""
    acc = rows[:, :, 0, :].astype(F32).copy()
    for a in range(1, rows.shape[2]):
        acc = (acc + rows[:, :, a, :]).astype(F32)
    return acc


def feat2emb_forward(params: Dict[str, np.ndarray], layout, seq: np.ndarray, tensors: Dict[str, np.ndarray],
                     mask: Optional[np.ndarray], include_user: bool):
    """Returns (out [B,L,H] float32, cache). ``params`` uses the reference's state_dict keys."""
    seq = np.asarray(seq).astype(np.int64)
    cache = {"include_user": include_user, "ids": {}, "B": seq.shape[0], "L": seq.shape[1]}
    item_list, user_list = [], []
    if include_user:
        m = np.asarray(mask)
        uid = (m == 2) * seq                      # model.py:240,242
        iid = (m == 1) * seq                      # model.py:241,243
        user_list.append(params["user_emb.weight"][uid])
        cache["ids"]["user_emb"] = uid
    else:
        iid = seq                                  # model.py:247
    item_list.append(params["item_emb.weight"][iid])
    cache["ids"]["item_emb"] = iid

    groups = [(layout.item_sparse, False, item_list), (layout.item_array, True, item_list)]
    if include_user:
        groups += [(layout.user_sparse, False, user_list), (layout.user_array, True, user_list)]
    for feats, is_arr, dst in groups:              # model.py:267-279
        for k in feats:
            ids = tensors[k]
            w = params[f"sparse_emb.{k}.weight"]
            cache["ids"][f"sparse_emb.{k}"] = ids
            dst.append(_pool_left_to_right(w[ids]) if is_arr else w[ids])
    for k in layout.item_emb_feat:                 # model.py:281-299
        x = tensors[k].astype(F32)
        w, b = params[f"emb_transform.{k}.weight"], params[f"emb_transform.{k}.bias"]
        item_list.append((x @ w.T + b).astype(F32))
        cache.setdefault("mm_x", {})[k] = x

    item_cat = np.concatenate(item_list, axis=2).astype(F32)      # model.py:302
    zi = (item_cat @ params["itemdnn.weight"].T + params["itemdnn.bias"]).astype(F32)
    yi = np.maximum(zi, 0)                                          # model.py:303
    cache.update(item_cat=item_cat, yi=yi)
    out = yi
    if include_user:
        user_cat = np.concatenate(user_list, axis=2).astype(F32)  # model.py:305
        zu = (user_cat @ params["userdnn.weight"].T + params["userdnn.bias"]).astype(F32)
        yu = np.maximum(zu, 0)                                      # model.py:306
        cache.update(user_cat=user_cat, yu=yu)
        out = (yi + yu).astype(F32)                                 # model.py:307
    return out.astype(F32), cache


# ---------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------
def embedding_dense_backward(grad: np.ndarray, ids: np.ndarray, rows: int) -> np.ndarray:
    """Dense [rows, H] fp32; per row a sequential sum in ascending flat position; padding row 0 skipped."""
    H = grad.shape[-1]
    g = np.zeros((rows, H), F32)
    flat = ids.reshape(-1)
    np.add.at(g, flat, grad.reshape(-1, H).astype(F32))   # unbuffered, in index order
    g[0] = 0
    return g


def concat_backward(params, layout, cache, d_out: np.ndarray):
    """dOut [B,L,H] -> (d_item_cat, d_user_cat | None, dense-param grads of itemdnn/userdnn)."""
    H = layout.H
    g: Dict[str, np.ndarray] = {}
    d_out = d_out.astype(F32)
    dzi = (d_out * (cache["yi"] > 0)).astype(F32)
    T = dzi.shape[0] * dzi.shape[1]
    g["itemdnn.weight"] = (dzi.reshape(T, H).T @ cache["item_cat"].reshape(T, -1)).astype(F32)
    g["itemdnn.bias"] = dzi.reshape(T, H).sum(0).astype(F32)
    d_item = (dzi @ params["itemdnn.weight"]).astype(F32)
    d_user = None
    if cache["include_user"]:
        dzu = (d_out * (cache["yu"] > 0)).astype(F32)
        g["userdnn.weight"] = (dzu.reshape(T, H).T @ cache["user_cat"].reshape(T, -1)).astype(F32)
        g["userdnn.bias"] = dzu.reshape(T, H).sum(0).astype(F32)
        d_user = (dzu @ params["userdnn.weight"]).astype(F32)
    return d_item, d_user, g


def feat2emb_backward(params, layout, cache, d_out: np.ndarray) -> Dict[str, np.ndarray]:
    """Dense grads for every hot-path parameter of ONE call, keyed like the state_dict."""
    H = layout.H
    d_item, d_user, g = concat_backward(params, layout, cache, d_out)
    call = layout.calls[cache["include_user"]]
    for s in call.slots:
        d_side = d_item if s.side == 0 else d_user
        chunk = d_side[:, :, s.col:s.col + H]
        if s.kind == 2:
            x = cache["mm_x"][s.name]
            T = x.shape[0] * x.shape[1]
            g[f"emb_transform.{s.name}.weight"] = (chunk.reshape(T, H).T @ x.reshape(T, -1)).astype(F32)
            g[f"emb_transform.{s.name}.bias"] = chunk.reshape(T, H).sum(0).astype(F32)
            continue
        t = layout.tables[s.table]
        ids = cache["ids"][t.name]
        if s.kind == 1:   # grad broadcast over the array axis (backward of sum(2))
            A = ids.shape[2]
            chunk = np.broadcast_to(chunk[:, :, None, :], chunk.shape[:2] + (A, H))
        g[f"{t.name}.weight"] = embedding_dense_backward(np.ascontiguousarray(chunk), ids, t.rows)
    return g


def accumulate(grads_in_autograd_order: Sequence[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    """AccumulateGrad: first grad stored, later ones added in arrival order (fp32)."""
    tot: Dict[str, np.ndarray] = {}
    for g in grads_in_autograd_order:
        for k, v in g.items():
            tot[k] = v.copy() if k not in tot else (tot[k] + v).astype(F32)
    return tot


# ---------------------------------------------------------------------------------------------
# AdamW (torch/optim/adam.py _single_tensor_adam, decoupled weight decay)
# ---------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    """float32 fma(a, b, c): the product of two float32 is exact in float64; one final rounding."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def adamw_dense(w, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-8, wd=1e-2):
    """One dense AdamW step on one tensor; returns new (w, m, v). ``step`` counts from 1.

    Rounding order = torch 2.11 CPU kernels (probed): lerp_ and addcmul_ fuse their last multiply-add.
    """
    w = (w * F32(1 - lr * wd)).astype(F32)                             # adam.py:416-419 (param.mul_)
    m = _fma32(F32(1 - beta1), (g - m).astype(F32), m)                 # adam.py:457  (lerp_)
    v = _fma32((F32(1 - beta2) * g).astype(F32), g, (v * F32(beta2)).astype(F32))   # adam.py:476
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = ((np.sqrt(v) / F32(bc2 ** 0.5)).astype(F32) + F32(eps)).astype(F32)     # adam.py:536-547
    w = (w + ((F32(-step_size) * m).astype(F32) / denom).astype(F32)).astype(F32)
    return w, m, v


def adamw_rows(w, m, v, rows: np.ndarray, g_rows: np.ndarray, step: int, **hp):
    """The sparse ("lazy") row update: the dense formula applied to the touched rows only, in place.

    Equals the reference's dense AdamW on those rows whenever they have the same (w, m, v) going in —
    e.g. at step 1 from zero state. Untouched rows are left alone, whereas dense AdamW would scale
    them by (1 - lr*wd) and keep moving them on stale momentum.
    """
    nw, nm, nv = adamw_dense(w[rows], g_rows, m[rows], v[rows], step, **hp)
    w[rows], m[rows], v[rows] = nw, nm, nv


# ---------------------------------------------------------------------------------------------
# dedup / key building / routing (no reference counterpart: the build's own definitions)
# ---------------------------------------------------------------------------------------------
def build_keys(layout, packed_calls) -> Tuple[np.ndarray, np.ndarray]:
    """Global keys (table.key_base + id, padding dropped) and their source code, in generation order.

    Generation order = call, then slot (layout order), then token (arrays: CSR order). Source code =
    call << 29 | slot_index << 24 | token  (include/tgr_embed.h TGR_SRC_*).
    """
    keys, srcs = [], []
    for c, pc in enumerate(packed_calls):
        call = layout.calls[pc.include_user]
        tok = np.arange(pc.T, dtype=np.int64)
        for si, s in enumerate(call.slots):
            if s.kind == 0:
                ids = pc.ids[:, s.src].astype(np.int64)
                toks = tok
            elif s.kind == 1:
                off = pc.arr_off[s.src].astype(np.int64)
                ids = pc.arr_val[off[0]:off[-1]].astype(np.int64)
                toks = np.repeat(tok, off[1:] - off[:-1])
            else:
                continue
            keep = ids != 0
            keys.append(layout.tables[s.table].key_base + ids[keep])
            srcs.append((c << 29) | (si << 24) | toks[keep])
    if not keys:
        return np.zeros(0, np.uint32), np.zeros(0, np.int64)
    # sources stay int64 here so the oracle also handles more than 8 calls (the C ABI packs 3 call bits)
    return np.concatenate(keys).astype(np.uint32), np.concatenate(srcs).astype(np.int64)


def sort_dedup(keys: np.ndarray):
    """Stable sort by key -> (sorted order, unique keys, segment offsets [U+1], counts)."""
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    uniq, start, counts = np.unique(sk, return_index=True, return_counts=True)
    seg = np.concatenate([start, [sk.size]]).astype(np.int64)
    return order, uniq, seg, counts


def segment_reduce_fp64(layout, packed_calls, d_cats, keys_src=None):
    """fp64 ground truth of the per-unique-row gradient. d_cats[c] = (d_item [T,Di], d_user [T,Du] | None)."""
    keys, srcs = keys_src if keys_src is not None else build_keys(layout, packed_calls)
    order, uniq, seg, _ = sort_dedup(keys)
    H = layout.H
    rows = np.zeros((keys.size, H), np.float64)
    c = (srcs >> 29).astype(np.int64)
    si = ((srcs >> 24) & 31).astype(np.int64)
    tok = (srcs & 0xFFFFFF).astype(np.int64)
    for ci, pc in enumerate(packed_calls):
        call = layout.calls[pc.include_user]
        for sidx, s in enumerate(call.slots):
            sel = (c == ci) & (si == sidx)
            if not sel.any():
                continue
            d_side = d_cats[ci][s.side]
            rows[sel] = d_side.reshape(pc.T, -1)[tok[sel], s.col:s.col + H]
    rows = rows[order]
    out = np.add.reduceat(rows, seg[:-1], axis=0) if uniq.size else np.zeros((0, H))
    return uniq, out


def route(keys: np.ndarray, W: int):
    """Row ownership for W ranks: owner = key mod W, local_row = key div W; stable bucket order.

    Returns (owner, local_row, send_counts [W], bucket_order) where bucket_order lists the positions
    of ``keys`` grouped by owner, ascending key order preserved inside each bucket.
    """
    keys = keys.astype(np.int64)
    owner = keys % W
    local = keys // W
    order = np.argsort(owner, kind="stable")
    counts = np.bincount(owner, minlength=W).astype(np.int64)
    return owner.astype(np.int32), local.astype(np.int64), counts, order
