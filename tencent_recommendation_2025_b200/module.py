"""Drop-in for ``BaselineModel.feat2emb`` (+ backward + embedding-row update).

Two ways in, same engine underneath:

  * ``BaselineEmbedding`` — a standalone nn.Module with the reference's constructor signature
    ``(user_num, item_num, feat_statistics, feat_types, args)`` and the hot path's attribute names
    (``item_emb``, ``user_emb``, ``sparse_emb[fid]``, ``emb_transform[fid]``, ``itemdnn``, ``userdnn``;
    model/BaseLine/model.py:104-167), so ``named_parameters()`` / ``state_dict()`` keys, the xavier /
    row-0 init of main.py:95-111 and checkpoints interchange with the reference for those keys.
  * ``install(model)`` — graft onto an existing reference ``BaselineModel``: its ``feat2emb`` is
    replaced by the CUDA path operating on the model's OWN parameters; trunk, loss and training loop
    stay untouched (model.py:324,376-377,425 keep calling ``self.feat2emb`` with the same arguments).

``feat2emb(seq, feature_array, mask=None, include_user=False)`` keeps the reference signature and
returns ``[B, L, H]`` on the module's device; ``feat2emb_packed`` takes a pre-tensorized
``PackedBatch`` (the fast entry the benchmark times). itemdnn / userdnn + ReLU + add stay torch
calls and consume the fused concat buffers.
"""
from __future__ import annotations

import types
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import EmbeddingEngine, GatherConcatFn
from .factored import FactoredEngine, FactoredFn
from .layout import FeatureLayout
from .packed import HostPacked, PackedBatch, pack_from_dicts, to_device
from .synth import PackedCall


def _concat_dtype() -> torch.dtype:
    # Under bf16 autocast the reference's fp32 concat is rounded to bf16 by the autocast Linear that
    # consumes it; writing bf16 directly feeds itemdnn bit-identical inputs.
    if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    return torch.float32


class _FeatEmbMixin:
    """feat2emb family shared by the standalone module and by ``install``-ed reference models."""

    _tgr_engine: EmbeddingEngine
    _tgr_layout: FeatureLayout

    # -- reference-compatible tensorizer (model.py:186-224), kept for callers that use it directly
    def feat2tensor(self, seq_feature, k):
        lay = self._tgr_layout
        is_array = k in lay.item_array or k in lay.user_array
        B = len(seq_feature)
        if is_array:
            max_a = max(max(len(tok[k]) for tok in row) for row in seq_feature)
            max_l = max(len(row) for row in seq_feature)
            out = np.zeros((B, max_l, max_a), dtype=np.int64)
            for i, row in enumerate(seq_feature):
                for j, tok in enumerate(row):
                    v = tok[k][:max_a]
                    out[i, j, :len(v)] = v
        else:
            max_l = max(len(row) for row in seq_feature)
            out = np.zeros((B, max_l), dtype=np.int64)
            for i, row in enumerate(seq_feature):
                out[i] = [tok[k] for tok in row]
        return torch.from_numpy(out).to(self.dev)

    def pack(self, seq, feature_array, mask=None, include_user=False, mm_dtype=torch.float32) -> PackedBatch:
        """Dict-form call -> device PackedBatch: one pass over the dicts, one pinned H2D copy."""
        pc = pack_from_dicts(self._tgr_layout, seq, feature_array, mask, include_user)
        return to_device(self._tgr_layout, pc, self._tgr_engine._device(), mm_dtype=mm_dtype)

    @torch.compiler.disable
    def feat2emb(self, seq, feature_array, mask=None, include_user=False):
        """Same signature and result as model/BaseLine/model.py:226-310. ``feature_array`` may also be an already packed
        call (``PackedBatch`` / ``HostPacked`` from ``packed.PackingCollate`` / ``PackedCall``): the dict walk is then
        skipped, ``seq`` / ``mask`` only have to agree in shape (their values are inside the packed ids)."""
        if isinstance(feature_array, (PackedBatch, HostPacked, PackedCall)):
            return self.feat2emb_packed(self._resolve_packed(seq, feature_array, include_user))
        return self.feat2emb_packed(self.pack(seq, feature_array, mask, include_user))

    def _resolve_packed(self, seq, packed, include_user) -> PackedBatch:
        if bool(packed.include_user) != bool(include_user):
            raise ValueError(f"packed call was built with include_user={packed.include_user}, feat2emb got {include_user}")
        if seq is not None and tuple(seq.shape) != (packed.B, packed.L):
            raise ValueError(f"seq is {tuple(seq.shape)}, the packed call holds [{packed.B}, {packed.L}]")
        if isinstance(packed, PackedBatch):
            return packed
        eng = self._tgr_engine
        dev = eng._device()
        if isinstance(packed, PackedCall):
            return to_device(self._tgr_layout, packed, dev)
        group = packed.group
        if group is None:
            return packed.upload(dev)
        first = group.device is None
        pb = group.batch_of(packed, dev)
        if first and torch.is_grad_enabled():
            self.prefetch(group.device)      # factored path: the step's calls become one group; no-op on the concat path
        return pb

    @torch.compiler.disable
    def feat2emb_packed(self, pb: PackedBatch):
        eng = self._tgr_engine
        lay = self._tgr_layout
        params = [t for t in eng.tables]
        for k in lay.item_emb_feat:
            params += [self.emb_transform[k].weight, self.emb_transform[k].bias]
        if getattr(eng, "path", "concat") == "factored":
            # itemdnn / userdnn + ReLU + add (model.py:303-307) happen inside the factored kernels
            params += [self.itemdnn.weight, self.itemdnn.bias, self.userdnn.weight, self.userdnn.bias]
            needs = torch.is_grad_enabled() and any(p.requires_grad for p in params)
            out = FactoredFn.apply(eng, pb, needs, *params).view(pb.B, pb.L, -1)
            if _concat_dtype() == torch.bfloat16:
                out = out.to(torch.bfloat16)   # the autocast Linear of the reference returns bf16
            return out
        item_cat, user_cat = GatherConcatFn.apply(eng, pb, _concat_dtype(), *params)
        B, L = pb.B, pb.L
        out = torch.relu(self.itemdnn(item_cat.view(B, L, -1)))            # model.py:303
        if pb.include_user:
            out = out + torch.relu(self.userdnn(user_cat.view(B, L, -1)))  # model.py:306-307
        return out

    def prefetch(self, pbs):
        """Optional step-level fast path of the factored engine: prepare all of a step's calls as ONE group (one
        sort / dedup, rows shared between the seq / pos / neg calls projected once). No-op on the concat path."""
        eng = self._tgr_engine
        if getattr(eng, "path", "concat") == "factored":
            eng.prefetch(list(pbs))

    def fused_step(self, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2, grad_scale=1.0, dense=False):
        """Row update of the step's touched rows. ``dense=True`` (factored path, after ``own_dense_parameters()``): the same
        AdamW also updates itemdnn / userdnn / emb_transform from the gradients the kernels accumulated."""
        if dense:
            return self._tgr_engine.fused_step(lr, betas, eps, weight_decay, grad_scale, dense=True)
        return self._tgr_engine.fused_step(lr, betas, eps, weight_decay, grad_scale)

    def own_dense_parameters(self, own: bool = True):
        """Fused mode, factored path: the engine also owns the path's Linear layers (model.py:150-151,166-167). Their
        gradients are no longer handed to autograd (``.grad`` stays None — leave them out of the outer optimizer) and
        ``fused_step(dense=True)`` updates them in one launch next to the row update."""
        eng = self._tgr_engine
        if getattr(eng, "path", "concat") != "factored" or eng.mode != "fused":
            raise ValueError("own_dense_parameters needs the factored path in fused mode")
        eng.own_dense = bool(own)

    def save_item_emb(self, item_ids, retrieval_ids, feat_dict, save_path, batch_size=1024):
        """Candidate-embedding sweep with the reference's signature and outputs (model/BaseLine/model.py:402-433):
        ``feat2emb`` over ``[1, n]`` id blocks with ``[np.array(dicts)]`` features, then ``embedding.fbin`` (float32
        [N, H]) and ``id.u64bin`` (uint64 [N, 1]) in ``save_path``."""
        import os

        from .binfmt import save_emb
        all_embs = []
        with torch.no_grad():
            for start in range(0, len(item_ids), batch_size):
                end = min(start + batch_size, len(item_ids))
                item_seq = torch.tensor(item_ids[start:end], device=self.dev).unsqueeze(0)
                batch_feat = np.array([feat_dict[i] for i in range(start, end)], dtype=object)
                emb = self.feat2emb(item_seq, [batch_feat], include_user=False).squeeze(0)
                all_embs.append(emb.detach().float().cpu().numpy().astype(np.float32))
        final_ids = np.array(retrieval_ids, dtype=np.uint64).reshape(-1, 1)
        final_embs = np.concatenate(all_embs, axis=0) if all_embs else np.zeros((0, self._tgr_layout.H), np.float32)
        save_emb(final_embs, os.path.join(save_path, "embedding.fbin"))
        save_emb(final_ids, os.path.join(save_path, "id.u64bin"))

    def save_item_emb_resident(self, store, item_ids, retrieval_ids, save_path, chunk: int = 1 << 16) -> dict:
        """The candidate-embedding sweep (model/BaseLine/model.py:402-433) as a STREAM: the item features
        are resident in HBM (``resident.ResidentItemFeatures``), so a chunk of the sweep is its item ids alone — uploaded on
        a copy stream one chunk ahead, expanded and embedded on the device (forward only), and copied back into a ring of
        pinned buffers that a writer drains into ``embedding.fbin`` while the next chunks compute. No per-chunk host sync, no
        dict walk, never more than three chunks of output on the host. Same files as ``save_item_emb``; returns timings."""
        import os
        import time

        from .binfmt import EmbWriter, save_emb
        from .resident import ResidentFeeder
        ids = np.ascontiguousarray(item_ids, np.int64).reshape(-1)
        N, H = ids.size, self._tgr_layout.H
        dev = store.device
        feeder = ResidentFeeder(store, slots=3)
        ring = [torch.empty((chunk, H), dtype=torch.float32, pin_memory=True) for _ in range(3)]
        done = [None, None, None]
        pend = []                                     # (ring slot, rows) whose D2H copy is in flight, oldest first
        t0 = time.perf_counter()
        starts = list(range(0, N, chunk))
        with EmbWriter(os.path.join(save_path, "embedding.fbin"), N, H) as out, torch.no_grad():
            if starts:
                feeder.submit([store.slim_items(ids[0:chunk])])
            for ci, a in enumerate(starts):
                n = min(chunk, N - a)
                pb = feeder.take()[0]
                if ci + 1 < len(starts):
                    b = starts[ci + 1]
                    feeder.submit([store.slim_items(ids[b:b + chunk])])      # next chunk's ids travel while this one computes
                emb = self.feat2emb_packed(pb).reshape(n, H)
                feeder.retire()
                slot = ci % 3
                if done[slot] is not None:                                    # the writer is two chunks behind the GPU
                    s_, n_ = pend.pop(0)
                    done[s_].synchronize()
                    out.append(ring[s_][:n_].numpy())
                ring[slot][:n].copy_(emb, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                done[slot] = ev
                pend.append((slot, n))
            for s_, n_ in pend:
                done[s_].synchronize()
                out.append(ring[s_][:n_].numpy())
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        save_emb(np.array(retrieval_ids, dtype=np.uint64).reshape(-1, 1), os.path.join(save_path, "id.u64bin"))
        return {"items": N, "seconds": dt, "items_per_s": N / dt if dt > 0 else float("inf"), "chunk": chunk}

    def check_padding_rows(self):
        """Row 0 of every table must be all-zero (the reference keeps it so: main.py:106-111, padding_idx
        gradient masking, AdamW fixed point). Array pooling drops padding ids on that basis."""
        for p, t in zip(self._tgr_engine.tables, self._tgr_layout.tables):
            if bool(torch.any(p.data[0] != 0)):
                raise ValueError(f"{t.name}: padding row 0 is not zero")


def _table_params(mod, layout: FeatureLayout) -> List[torch.nn.Parameter]:
    out = []
    for t in layout.tables:
        if t.name in ("item_emb", "user_emb"):
            out.append(getattr(mod, t.name).weight)
        else:
            out.append(mod.sparse_emb[t.name.split(".", 1)[1]].weight)
    return out


def _make_engine(mod, lay: FeatureLayout, mode: str, path: str):
    """path 'concat': fused gather/pool/concat kernels + the caller's torch itemdnn/userdnn;
    path 'factored': DNN folded into the deduplicated rows (§8(f) N4) — same results, no concat buffers."""
    if path == "concat":
        return EmbeddingEngine(lay, _table_params(mod, lay), dict(mod.emb_transform.items()), mode)
    if path == "factored":
        return FactoredEngine(lay, _table_params(mod, lay), dict(mod.emb_transform.items()),
                              {"item": mod.itemdnn, "user": mod.userdnn}, mode)
    raise ValueError("path must be 'concat' or 'factored'")


class BaselineEmbedding(_FeatEmbMixin, torch.nn.Module):
    """The hot-path slice of the reference ``BaselineModel`` with the CUDA ``feat2emb``."""

    def __init__(self, user_num, item_num, feat_statistics, feat_types, args, mode: str = "parity",
                 path: str = "concat"):
        super().__init__()
        self.user_num, self.item_num = user_num, item_num
        self.dev = args.device
        H = args.hidden_units
        lay = FeatureLayout(user_num, item_num, feat_statistics, feat_types, H)
        self._tgr_layout = lay
        # same declarations, same insertion order as model.py:115-116,158-167
        self.item_emb = torch.nn.Embedding(item_num + 1, H, padding_idx=0)
        self.user_emb = torch.nn.Embedding(user_num + 1, H, padding_idx=0)
        self.sparse_emb = torch.nn.ModuleDict()
        self.emb_transform = torch.nn.ModuleDict()
        self.userdnn = torch.nn.Linear(lay.user_dim, H)
        self.itemdnn = torch.nn.Linear(lay.item_dim, H)
        for group in (lay.user_sparse, lay.item_sparse, lay.item_array, lay.user_array):
            for k, vocab in group.items():
                self.sparse_emb[k] = torch.nn.Embedding(vocab + 1, H, padding_idx=0)
        for k, d in lay.item_emb_feat.items():
            self.emb_transform[k] = torch.nn.Linear(d, H)
        self._tgr_engine = _make_engine(self, lay, mode, path)

    @property
    def layout(self) -> FeatureLayout:
        return self._tgr_layout

    @property
    def engine(self) -> EmbeddingEngine:
        return self._tgr_engine

    def dense_parameters(self):
        """Parameters the outer (dense) optimizer should own in fused mode: everything but the tables."""
        tabs = {id(p) for p in self._tgr_engine.tables}
        return [p for p in self.parameters() if id(p) not in tabs]


def install(model, optimizer: Optional[torch.optim.Optimizer] = None, mode: str = "parity", path: str = "concat",
            scaler=None):
    """Replace ``model.feat2emb`` of a reference-style ``BaselineModel`` with the CUDA path, in place.

    The model keeps its own nn.Embedding / nn.Linear parameters (so init, checkpoints and, in parity
    mode, the optimizer are untouched). In fused mode pass the optimizer: after every
    ``optimizer.step()`` a post-hook applies the queued row updates with that optimizer's lr / betas /
    eps / weight_decay; the tables receive no ``.grad`` so the dense optimizer skips them. The reference routes
    every step through ``scaler.step(optimizer)`` (main.py:189): pass that ``GradScaler`` as ``scaler`` and the row
    update is un-scaled by the same factor (``grad_scale = 1 / scale``), and a step the scaler SKIPS (inf / NaN
    gradients: ``optimizer.step()`` never runs, so the hook never fires) drops the queued row gradients instead of
    leaving them to be merged into the next step.
    """
    H = model.item_emb.embedding_dim
    feat_types = {
        "user_sparse": list(model.USER_SPARSE_FEAT), "item_sparse": list(model.ITEM_SPARSE_FEAT),
        "user_array": list(model.USER_ARRAY_FEAT), "item_array": list(model.ITEM_ARRAY_FEAT),
        "item_emb": list(model.ITEM_EMB_FEAT), "user_continual": list(model.USER_CONTINUAL_FEAT),
        "item_continual": list(model.ITEM_CONTINUAL_FEAT),
    }
    stats: Dict[str, int] = {}
    for d in (model.USER_SPARSE_FEAT, model.ITEM_SPARSE_FEAT, model.USER_ARRAY_FEAT, model.ITEM_ARRAY_FEAT):
        stats.update(d)
    lay = FeatureLayout(model.user_num, model.item_num, stats, feat_types, H)
    model._tgr_layout = lay
    model._tgr_engine = _make_engine(model, lay, mode, path)
    for name, fn in vars(_FeatEmbMixin).items():     # every method of the mixin, private helpers included
        if callable(fn) and not name.startswith("__"):
            setattr(model, name, types.MethodType(fn, model))
    if mode == "fused":
        if optimizer is None:
            raise ValueError("fused mode needs the optimizer whose step() should trigger the row update")

        def _post_step(opt, args, kwargs):
            g = opt.param_groups[0]
            scale = float(scaler.get_scale()) if scaler is not None and scaler.is_enabled() else 1.0
            model._tgr_engine.fused_step(g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"], grad_scale=1.0 / scale)

        model._tgr_hook = optimizer.register_step_post_hook(_post_step)
        if scaler is not None:
            inner = scaler.step

            def _scaler_step(opt, *a, **kw):
                out = inner(opt, *a, **kw)
                if opt is optimizer:
                    model._tgr_engine.discard_pending()    # non-empty only when the scaler skipped optimizer.step()
                return out

            scaler.step = _scaler_step
    return model
