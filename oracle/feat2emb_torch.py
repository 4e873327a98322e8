"""CPU ORACLE #2 — the reference's op sequence on torch CPU tensors, driven by the slot layout.

TEST INFRASTRUCTURE ONLY (see oracle/feat2emb_numpy.py for the rules). Used (a) as a second,
independent check of the numpy restatement through torch's own autograd and ``torch.optim.AdamW``,
and (b) as the multi-threaded CPU baseline that ``bench.py`` times (``cpu_baseline.kind = "port"``):
it runs exactly the library calls the reference's ``feat2emb`` makes (aten::embedding, sum, cat,
addmm, relu — SURVEY.md §2.2 K1-K7) on pre-tensorized inputs, i.e. SURVEY.md §6 figure (ii), the
reference path without its Python dict walk.

Follows model/BaseLine/model.py:115-116,138-139,158-167 (parameters) and :226-310 (op order);
written from the op semantics over ``FeatureLayout`` slots, not copied.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


class TorchOracle(torch.nn.Module):
    def __init__(self, layout):
        super().__init__()
        self.layout = layout
        H = layout.H
        P = torch.nn.Parameter
        self.p = torch.nn.ParameterDict()
        self._names: List[str] = []
        for t in layout.tables:
            self._add(f"{t.name}.weight", torch.zeros(t.rows, H))
        for k, d in layout.item_emb_feat.items():
            self._add(f"emb_transform.{k}.weight", torch.zeros(H, d))
            self._add(f"emb_transform.{k}.bias", torch.zeros(H))
        self._add("itemdnn.weight", torch.zeros(H, layout.item_dim))
        self._add("itemdnn.bias", torch.zeros(H))
        self._add("userdnn.weight", torch.zeros(H, layout.user_dim))
        self._add("userdnn.bias", torch.zeros(H))

    def _add(self, name, t):
        self.p[name.replace(".", "/")] = torch.nn.Parameter(t)
        self._names.append(name)

    def param(self, name: str) -> torch.nn.Parameter:
        return self.p[name.replace(".", "/")]

    def load_numpy(self, params: Dict[str, np.ndarray]):
        with torch.no_grad():
            for n in self._names:
                self.param(n).copy_(torch.from_numpy(np.ascontiguousarray(params[n])))

    def to_numpy(self) -> Dict[str, np.ndarray]:
        return {n: self.param(n).detach().numpy().copy() for n in self._names}

    def grads_numpy(self) -> Dict[str, np.ndarray]:
        return {n: self.param(n).grad.numpy().copy() for n in self._names if self.param(n).grad is not None}

    def table_params(self):
        return [self.param(f"{t.name}.weight") for t in self.layout.tables]

    def feat2emb(self, seq: torch.Tensor, tensors: Dict[str, torch.Tensor], mask: Optional[torch.Tensor],
                 include_user: bool, return_concat: bool = False):
        lay = self.layout
        if include_user:
            ids_main = {"item_id": (mask == 1) * seq, "user_id": (mask == 2) * seq}
        else:
            ids_main = {"item_id": seq}
        parts = {0: [], 1: []}
        for s in lay.calls[include_user].slots:
            if s.kind == 2:
                y = F.linear(tensors[s.name], self.param(f"emb_transform.{s.name}.weight"),
                             self.param(f"emb_transform.{s.name}.bias"))
            else:
                w = self.param(f"{lay.tables[s.table].name}.weight")
                ids = ids_main[s.name] if s.name in ids_main else tensors[s.name]
                y = F.embedding(ids, w, padding_idx=0)
                if s.kind == 1:
                    y = y.sum(2)
            parts[s.side].append(y)
        item_cat = torch.cat(parts[0], dim=2)
        out = torch.relu(F.linear(item_cat, self.param("itemdnn.weight"), self.param("itemdnn.bias")))
        user_cat = None
        if include_user:
            user_cat = torch.cat(parts[1], dim=2)
            out = out + torch.relu(F.linear(user_cat, self.param("userdnn.weight"), self.param("userdnn.bias")))
        if return_concat:
            return out, item_cat, user_cat
        return out


def tensors_to_torch(tensors: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in tensors.items()}
