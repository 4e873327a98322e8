"""Staging + import of the UNMODIFIED reference model for the benchmark's reference arms and the install() tests.

The reference (Puiching-Memory/Tencent_Recommendation_2025) is pure Python and has no setup.py / pyproject, so the
"offline install" of the bench contract is a verbatim copy of the files the hot path lives in:

    /root/reference/model/BaseLine/{model.py,dataset.py}  ->  baseline/_ref/BaseLine/
    /root/reference/model/BaseLineO1/model.py             ->  baseline/_ref/BaseLineO1/

``baseline/_ref/`` is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored, so
the copy travels to the GPU box, where /root/reference does not exist. ``stage()`` runs from ``__graft_entry__.build``
whenever /root/reference is present; nothing here is imported by the product package.

O1's ``dataset.py`` runs ``pip install`` at import: it is never staged; ``model.py`` only needs
``dataset.save_emb`` (model.py:7 / O1 model.py:10), which the BaseLine ``dataset.py`` (no side effects) provides.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/model"
REF_DST = os.path.join(HERE, "_ref")
FILES = [("BaseLine", "model.py"), ("BaseLine", "dataset.py"), ("BaseLineO1", "model.py")]


def stage(force: bool = False) -> bool:
    """Copy the reference files into baseline/_ref (only where /root/reference exists). True when staged files exist."""
    if os.path.isdir(REF_SRC):
        for sub, name in FILES:
            src, dst = os.path.join(REF_SRC, sub, name), os.path.join(REF_DST, sub, name)
            if force or not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DST, sub, name)) for sub, name in FILES)


def _load_file(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_model_class(variant: str = "BaseLine"):
    """``BaselineModel`` of the staged reference (variant 'BaseLine' or 'BaseLineO1'), imported unmodified."""
    if not available():
        raise FileNotFoundError("baseline/_ref is empty: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "in the build container (it stages the reference files)")
    prev = sys.modules.get("dataset")
    try:
        try:
            ds = _load_file(os.path.join(REF_DST, "BaseLine", "dataset.py"), "dataset")
        except Exception:       # e.g. a missing optional import inside dataset.py: model.py only needs save_emb
            ds = types.ModuleType("dataset")
            ds.save_emb = lambda *a, **k: None
        sys.modules["dataset"] = ds
        mod = _load_file(os.path.join(REF_DST, variant, "model.py"), f"tgr_ref_{variant}")
    finally:
        if prev is not None:
            sys.modules["dataset"] = prev
        else:
            sys.modules.pop("dataset", None)
    return mod.BaselineModel


def build_model(cfg, device: str, variant: str = "BaseLine", seed: int = 0):
    """Reference ``BaselineModel`` for a synth config, parameters drawn like bench.init_module (N(0, 0.05) matrices,
    N(0, 0.1) vectors, padding rows zero) so both arms run on statistically identical tables."""
    import torch
    Model = load_model_class(variant)
    args = types.SimpleNamespace(device=device, norm_first=False, maxlen=cfg.L - 1, hidden_units=cfg.H, num_blocks=1,
                                 num_heads=1, dropout_rate=0.0)
    model = Model(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.05, generator=g)
            else:
                p.normal_(0.0, 0.1, generator=g)
        model.item_emb.weight[0].zero_()
        model.user_emb.weight[0].zero_()
        for k in model.sparse_emb:
            model.sparse_emb[k].weight[0].zero_()
    return model


def hot_params(model):
    """Parameters of the hot path (AdamW restricted to them)."""
    keep = ("item_emb", "user_emb", "sparse_emb", "emb_transform", "itemdnn", "userdnn")
    return [p for n, p in model.named_parameters() if n.split(".")[0] in keep]
