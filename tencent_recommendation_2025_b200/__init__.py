"""B200-native sparse-feature embedding path for the TencentGR baseline models.

Drop-in for ``BaselineModel.feat2emb`` (+ backward + embedding-row update) of
Puiching-Memory/Tencent_Recommendation_2025 (model/BaseLine/model.py:226-310,
model/BaseLineO1/model.py:327-416). Hand-written CUDA for sm_100a behind a C ABI
(``include/tgr_embed.h``); host side stays Python/PyTorch. No CPU fallback: the product path
raises if ``libtgr_embed.so`` is missing.
"""
from .layout import FeatureLayout, DEFAULT_FEAT_TYPES, EMB_SHAPE_DICT, default_feat_statistics  # noqa: F401



def __getattr__(name):      # lazy: importing the package must not need torch extensions or a built library
    if name in ("install", "BaselineEmbedding"):
        from . import module
        return getattr(module, name)
    if name == "PackingCollate":
        from .packed import PackingCollate
        return PackingCollate
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


__all__ = ["FeatureLayout", "DEFAULT_FEAT_TYPES", "EMB_SHAPE_DICT", "default_feat_statistics", "install",
           "BaselineEmbedding", "PackingCollate"]
