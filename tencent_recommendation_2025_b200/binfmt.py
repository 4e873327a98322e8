"""The `.fbin` / `.u64bin` wire format of the reference's candidate-embedding dump: what
``save_item_emb`` hands to the external ANN step and what ``infer.py`` reads back.

    uint32 num_points, uint32 num_dimensions, then the row-major array in its own dtype
    (model/BaseLine/dataset.py:421-434 ``save_emb``; float32 for embedding.fbin / query.fbin, uint64 for id.u64bin;
    result files: model/BaseLine/infer.py:51-65 ``read_result_ids``)
"""
from __future__ import annotations

import os
import struct
from typing import Union

import numpy as np
import torch


def save_emb(emb: Union[np.ndarray, torch.Tensor], save_path) -> None:
    """Same bytes as the reference's ``save_emb(emb, save_path)``. A CUDA tensor is copied to the host first."""
    if isinstance(emb, torch.Tensor):
        emb = emb.detach().cpu().numpy()
    if emb.ndim != 2:
        raise ValueError("save_emb expects a [num_points, num_dimensions] array")
    emb = np.ascontiguousarray(emb)
    with open(os.fspath(save_path), "wb") as f:
        f.write(struct.pack("II", emb.shape[0], emb.shape[1]))
        emb.tofile(f)


def load_emb(path, dtype=np.float32) -> np.ndarray:
    with open(os.fspath(path), "rb") as f:
        n, d = struct.unpack("II", f.read(8))
        return np.fromfile(f, dtype=dtype, count=n * d).reshape(n, d)


def read_result_ids(path) -> np.ndarray:
    """uint64 [num_queries, top_k] (model/BaseLine/infer.py:51-65)."""
    return load_emb(path, np.uint64)


class EmbWriter:
    """Incremental writer of the same format: the header is written up front from the known shape, row blocks are appended
    as they arrive (the streaming candidate sweep never holds the whole [N, H] array on the host)."""

    def __init__(self, save_path, num_points: int, num_dimensions: int, dtype=np.float32):
        self.f = open(os.fspath(save_path), "wb")
        self.f.write(struct.pack("II", int(num_points), int(num_dimensions)))
        self.n, self.d, self.dtype, self.written = int(num_points), int(num_dimensions), np.dtype(dtype), 0

    def append(self, rows: np.ndarray) -> None:
        rows = np.ascontiguousarray(rows, self.dtype)
        if rows.ndim != 2 or rows.shape[1] != self.d:
            raise ValueError(f"expected [k, {self.d}] rows")
        rows.tofile(self.f)
        self.written += rows.shape[0]

    def close(self) -> None:
        self.f.close()
        if self.written != self.n:
            raise ValueError(f"wrote {self.written} of {self.n} rows")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.f.close()
        return False
