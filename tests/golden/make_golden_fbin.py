"""Generates tests/golden/fbin_*.bin with the UNMODIFIED reference writer (model/BaseLine/dataset.py:421-434 save_emb):
a float32 [7, 5] embedding block (`embedding.fbin` layout) and a uint64 [7, 1] id block (`id.u64bin` layout), as
model.py:431-433 writes them. Run in the build container:  python tests/golden/make_golden_fbin.py"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_dataset", "/root/reference/model/BaseLine/dataset.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(0)
emb = rng.standard_normal((7, 5)).astype(np.float32)
ids = np.arange(100, 107, dtype=np.uint64).reshape(-1, 1)
ref.save_emb(emb, os.path.join(HERE, "fbin_embedding.bin"))
ref.save_emb(ids, os.path.join(HERE, "fbin_ids.bin"))
np.savez(os.path.join(HERE, "fbin_inputs.npz"), emb=emb, ids=ids)
