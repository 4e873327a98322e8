"""Candidate-embedding sweep at scale: python tools/sweep_bench.py [items] [chunk]
save_item_emb_resident over `items` ids of the C2 tables (features resident in HBM), written to /dev/shm; prints one JSON
line: items/s end to end (ids H2D -> expand -> forward -> D2H ring -> .fbin) and the same with the file write replaced by a
no-op, against the output-stream roofline (256 B per item over PCIe)."""
import json, os, sys, tempfile, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from tencent_recommendation_2025_b200 import synth, binfmt
from tencent_recommendation_2025_b200.resident import ResidentItemFeatures

items = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
cfg = synth.config_c2(1024)
world = synth.SynthWorld(cfg, 0)
m = bench.init_module(cfg, torch.device("cuda"), "fused", "factored")
t0 = time.perf_counter()
store = ResidentItemFeatures.from_world(world, "cuda")
t_tab = time.perf_counter() - t0
ids = np.arange(1, min(items, cfg.item_num) + 1, dtype=np.int64)
out_dir = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
res = {}
for label, nowrite in (("warmup", True), ("no_file_write", True), ("end_to_end", False)):
    if nowrite:
        orig = binfmt.EmbWriter.append
        binfmt.EmbWriter.append = lambda self, rows: setattr(self, "written", self.written + rows.shape[0])
    n_use = ids[: 1 << 20] if label == "warmup" else ids
    info = m.save_item_emb_resident(store, n_use, n_use, out_dir, chunk=chunk)
    if nowrite:
        binfmt.EmbWriter.append = orig
    res[label] = info
e = binfmt.load_emb(os.path.join(out_dir, "embedding.fbin"))
assert e.shape == (ids.size, cfg.H) and np.isfinite(e).all()
line = {"bench": "save_item_emb sweep (N2)", "items": int(ids.size), "chunk": chunk, "hidden": cfg.H,
        "end_to_end_items_per_s": res["end_to_end"]["items_per_s"], "end_to_end_s": res["end_to_end"]["seconds"],
        "no_file_write_items_per_s": res["no_file_write"]["items_per_s"], "no_file_write_s": res["no_file_write"]["seconds"],
        "d2h_GBs_no_file_write": ids.size * cfg.H * 4 / res["no_file_write"]["seconds"] / 1e9,
        "resident_table_build_s": round(t_tab, 1),
        "note": "forward-only factored path per chunk of ids (expand from the resident tables, keys/sort/dedup, projection of the "
                "chunk's unique rows, gather-sum), output through a 3-slot pinned ring; the reference walks 1024 dicts per chunk "
                "and syncs on .cpu() (model.py:402-433)"}
print(json.dumps(line))
