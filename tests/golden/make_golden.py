"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures are committed; the GPU box never reads /root/reference.

For each case: import model/{BaseLine,BaseLineO1}/model.py (O1 with a stub ``dataset`` module,
SURVEY.md F12 / Appendix A), build ``BaselineModel`` on CPU, apply the reference's init
(model/BaseLine/main.py:95-111) then re-randomise the 1-D parameters (the shipped
init zeroes every bias), call ``feat2emb`` three times exactly as a training step does
(model.py:324,376-377) on list-of-dict inputs, inject seeded upstream gradients at feat2emb's
output, run ``torch.optim.AdamW(lr, betas=(0.9,0.98))`` (main.py:131) for two steps on two
different batches, and store inputs (packed form), parameters, outputs, gradients and updated
parameters/optimizer state.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tencent_recommendation_2025_b200.synth import SynthConfig, SynthWorld, packed_to_dicts  # noqa: E402

REF = "/root/reference/model"


def load_ref(variant: str):
    stub = types.ModuleType("dataset")
    stub.save_emb = lambda *a, **k: None
    sys.modules["dataset"] = stub
    spec = importlib.util.spec_from_file_location(f"ref_{variant}", f"{REF}/{variant}/model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.BaselineModel


def hot_params(model):
    out = {}
    for n, p in model.named_parameters():
        if n.split(".")[0] in ("item_emb", "user_emb", "sparse_emb", "emb_transform", "itemdnn", "userdnn"):
            out[n] = p
    return out


def make_case(name, variant, cfg: SynthConfig, seed, lr, wd, n_steps=2, store_state=True):
    torch.manual_seed(seed)
    Model = load_ref(variant)
    args = types.SimpleNamespace(device="cpu", norm_first=False, maxlen=cfg.L - 1, hidden_units=cfg.H,
                                 num_blocks=1, num_heads=1, dropout_rate=0.0)
    world = SynthWorld(cfg, seed)
    lay = world.layout
    model = Model(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args)
    # reference init (main.py:95-111)
    for _, p in model.named_parameters():
        if p.dim() >= 2:
            torch.nn.init.xavier_normal_(p.data)
        elif p.dim() == 1:
            torch.nn.init.constant_(p.data, 0.0)
    model.pos_emb.weight.data[0, :] = 0
    model.item_emb.weight.data[0, :] = 0
    model.user_emb.weight.data[0, :] = 0
    for k in model.sparse_emb:
        model.sparse_emb[k].weight.data[0, :] = 0
    hp = hot_params(model)
    g = torch.Generator().manual_seed(seed + 17)
    for n, p in hp.items():
        if p.dim() == 1:
            p.data.copy_(0.1 * torch.randn(p.shape, generator=g))
        elif n.split(".")[0] in ("item_emb", "user_emb", "sparse_emb"):
            # xavier std on a tiny table is ~0.1; keep but make sure values are O(0.1) for all tables
            p.data[1:].copy_(0.1 * torch.randn(p.data[1:].shape, generator=g))

    opt = torch.optim.AdamW(list(hp.values()), lr=lr, betas=(0.9, 0.98), weight_decay=wd)
    blob = {"variant": variant, "seed": seed, "lr": lr, "wd": wd, "n_steps": n_steps,
            "B": cfg.B, "L": cfg.L, "H": cfg.H, "item_num": cfg.item_num, "user_num": cfg.user_num,
            "mm_ids": np.array(list(cfg.mm_ids)),
            "stat_keys": np.array(list(cfg.statistics().keys())),
            "stat_vals": np.array(list(cfg.statistics().values()), np.int64)}
    for n, p in hp.items():
        blob[f"param0/{n}"] = p.detach().numpy().copy()
    for step in range(n_steps):
        st = world.make_step(step, with_dicts=True)
        opt.zero_grad(set_to_none=True)
        outs = []
        for c, pc in enumerate(st.calls):
            seq = torch.from_numpy(pc.seq)
            mask = torch.from_numpy(pc.mask) if pc.include_user else None
            outs.append(model.feat2emb(seq, st.dicts[c], mask=mask, include_user=pc.include_user))
        loss = sum((o * torch.from_numpy(r)).sum() for o, r in zip(outs, st.upstream))
        loss.backward()
        for c, pc in enumerate(st.calls):
            pre = f"s{step}/c{c}/"
            blob[pre + "ids"] = pc.ids
            blob[pre + "arr_off"] = pc.arr_off
            blob[pre + "arr_val"] = pc.arr_val
            for j, x in enumerate(pc.mm_x):
                blob[pre + f"mm_x{j}"] = x
            blob[pre + "seq"] = pc.seq
            if pc.include_user:
                blob[pre + "mask"] = pc.mask
            blob[pre + "out"] = outs[c].detach().numpy().copy()
            blob[pre + "upstream"] = st.upstream[c]
        for n, p in hp.items():
            blob[f"s{step}/grad/{n}"] = p.grad.numpy().copy() if p.grad is not None else np.zeros(0, np.float32)
        opt.step()
        for n, p in hp.items():
            if not store_state:
                continue
            blob[f"s{step}/param/{n}"] = p.detach().numpy().copy()
            if step == n_steps - 1:
                blob[f"s{step}/exp_avg/{n}"] = opt.state[p]["exp_avg"].numpy().copy()
                blob[f"s{step}/exp_avg_sq/{n}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(blob)} arrays")


def make_item_sweep(name, cfg: SynthConfig, seed, n):
    """save_item_emb's call shape: int64 seq [1, n], feature_array = [object-array[n] of dict] (model.py:418-425)."""
    torch.manual_seed(seed)
    Model = load_ref("BaseLine")
    args = types.SimpleNamespace(device="cpu", norm_first=False, maxlen=cfg.L - 1, hidden_units=cfg.H,
                                 num_blocks=1, num_heads=1, dropout_rate=0.0)
    world = SynthWorld(cfg, seed)
    model = Model(cfg.user_num, cfg.item_num, cfg.statistics(), cfg.feat_types(), args)
    g = torch.Generator().manual_seed(seed + 3)
    hp = hot_params(model)
    for n_, p in hp.items():
        p.data.copy_(0.1 * torch.randn(p.shape, generator=g))
    for n_ in hp:
        if n_.split(".")[0] in ("item_emb", "user_emb", "sparse_emb"):
            hp[n_].data[0] = 0
    items = np.arange(1, n + 1, dtype=np.int64)[None, :]
    pc = world.pack_call(items, None, False)
    dicts = packed_to_dicts(world.layout, pc)
    with torch.no_grad():
        out = model.feat2emb(torch.from_numpy(items), dicts, include_user=False)
    blob = {"variant": "BaseLine", "seed": seed, "B": 1, "L": n, "H": cfg.H, "item_num": cfg.item_num,
            "user_num": cfg.user_num, "mm_ids": np.array(list(cfg.mm_ids)),
            "stat_keys": np.array(list(cfg.statistics().keys())),
            "stat_vals": np.array(list(cfg.statistics().values()), np.int64),
            "ids": pc.ids, "arr_off": pc.arr_off, "arr_val": pc.arr_val, "seq": items, "out": out.numpy()}
    for j, x in enumerate(pc.mm_x):
        blob[f"mm_x{j}"] = x
    for n_, p in hp.items():
        blob[f"param0/{n_}"] = p.detach().numpy().copy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def make_bf16_case(name, base_name):
    """Same inputs/weights as ``base_name`` (read back from its fixture), reference run under
    ``torch.autocast('cpu', dtype=torch.bfloat16)`` — the bf16 oracle of SURVEY.md F15 / §4."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden
    g = Golden(base_name)
    Model = load_ref(str(g.z["variant"]))
    args = types.SimpleNamespace(device="cpu", norm_first=False, maxlen=g.L - 1, hidden_units=g.H,
                                 num_blocks=1, num_heads=1, dropout_rate=0.0)
    model = Model(g.user_num, g.item_num, g.feat_statistics, g.feat_types, args)
    hp = hot_params(model)
    for n, p in hp.items():
        p.data.copy_(torch.from_numpy(g.z[f"param0/{n}"]))
    outs = []
    calls = g.calls(0)
    for c, pc in enumerate(calls):
        dicts = packed_to_dicts(g.layout, pc)
        seq = torch.from_numpy(pc.seq)
        mask = torch.from_numpy(pc.mask) if pc.include_user else None
        with torch.autocast("cpu", dtype=torch.bfloat16):
            outs.append(model.feat2emb(seq, dicts, mask=mask, include_user=pc.include_user))
    loss = sum((o.float() * torch.from_numpy(r)).sum() for o, r in zip(outs, g.upstream(0)))
    loss.backward()
    blob = {"base": base_name}
    for c, o in enumerate(outs):
        blob[f"out{c}"] = o.detach().float().numpy()
    for n, p in hp.items():
        if p.grad is not None:
            blob[f"grad/{n}"] = p.grad.float().numpy().copy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


SMALL_STATS = {"103": 3, "104": 8, "105": 15, "109": 30,
               "100": 5, "117": 12, "111": 30, "118": 50, "101": 90, "102": 6, "119": 11, "120": 25,
               "114": 40, "112": 80, "121": 4, "115": 15, "122": 28, "116": 45,
               "106": 20, "107": 35, "108": 60, "110": 10}

if __name__ == "__main__":
    if "--bf16-only" in sys.argv:
        make_bf16_case("o1_h64_bf16", "o1_h64")
        sys.exit(0)
    # BaseLine: H=32, mm '81', duplicate-heavy ids (alpha 1.2 on a small table)
    make_case("baseline_h32", "BaseLine",
              SynthConfig(B=6, L=12, H=32, item_num=150, user_num=20, alpha=1.2, mm_ids=("81",), min_len=3,
                          feat_statistics=SMALL_STATS), seed=1, lr=1e-3, wd=1e-2)
    # BaseLineO1: H=64, O1's lr / weight decay (BaseLineO1/main.py:174)
    make_case("o1_h64", "BaseLineO1",
              SynthConfig(B=3, L=10, H=64, item_num=80, user_num=12, alpha=1.05, mm_ids=("81",), min_len=2,
                          feat_statistics=SMALL_STATS), seed=2, lr=5e-3, wd=1e-2, n_steps=1)
    # BaseLineO1 with two mm features (32-d and 1024-d): outputs and gradients only
    make_case("o1_h64_mm2", "BaseLineO1",
              SynthConfig(B=2, L=6, H=64, item_num=40, user_num=6, alpha=1.05, mm_ids=("81", "82"), min_len=2,
                          feat_statistics=SMALL_STATS), seed=5, lr=5e-3, wd=1e-2, n_steps=1, store_state=False)
    # L=102 (maxlen+1 with the reference default --maxlen 101), no mm feature
    make_case("baseline_l102_nomm", "BaseLine",
              SynthConfig(B=2, L=102, H=32, item_num=300, user_num=25, alpha=1.05, mm_ids=(), min_len=20,
                          feat_statistics=SMALL_STATS), seed=3, lr=1e-3, wd=1e-2, n_steps=1, store_state=False)
    make_item_sweep("item_sweep", SynthConfig(B=1, L=37, H=32, item_num=300, user_num=10, mm_ids=("81",),
                                              feat_statistics=SMALL_STATS), seed=4, n=37)
    make_bf16_case("o1_h64_bf16", "o1_h64")
