"""Pin the CPU oracles against outputs of the unmodified reference (tests/golden/*.npz).

Tolerances: forward / gradient / updated-row values 1e-5 relative (north_star, fp32) with an absolute
floor tied to the tensor's scale; integer work (ids, keys, dedup) bit-exact.
"""
import numpy as np
import pytest
import torch

from golden_util import Golden, TRAIN_CASES
from oracle import feat2emb_numpy as onp
from oracle.feat2emb_torch import TorchOracle, tensors_to_torch
from tencent_recommendation_2025_b200.synth import packed_to_dicts

RTOL = 1e-5


def close(a, b, rtol=RTOL, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    err = np.abs(a - b).max() if b.size else 0.0
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_numpy_oracle_forward_backward_matches_reference(name):
    g = Golden(name)
    lay = g.layout
    params = g.params0()
    step = 0
    calls = g.calls(step)
    grads = []
    for c, pc in enumerate(calls):
        tensors = onp.tensors_from_packed(lay, pc)
        out, cache = onp.feat2emb_forward(params, lay, pc.seq, tensors, pc.mask, pc.include_user)
        close(out, g.outs(step)[c], what=f"{name} out c{c}")
        grads.append(onp.feat2emb_backward(params, lay, cache, g.upstream(step)[c]))
    tot = onp.accumulate(list(reversed(grads)))   # autograd runs the later calls' nodes first
    ref = g.group(f"s{step}/grad/")
    for k, v in ref.items():
        if v.size == 0:
            assert k not in tot or not np.any(tot[k]), k
            continue
        close(tot[k], v, what=f"{name} grad {k}")
        if k.split(".")[0] in ("item_emb", "user_emb", "sparse_emb"):
            assert not np.any(tot[k][0]), "padding row must get exactly zero grad"


@pytest.mark.parametrize("name", ["baseline_h32", "o1_h64"])
def test_numpy_adamw_matches_reference(name):
    g = Golden(name)
    params = g.params0()
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(x) for k, x in params.items()}
    for step in range(g.n_steps):
        ref_g = g.group(f"s{step}/grad/")
        ref_p = g.group(f"s{step}/param/")
        for k in params:
            params[k], m[k], v[k] = onp.adamw_dense(params[k], ref_g[k], m[k], v[k], step + 1, lr=g.lr, wd=g.wd)
            close(params[k], ref_p[k], rtol=2e-6, what=f"{name} step{step} {k}")
    last = g.n_steps - 1
    for k in params:
        close(m[k], g.z[f"s{last}/exp_avg/{k}"], rtol=2e-6, what=f"exp_avg {k}")
        close(v[k], g.z[f"s{last}/exp_avg_sq/{k}"], rtol=2e-6, what=f"exp_avg_sq {k}")


def test_lazy_rows_equal_dense_on_touched_rows_at_step1():
    """from zero state the sparse row update == the reference's dense AdamW on touched rows."""
    g = Golden("baseline_h32")
    params = g.params0()
    ref_g, ref_p = g.group("s0/grad/"), g.group("s0/param/")
    for k in params:
        if k.split(".")[0] not in ("item_emb", "user_emb", "sparse_emb"):
            continue
        rows = np.nonzero(np.any(ref_g[k] != 0, axis=1))[0]
        w, m, v = params[k].copy(), np.zeros_like(params[k]), np.zeros_like(params[k])
        onp.adamw_rows(w, m, v, rows, ref_g[k][rows], 1, lr=g.lr, wd=g.wd)
        close(w[rows], ref_p[k][rows], rtol=2e-6, what=k)
        untouched = np.setdiff1d(np.arange(w.shape[0]), rows)
        # dense AdamW scales untouched rows by exactly (1 - lr*wd); the lazy update leaves them alone
        close(params[k][untouched] * np.float32(1 - g.lr * g.wd), ref_p[k][untouched], rtol=2e-7, what=k + " untouched")
        assert not np.any(ref_p[k][0]), "row 0 stays zero under the reference flow"


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_torch_oracle_matches_reference(name):
    g = Golden(name)
    lay = g.layout
    orc = TorchOracle(lay)
    orc.load_numpy(g.params0())
    outs = []
    calls = g.calls(0)
    for c, pc in enumerate(calls):
        t = tensors_to_torch(onp.tensors_from_packed(lay, pc))
        seq = torch.from_numpy(pc.seq.astype(np.int64))
        mask = torch.from_numpy(pc.mask.astype(np.int64)) if pc.include_user else None
        outs.append(orc.feat2emb(seq, t, mask, pc.include_user))
        close(outs[-1].detach().numpy(), g.outs(0)[c], what=f"{name} torch out c{c}")
    loss = sum((o * torch.from_numpy(r)).sum() for o, r in zip(outs, g.upstream(0)))
    loss.backward()
    ref = g.group("s0/grad/")
    got = orc.grads_numpy()
    for k, v in ref.items():
        if v.size:
            close(got[k], v, what=f"{name} torch grad {k}")


def test_item_sweep_shape():
    """save_item_emb's [1, n] int64 call (model.py:418-425)."""
    g = Golden("item_sweep")
    pc = g.sweep_call()
    out, _ = onp.feat2emb_forward(g.params0(), g.layout, pc.seq, onp.tensors_from_packed(g.layout, pc), None, False)
    close(out, g.z["out"], what="item sweep")


def test_feat2tensor_round_trip():
    """dict form -> reference-style padded tensors == packed form re-padded (the tensorizer contract)."""
    g = Golden("baseline_h32")
    lay = g.layout
    for pc in g.calls(0):
        d = packed_to_dicts(lay, pc)
        a = onp.tensors_from_dicts(lay, d, pc.include_user)
        b = onp.tensors_from_packed(lay, pc)
        assert set(a) == set(b)
        for k in a:
            assert a[k].shape == b[k].shape, k
            assert np.array_equal(a[k], b[k]), k


def test_feat2tensor_ragged_raises():
    g = Golden("baseline_h32")
    pc = g.calls(0)[1]
    d = packed_to_dicts(g.layout, pc)
    d[0] = d[0][:-1]
    with pytest.raises(ValueError):
        onp.feat2tensor(d, "100", False)


def test_dedup_matches_torch_unique():
    g = Golden("baseline_h32")
    keys, srcs = onp.build_keys(g.layout, g.calls(0))
    order, uniq, seg, counts = onp.sort_dedup(keys)
    tu, tinv, tc = torch.unique(torch.from_numpy(keys.astype(np.int64)), sorted=True, return_inverse=True, return_counts=True)
    assert np.array_equal(uniq.astype(np.int64), tu.numpy())
    assert np.array_equal(counts, tc.numpy())
    assert np.array_equal(np.diff(seg), counts)
    # stable: sources ascending inside each segment
    ss = srcs[order].astype(np.int64)
    for u in range(len(uniq)):
        s = ss[seg[u]:seg[u + 1]]
        assert np.all(np.diff(s) >= 0)   # equal only for a repeated id inside one array


def test_segment_reduce_truth_matches_reference_grads():
    """fp64 per-unique-row gradient == rows of the reference's dense .grad (ties keys/sources to the reference)."""
    g = Golden("baseline_h32")
    lay = g.layout
    params = g.params0()
    calls = g.calls(0)
    d_cats = []
    for c, pc in enumerate(calls):
        _, cache = onp.feat2emb_forward(params, lay, pc.seq, onp.tensors_from_packed(lay, pc), pc.mask, pc.include_user)
        di, du, _ = onp.concat_backward(params, lay, cache, g.upstream(0)[c])
        d_cats.append((di, du))
    uniq, rows = onp.segment_reduce_fp64(lay, calls, d_cats)
    ref = g.group("s0/grad/")
    seen = 0
    for t in lay.tables:
        gk = ref[f"{t.name}.weight"]
        sel = (uniq >= t.key_base) & (uniq < t.key_base + t.rows)
        local = uniq[sel].astype(np.int64) - t.key_base
        close(rows[sel], gk[local], what=t.name)
        touched = np.nonzero(np.any(gk != 0, axis=1))[0]
        assert set(touched) <= set(local)
        seen += sel.sum()
    assert seen == uniq.size


def test_route_w1_identity_and_partition():
    keys = np.array([5, 9, 12, 12, 40, 41, 77], np.uint32)
    for W in (1, 2, 4, 8):
        owner, local, counts, order = onp.route(keys, W)
        assert counts.sum() == keys.size
        assert np.array_equal(owner.astype(np.int64) + local * W, keys.astype(np.int64))
        assert np.all(np.diff(owner[order]) >= 0)


def test_fbin_writer_matches_reference_bytes(tmp_path):
    """binfmt.save_emb == the reference's save_emb byte for byte (fixtures written by the unmodified reference,
    tests/golden/make_golden_fbin.py), and the reader round-trips both files."""
    import os
    from golden_util import GOLDEN_DIR
    from tencent_recommendation_2025_b200 import binfmt
    z = np.load(os.path.join(GOLDEN_DIR, "fbin_inputs.npz"))
    for arr, name, dt in ((z["emb"], "fbin_embedding.bin", np.float32), (z["ids"], "fbin_ids.bin", np.uint64)):
        out = tmp_path / name
        binfmt.save_emb(arr, out)
        ref = open(os.path.join(GOLDEN_DIR, name), "rb").read()
        assert out.read_bytes() == ref
        back = binfmt.load_emb(out, dt)
        assert back.dtype == dt and np.array_equal(back, arr)
    binfmt.save_emb(torch.from_numpy(z["emb"]), tmp_path / "t.fbin")
    assert (tmp_path / "t.fbin").read_bytes() == open(os.path.join(GOLDEN_DIR, "fbin_embedding.bin"), "rb").read()
    assert np.array_equal(binfmt.read_result_ids(os.path.join(GOLDEN_DIR, "fbin_ids.bin")), z["ids"])
