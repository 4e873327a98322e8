"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/tgr_embed.h declares
(no compute calls without a GPU); host-side packing logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from golden_util import Golden
from oracle import feat2emb_numpy as onp
from tencent_recommendation_2025_b200 import _lib, build
from tencent_recommendation_2025_b200.packed import count_valid, pack_from_dicts, to_device
from tencent_recommendation_2025_b200.synth import packed_to_dicts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tgr_embed.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tgr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in tgr_embed.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.tgr_abi_version() == _lib.TGR_ABI_VERSION


def test_struct_layouts_match_header(lib):
    # sizes the C compiler produces for the header's structs (LP64)
    assert ctypes.sizeof(_lib.Table) == 48
    assert ctypes.sizeof(_lib.Slot) == 20
    assert ctypes.sizeof(_lib.Adam) == 48
    assert _lib.Call.ids.offset == 16 + 20 * 32
    assert ctypes.sizeof(_lib.Call) == 16 + 640 + 8 + 64 + 64 + 32 + 32 + 8 + 8 + 8 + 8 + 8 + 8 + 8


def test_ctypes_structs_match_the_c_compiler(lib, tmp_path):
    """sizeof / key offsets of every struct in include/tgr_embed.h as gcc lays them out == the ctypes mirrors."""
    import subprocess
    names = {"tgr_table_t": _lib.Table, "tgr_slot_t": _lib.Slot, "tgr_call_t": _lib.Call, "tgr_adam_t": _lib.Adam,
             "tgr_dnn_t": _lib.Dnn, "tgr_mm_feat_t": _lib.MmFeat, "tgr_fact_params_t": _lib.FactParams,
             "tgr_fact_grads_t": _lib.FactGrads, "tgr_fact_group_t": _lib.FactGroup,
             "tgr_row_source_t": _lib.RowSource}
    offs = [("tgr_fact_group_t", "calls"), ("tgr_fact_group_t", "mm_x"), ("tgr_fact_group_t", "src"), ("tgr_fact_group_t", "cap"), ("tgr_fact_group_t", "rows_local"),
            ("tgr_row_source_t", "save_rows"),
            ("tgr_fact_group_t", "P"), ("tgr_fact_group_t", "mmz"), ("tgr_fact_group_t", "ws_bytes"),
            ("tgr_fact_group_t", "n_backward"), ("tgr_fact_params_t", "b_item"), ("tgr_fact_params_t", "n_mm"),
            ("tgr_call_t", "err_flag"), ("tgr_dnn_t", "table_col")]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "tgr_embed.h"', 'int main(void) {']
    src += [f'printf("%zu\\n", sizeof({n}));' for n in names]
    src += [f'printf("%zu\\n", offsetof({n}, {f}));' for n, f in offs]
    src += ['return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(t) for t in names.values()] + [getattr(names[n], f).offset for n, f in offs]
    assert got == want, list(zip(list(names) + offs, got, want))


def test_size_queries_work_without_gpu(lib):
    assert lib.tgr_build_keys_workspace_bytes(1 << 20) > 0
    assert lib.tgr_dedup_workspace_bytes(1 << 20) > 0
    assert lib.tgr_reduce_workspace_bytes(1 << 20, 64) > 0
    assert lib.tgr_route_workspace_bytes(1 << 20, 8) > 0
    assert lib.tgr_mm_proj_bwd_workspace_bytes(103424, 32, 64) > 0


def test_argument_errors_return_negative_and_message(lib):
    rc = lib.tgr_fwd_gather_pool_concat(None, 0, 64, None, None)
    assert rc < 0 and b"null" in lib.tgr_last_error()
    rc = lib.tgr_sort_pairs(None, None, None, None, 10, 99, None, 0, None)
    assert rc < 0 and b"key_bits" in lib.tgr_last_error()


@pytest.mark.parametrize("name", ["baseline_h32", "o1_h64_mm2"])
def test_pack_from_dicts_round_trip(name):
    """The one-pass tensorizer reproduces the packed form (and hence feat2tensor's tensors) exactly."""
    g = Golden(name)
    lay = g.layout
    for pc in g.calls(0):
        d = packed_to_dicts(lay, pc)
        got = pack_from_dicts(lay, torch.from_numpy(pc.seq), d, None if pc.mask is None else torch.from_numpy(pc.mask),
                              pc.include_user)
        assert np.array_equal(got.ids, pc.ids)
        assert np.array_equal(got.arr_off, pc.arr_off)
        assert np.array_equal(got.arr_val, pc.arr_val)
        for a, b in zip(got.mm_x, pc.mm_x):
            assert np.array_equal(a, b)
        a = onp.tensors_from_dicts(lay, d, pc.include_user)
        b = onp.tensors_from_packed(lay, got)
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_pack_missing_mm_key_gives_zeros_and_ragged_raises():
    g = Golden("baseline_h32")
    lay = g.layout
    pc = g.calls(0)[1]
    d = packed_to_dicts(lay, pc)
    del d[0][3]["81"]
    got = pack_from_dicts(lay, pc.seq, d, None, False)
    assert not got.mm_x[0][3].any()
    d[1] = d[1][:-2]
    with pytest.raises(ValueError):
        pack_from_dicts(lay, pc.seq, d, None, False)


def test_to_device_cpu_layout_and_counts():
    g = Golden("baseline_h32")
    lay = g.layout
    pc = g.calls(0)[0]
    pb = to_device(lay, pc, "cpu", pin=False)
    assert pb.ids.dtype == torch.int32 and pb.ids.is_contiguous() and pb.ids.data_ptr() % 16 == 0
    assert torch.equal(pb.ids, torch.from_numpy(pc.ids))
    assert pb.arr_tok.numel() == pc.arr_val.size
    keys, _ = onp.build_keys(lay, [pc])
    assert pb.n_valid == keys.size == count_valid(lay, pc)
    # COO tokens agree with the CSR offsets
    for j in range(pc.arr_off.shape[0]):
        toks = pb.arr_tok[pb.arr_begin[j]:pb.arr_begin[j] + pb.arr_nnz[j]].numpy()
        for i, t in enumerate(toks):
            assert pc.arr_off[j, t] <= pb.arr_begin[j] + i < pc.arr_off[j, t + 1]


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtgr_embed.so")
    with pytest.raises(_lib.TgrError):
        _lib.load()


def _same_packed(a, b):
    assert np.array_equal(a.ids, b.ids) and a.ids.dtype == b.ids.dtype
    assert np.array_equal(a.arr_off, b.arr_off) and np.array_equal(a.arr_val, b.arr_val)
    assert len(a.mm_x) == len(b.mm_x)
    for x, y in zip(a.mm_x, b.mm_x):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    assert np.array_equal(a.seq, b.seq)
    assert (a.mask is None) == (b.mask is None) and (a.mask is None or np.array_equal(a.mask, b.mask))


@pytest.mark.parametrize("name", ["baseline_h32", "o1_h64_mm2", "baseline_l102_nomm"])
def test_native_tensorizer_equals_python_restatement(name):
    """csrc/tgr_pack.c (one dict walk in C) == packed.pack_from_dicts_py on every fixture call, and on the input
    variants the reference's pipeline produces: object-array rows (np.array(dicts)), numpy integer ids, tuple / ndarray
    id lists, list-typed mm vectors, mapping (non-dict) tokens."""
    import collections
    from tencent_recommendation_2025_b200.packed import pack_from_dicts_py
    g = Golden(name)
    lay = g.layout
    for pc in g.calls(0):
        d = packed_to_dicts(lay, pc)
        seq = torch.from_numpy(pc.seq)
        mask = None if pc.mask is None else torch.from_numpy(pc.mask)
        ref = pack_from_dicts_py(lay, seq, d, mask, pc.include_user)
        _same_packed(pack_from_dicts(lay, seq, d, mask, pc.include_user), ref)
        # variants
        v = []
        for row in d:
            new = []
            for t, tok in enumerate(row):
                tok2 = {}
                for k, val in tok.items():
                    if isinstance(val, (int, np.integer)):
                        tok2[k] = np.int64(val) if t % 2 else int(val)
                    elif isinstance(val, np.ndarray) and val.dtype == np.float32:
                        tok2[k] = val.tolist() if t % 3 == 0 else val
                    elif isinstance(val, (list, tuple, np.ndarray)):
                        tok2[k] = tuple(val) if t % 2 else np.asarray(val, np.int64)
                    else:
                        tok2[k] = val
                new.append(collections.ChainMap(tok2) if t % 5 == 4 else tok2)
            v.append(np.array(new, dtype=object))
        _same_packed(pack_from_dicts(lay, seq, v, mask, pc.include_user), ref)


def test_native_tensorizer_errors():
    g = Golden("baseline_h32")
    lay = g.layout
    pc = g.calls(0)[0]
    d = packed_to_dicts(lay, pc)
    seq, mask = torch.from_numpy(pc.seq), torch.from_numpy(pc.mask)
    bad = [list(r) for r in d]
    bad[0][1] = {k: v for k, v in bad[0][1].items() if k != "100"}
    with pytest.raises(KeyError):
        pack_from_dicts(lay, seq, bad, mask, True)                     # a dict lacking a sparse key (model.py:222)
    bad = [list(r) for r in d]
    bad[0][1] = dict(bad[0][1], **{"81": np.zeros(7, np.float32)})
    with pytest.raises(ValueError):
        pack_from_dicts(lay, seq, bad, mask, True)                     # mm vector of the wrong width
    bad = [list(r) for r in d]
    bad[0][1] = dict(bad[0][1], **{"100": 1 << 40})
    with pytest.raises(OverflowError):
        pack_from_dicts(lay, seq, bad, mask, True)
    with pytest.raises(ValueError):
        pack_from_dicts(lay, seq, d, None, True)                       # include_user without the mask
