"""ctypes binding of libtgr_embed.so (include/tgr_embed.h). Fails loudly when the library is missing —
there is no CPU or PyTorch fallback for the product path."""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libtgr_embed.so")

TGR_ABI_VERSION = 1
MAX_TABLES, MAX_SLOTS, MAX_ARRAYS, MAX_CALLS = 64, 32, 8, 4
MAX_PEERS = 16
KIND_SINGLE, KIND_ARRAY, KIND_MM = 0, 1, 2
DTYPE_F32, DTYPE_BF16 = 0, 1

i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)


class Table(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("grad", C.c_void_p),
                ("rows", C.c_int64), ("key_base", C.c_int64)]


class Slot(C.Structure):
    _fields_ = [("kind", C.c_int32), ("side", C.c_int32), ("col", C.c_int32), ("table", C.c_int32), ("src", C.c_int32)]


class Call(C.Structure):
    _fields_ = [("T", C.c_int32), ("n_slots", C.c_int32), ("n_single", C.c_int32), ("n_arrays", C.c_int32),
                ("slots", Slot * MAX_SLOTS),
                ("ids", C.c_void_p),
                ("arr_off", C.c_void_p * MAX_ARRAYS), ("arr_tok", C.c_void_p * MAX_ARRAYS),
                ("arr_begin", C.c_int32 * MAX_ARRAYS), ("arr_nnz", C.c_int32 * MAX_ARRAYS),
                ("arr_val", C.c_void_p),
                ("item_cat", C.c_void_p), ("user_cat", C.c_void_p),
                ("item_ld", C.c_int64), ("user_ld", C.c_int64),
                ("cat_dtype", C.c_int32), ("reserved", C.c_int32),
                ("err_flag", C.c_void_p)]


class Dnn(C.Structure):
    _fields_ = [("w_item", C.c_void_p), ("w_user", C.c_void_p), ("item_ld", C.c_int64), ("user_ld", C.c_int64),
                ("table_side", C.c_int32 * MAX_TABLES), ("table_col", C.c_int32 * MAX_TABLES)]


MAX_MM = 6


class RowSource(C.Structure):
    _fields_ = [("fetched_rows", C.c_void_p), ("fetched_perm", C.c_void_p), ("peer_rows", C.c_void_p * MAX_PEERS),
                ("n_peers", C.c_int32), ("reserved", C.c_int32), ("save_rows", C.c_void_p)]


class MmFeat(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("mm_dim", C.c_int32), ("col", C.c_int32)]


class FactParams(C.Structure):
    _fields_ = [("dnn", Dnn), ("b_item", C.c_void_p), ("b_user", C.c_void_p), ("mm", MmFeat * MAX_MM),
                ("n_mm", C.c_int32), ("reserved", C.c_int32)]


class FactGrads(C.Structure):
    _fields_ = [("dW_item", C.c_void_p), ("db_item", C.c_void_p), ("dW_user", C.c_void_p), ("db_user", C.c_void_p),
                ("dW_mm", C.c_void_p * MAX_MM), ("db_mm", C.c_void_p * MAX_MM)]


class FactGroup(C.Structure):
    _fields_ = [("n_calls", C.c_int32), ("H", C.c_int32), ("key_bits", C.c_int32), ("n_mm", C.c_int32),
                ("n", C.c_int64), ("mm_dim", C.c_int32 * MAX_MM), ("mm_x_dtype", C.c_int32), ("n_is_capacity", C.c_int32),
                ("calls", Call * MAX_CALLS), ("mm_x", (C.c_void_p * MAX_MM) * MAX_CALLS),
                ("src", RowSource),
                ("cap", C.c_int64),
                ("keys_in", C.c_void_p), ("srcs_in", C.c_void_p), ("keys", C.c_void_p), ("srcs", C.c_void_p),
                ("uniq", C.c_void_p),
                ("seg_off", C.c_void_p), ("seg_of", C.c_void_p), ("n_unique", C.c_void_p), ("n_valid", C.c_void_p),
                ("P", C.c_void_p), ("rows_local", C.c_void_p), ("G", C.c_void_p),
                ("ids_u", C.c_void_p * MAX_CALLS), ("arr_u", C.c_void_p * MAX_CALLS), ("mask", C.c_void_p * MAX_CALLS),
                ("dz_item", C.c_void_p * MAX_CALLS), ("dz_user", C.c_void_p * MAX_CALLS),
                ("mmz", (C.c_void_p * MAX_MM) * MAX_CALLS),
                ("fold_M", C.c_void_p * MAX_MM), ("fold_c", C.c_void_p * MAX_MM), ("mm_A", C.c_void_p * MAX_MM),
                ("mm_s", C.c_void_p * MAX_MM), ("fold_Mb", C.c_void_p * MAX_MM), ("dzb", C.c_void_p),
                ("ws", C.c_void_p), ("ws_bytes", C.c_size_t),
                ("projected", C.c_int32), ("n_backward", C.c_int32), ("mm_done", C.c_int32), ("mm_joined", C.c_int32),
                ("reduce_done_event", C.c_void_p)]


MAX_DENSE = 16


class DenseList(C.Structure):
    _fields_ = [("w", C.c_void_p * MAX_DENSE), ("g", C.c_void_p * MAX_DENSE), ("m", C.c_void_p * MAX_DENSE),
                ("v", C.c_void_p * MAX_DENSE), ("numel", C.c_int64 * MAX_DENSE), ("n", C.c_int32), ("reserved", C.c_int32)]


class Adam(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step_size", C.c_float), ("bc2_sqrt", C.c_float),
                ("grad_scale", C.c_float), ("decay", C.c_float), ("one_minus_beta1", C.c_float),
                ("one_minus_beta2", C.c_float), ("reserved", C.c_float)]


def make_adam(lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, step: int,
              grad_scale: float = 1.0) -> "Adam":
    """Scalars exactly as torch/optim/adam.py forms them (python doubles, rounded once to float)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    return Adam(lr, beta1, beta2, eps, weight_decay, lr / bc1, bc2 ** 0.5, grad_scale, 1 - lr * weight_decay,
                1 - beta1, 1 - beta2, 0.0)


# name -> (restype, argtypes); every symbol include/tgr_embed.h declares
SIGNATURES = {
    "tgr_abi_version": (C.c_int, []),
    "tgr_last_error": (C.c_char_p, []),
    "tgr_launch_count": (C.c_int64, []),
    "tgr_timing_enable": (C.c_int, [C.c_int]),
    "tgr_timing_collect": (C.c_int, [C.c_char_p, C.c_size_t, f32p, i32p, C.c_int]),
    "tgr_fwd_gather_pool_concat": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.POINTER(Call), C.c_void_p]),
    "tgr_mm_proj_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_int64, C.c_int, C.c_void_p]),
    "tgr_mm_proj_fwd_tc_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "tgr_mm_proj_fwd_tc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_int64, C.c_int, C.c_void_p]),
    "tgr_split_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_cast_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "tgr_mm_proj_bwd_tc_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "tgr_mm_proj_bwd_tc_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "tgr_mm_proj_bwd_tc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_mm_proj_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "tgr_mm_proj_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_bwd_max_entries": (C.c_int64, [C.POINTER(Call), C.c_int]),
    "tgr_build_keys_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "tgr_bwd_build_keys": (C.c_int, [C.POINTER(Table), C.c_int, C.POINTER(Call), C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_sort_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "tgr_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "tgr_dedup_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "tgr_dedup": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_size_t, C.c_void_p]),
    "tgr_reduce_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "tgr_bwd_reduce": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.POINTER(Call), C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Adam), C.c_void_p, C.c_size_t,
                                 C.c_void_p]),
    "tgr_bwd_reduce_rows": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.POINTER(Adam), C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_adam_rows": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.POINTER(Adam), C.c_void_p]),
    "tgr_adam_dense": (C.c_int, [C.POINTER(DenseList), C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_adam_rows_dev": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
    "tgr_scatter_rows": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p]),
    "tgr_route_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "tgr_route_bucket": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_remap_ids": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, u32p, i32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "tgr_remap_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(Call), C.c_int,
                                    C.POINTER(C.c_void_p), C.c_void_p]),
    "tgr_fetch_peer_rows": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p]),
    "tgr_expand_item_features": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64,
                                           C.c_void_p, C.c_void_p]),
    "tgr_scatter_user_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tgr_gather_mm_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "tgr_peer_put": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "tgr_peer_pull": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_void_p]),
    "tgr_merge_buckets": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_peer_barrier": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]),
    "tgr_allreduce_peers": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "tgr_remap_arrays": (C.c_int, [C.POINTER(Table), C.c_int, C.POINTER(Call), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_void_p), C.c_void_p]),
    "tgr_permute_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "tgr_gather_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "tgr_fact_project_rows": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.POINTER(Dnn), C.c_void_p, C.c_void_p,
                                        C.c_int64, C.POINTER(RowSource), C.c_void_p, C.c_void_p]),
    "tgr_fact_forward": (C.c_int, [C.POINTER(Call), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tgr_fact_relu_mask_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "tgr_fact_relu_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    "tgr_fact_backward_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "tgr_fact_unique_backward": (C.c_int, [C.POINTER(Table), C.c_int, C.c_int, C.POINTER(Dnn), C.c_void_p, C.c_void_p,
                                           C.c_int64, C.POINTER(RowSource), C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_fact_mm_fold": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "tgr_fact_group_bytes": (C.c_size_t, [C.POINTER(FactGroup), C.c_int]),
    "tgr_fact_prepare": (C.c_int, [C.POINTER(Table), C.c_int, C.POINTER(FactGroup), C.c_void_p, C.c_size_t, C.c_void_p]),
    "tgr_fact_mm_branch": (C.c_int, [C.POINTER(FactParams), C.POINTER(FactGroup), C.c_int, C.c_void_p]),
    "tgr_fact_call_forward": (C.c_int, [C.POINTER(Table), C.c_int, C.POINTER(FactParams), C.POINTER(FactGroup), C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "tgr_fact_call_backward": (C.c_int, [C.POINTER(Table), C.c_int, C.POINTER(FactParams), C.POINTER(FactGroup), C.c_int,
                                         C.c_void_p, C.POINTER(FactGrads), C.c_int, C.c_void_p]),
    "tgr_fact_mm_chain_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


class TgrError(RuntimeError):
    pass


def load():
    """Load (once) and type the C-ABI library. Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise TgrError(f"{LIB_PATH} not found: build it with `python -m tencent_recommendation_2025_b200.build` "
                           "(there is no CPU fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.tgr_abi_version() != TGR_ABI_VERSION:
            raise TgrError(f"ABI mismatch: library {lib.tgr_abi_version()} vs binding {TGR_ABI_VERSION}")
        _lib = lib
        return lib


def launch_count() -> int:
    """Kernels launched by libtgr_embed.so in this process so far (counted at every launch site)."""
    return int(load().tgr_launch_count())


def timing_enable(on: bool = True):
    """Per-entry CUDA-event timing inside the library (bench / profiling)."""
    load().tgr_timing_enable(1 if on else 0)


def timing_collect():
    """{entry name: (total ms, calls)} since timing_enable(True); clears the records."""
    lib = load()
    names = C.create_string_buffer(4096)
    ms = (C.c_float * 64)()
    cnt = (C.c_int32 * 64)()
    n = lib.tgr_timing_collect(names, 4096, ms, cnt, 64)
    keys = names.value.decode().split("\n")[:n]
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(keys)}


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().tgr_last_error().decode("utf-8", "replace")
        raise TgrError(f"{what or 'tgr call'} failed ({rc}): {msg}")
