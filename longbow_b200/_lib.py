"""ctypes binding of liblongbow_b200.so (include/longbow_b200.h).

The library is the product; this module only declares signatures.  There is no CPU
fallback: if the shared object is missing, ``load()`` raises, and if there is no CUDA
device every call returns ``LB_ERR_NO_DEVICE`` which ``check()`` turns into an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblongbow_b200.so")

LB_OK, LB_ERR_INVALID, LB_ERR_CUDA, LB_ERR_OOM, LB_ERR_UNSUPPORTED, LB_ERR_STATE, LB_ERR_NO_DEVICE = range(7)
ERR_NAMES = {1: "LB_ERR_INVALID", 2: "LB_ERR_CUDA", 3: "LB_ERR_OOM", 4: "LB_ERR_UNSUPPORTED",
             5: "LB_ERR_STATE", 6: "LB_ERR_NO_DEVICE"}

METRIC_L2, METRIC_COSINE, METRIC_DOT = 0, 1, 2
F32, F16, I8, U8 = 0, 1, 2, 3

vp, i32, i64, sz, fp = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float

# name -> (restype, argtypes); every function include/longbow_b200.h declares
SIGNATURES = {
    "lb_last_error": (C.c_char_p, []),
    "lb_device_info": (i32, [i32, C.POINTER(i32), C.POINTER(sz), C.c_char_p, sz]),
    "faiss_gpu_resources_new": (vp, [i32]),
    "faiss_gpu_resources_free": (None, [vp]),
    "faiss_gpu_index_flat_l2_new": (vp, [vp, i32]),
    "faiss_gpu_index_flat_l2_free": (None, [vp]),
    "faiss_gpu_index_add": (i32, [vp, i64, vp]),
    "faiss_gpu_index_search": (i32, [vp, i64, vp, i32, vp, vp]),
    "lb_index_create": (i32, [i32, i32, i32, i32, C.POINTER(vp)]),
    "lb_index_free": (None, [vp]),
    "lb_index_reserve": (i32, [vp, i64]),
    "lb_index_add": (i32, [vp, vp, i64]),
    "lb_index_add_device": (i32, [vp, vp, i64, vp]),
    "lb_index_add_arrow": (i32, [vp, vp, sz, i64, i64, i32]),
    "lb_index_size": (i64, [vp]),
    "lb_index_dim": (i32, [vp]),
    "lb_index_set_id_base": (i32, [vp, i64]),
    "lb_index_set_tombstones": (i32, [vp, vp, i64]),
    "lb_index_set_tombstones_device": (i32, [vp, vp, i64, vp]),
    "lb_index_search": (i32, [vp, vp, i64, i32, vp, vp, vp]),
    "lb_index_search_device": (i32, [vp, vp, i64, i32, vp, vp, vp, vp]),
    "lb_index_search_device_cert": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, vp, vp]),
    "lb_index_search_exact_device": (i32, [vp, vp, i64, i32, vp, vp, vp, vp, vp]),
    "lb_index_rerank": (i32, [vp, vp, i64, vp, i32, i32, vp, vp, vp]),
    "lb_index_rerank_device": (i32, [vp, vp, i64, vp, i32, i32, vp, vp, vp, vp]),
    "lb_index_distances": (i32, [vp, vp, vp]),
    "lb_index_last_uncertified": (i64, [vp]),
    "lb_index_coarse_keys": (i32, [vp, vp, i64, i64, vp]),
    "lb_simd_distance_batch_flat": (i32, [i32, i32, i32, vp, vp, i64, i32, vp]),
    "lb_simd_adc_distance_batch": (i32, [i32, vp, vp, i32, i64, vp]),
    "lb_simd_quantize_sq8": (i32, [i32, vp, i64, fp, fp, vp]),
    "lb_simd_dequantize_sq8": (i32, [i32, vp, i64, fp, fp, vp]),
    "lb_simd_compute_bounds": (i32, [i32, vp, i64, C.POINTER(fp), C.POINTER(fp)]),
    "lb_simd_sq8_dequant_distance_batch": (i32, [i32, vp, vp, i64, i32, fp, fp, vp]),
    "lb_simd_find_nearest_centroid": (i32, [i32, vp, vp, i32, i32, C.POINTER(i32), C.POINTER(fp)]),
    "lb_filter_scatter_device": (i32, [i32, vp, i64, vp, C.c_uint32, i64, vp, vp]),
    "lb_select_k": (i32, [i32, vp, i64, i32, vp, vp]),
    "lb_merge_topk": (i32, [i32, vp, vp, i32, i64, i32, i32, vp, vp]),
    "lb_merge_topk_device": (i32, [i32, vp, vp, i32, i64, i32, i32, vp, vp, vp]),
    "lb_merge_topk_packed_device": (i32, [i32, vp, sz, sz, i32, i64, i32, i32, vp, vp, vp]),
    "lb_exchange_create": (i32, [i32, i32, i32, sz, C.POINTER(vp)]),
    "lb_exchange_free": (None, [vp]),
    "lb_exchange_handle": (i32, [vp, vp]),
    "lb_exchange_connect_ipc": (i32, [vp, vp]),
    "lb_exchange_connect_local": (i32, [vp, C.POINTER(vp)]),
    "lb_exchange_slot": (i32, [vp, i64, i32, C.POINTER(vp), C.POINTER(vp)]),
    "lb_exchange_all_gather_merge": (i32, [vp, i64, i32, i32, vp, vp, vp]),
    "lb_exchange_error": (i32, [vp]),
    "lb_shard_create": (i32, [C.POINTER(i32), i32, i32, i32, i32, i64, C.POINTER(vp)]),
    "lb_shard_free": (None, [vp]),
    "lb_shard_add": (i32, [vp, vp, i64]),
    "lb_shard_size": (i64, [vp]),
    "lb_shard_count": (i32, [vp]),
    "lb_shard_rows_per_shard": (i64, [vp]),
    "lb_shard_set_tombstones": (i32, [vp, vp, i64]),
    "lb_shard_search": (i32, [vp, vp, i64, i32, vp, vp, vp]),
    "lb_shard_last_uncertified": (i64, [vp]),
    "lb_pq_create": (i32, [i32, vp, sz, C.POINTER(vp)]),
    "lb_pq_free": (None, [vp]),
    "lb_pq_params": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "lb_pq_add_codes": (i32, [vp, vp, i64]),
    "lb_pq_add_codes_device": (i32, [vp, vp, i64, vp]),
    "lb_pq_size": (i64, [vp]),
    "lb_pq_attach_raw": (i32, [vp, vp]),
    "lb_pq_set_tombstones": (i32, [vp, vp, i64]),
    "lb_pq_build_adc_table": (i32, [vp, vp, vp]),
    "lb_pq_encode": (i32, [vp, vp, i64, vp]),
    "lb_pq_train": (i32, [i32, vp, i64, i32, i32, i32, i32, vp, vp, vp]),
    "lb_pq_search": (i32, [vp, vp, i64, i32, i32, vp, vp, vp]),
    "lb_pq_search_device": (i32, [vp, vp, i64, i32, i32, vp, vp, vp, vp]),
    "lb_pq_last_uncertified": (i64, [vp]),
    "lb_pq_search_device_cert": (i32, [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
    "lb_graph_create": (i32, [vp, i32, C.POINTER(vp)]),
    "lb_graph_free": (None, [vp]),
    "lb_graph_set_layer": (i32, [vp, vp, vp, i64]),
    "lb_graph_set_layer_device": (i32, [vp, vp, vp, i64, vp]),
    "lb_graph_search_layer": (i32, [vp, vp, i64, vp, i32, vp, vp, vp]),
    "lb_graph_search": (i32, [vp, vp, i64, vp, i32, i32, vp, vp, vp]),
    "lb_graph_search_layer_device": (i32, [vp, vp, i64, vp, i32, vp, vp, vp, vp, vp]),
    "lb_graph_search_device": (i32, [vp, vp, i64, vp, i32, i32, vp, vp, vp, vp, vp]),
    "lb_filter_i64": (i32, [i32, vp, i64, i32, i64, i32, vp]),
    "lb_filter_f32": (i32, [i32, vp, i64, i32, fp, i32, vp]),
    "lb_filter_i64_device": (i32, [i32, vp, i64, i32, i64, i32, vp, vp]),
    "lb_filter_f32_device": (i32, [i32, vp, i64, i32, fp, i32, vp, vp]),
    "lb_kernel_launch_count": (i64, []),
    "lb_set_option": (i32, [C.c_char_p, i32]),
    "lb_prof_enable": (i32, [i32]),
    "lb_prof_read": (i32, [C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double), i32]),
    "lb_prof_read_aux": (i32, [C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double), i32]),
}

_lib = None


class LongbowError(RuntimeError):
    def __init__(self, code: int, text: str):
        self.code = code
        super().__init__(f"{ERR_NAMES.get(code, code)}: {text}")


def load() -> C.CDLL:
    """Load the CUDA library.  Fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C longbow_b200/csrc`.  longbow_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise LongbowError(rc, load().lb_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().lb_kernel_launch_count())


def set_option(name: str, value: int) -> None:
    check(load().lb_set_option(name.encode(), int(value)))


def prof_enable(on: bool) -> None:
    load().lb_prof_enable(int(on))


def prof_read(reset: bool = True):
    """(total_ms, launches, query-row pairs scanned) of the dominant scan kernel since the last reset."""
    ms, n, u = C.c_double(), i64(), C.c_double()
    load().lb_prof_read(C.byref(ms), C.byref(n), C.byref(u), int(reset))
    return ms.value, n.value, u.value


def prof_read_aux(reset: bool = True):
    """(total_ms, launches, bytes moved) of the auxiliary streaming kernels (the batched PQ path's decode)."""
    ms, n, u = C.c_double(), i64(), C.c_double()
    load().lb_prof_read_aux(C.byref(ms), C.byref(n), C.byref(u), int(reset))
    return ms.value, n.value, u.value
