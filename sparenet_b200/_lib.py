"""ctypes binding of libsparenet_b200.so (the C ABI declared in include/sparenet_b200.h).

There is no CPU fallback and no PyTorch re-implementation behind these calls: if the shared library is
missing the import fails loudly, and every op raises on a non-zero return code.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsparenet_b200.so")

c_int, c_float, c_double, c_size_t, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t, ctypes.c_void_p
P = c_void_p



class GemmDesc(ctypes.Structure):
    """snb_gemm_desc of include/sparenet_b200.h (field order and types must match)."""
    _fields_ = [("mode", c_int), ("G", c_int), ("BI", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
                ("A", c_void_p), ("lda", ctypes.c_longlong), ("a_batch_stride", ctypes.c_longlong),
                ("B", c_void_p), ("ldb", ctypes.c_longlong), ("b_batch_stride", ctypes.c_longlong),
                ("D", c_void_p), ("ldd", ctypes.c_longlong), ("d_batch_stride", ctypes.c_longlong),
                ("block_n", c_int), ("store", c_int), ("split", c_int),
                ("scale", c_void_p), ("shift", c_void_p), ("slope", c_float), ("seg", c_int),
                ("pmean", c_void_p), ("pm2", c_void_p), ("pmax", c_void_p), ("pmin", c_void_p), ("pimax", c_void_p), ("pimin", c_void_p),
                ("b_pos_mod", c_int)]


# name -> (restype, argtypes); must list every symbol of include/sparenet_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "snb_version": (c_int, []),
    "snb_strerror": (ctypes.c_char_p, [c_int]),
    "snb_chamfer_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "snb_chamfer_fwd": (c_int, [P, P, c_int, c_int, c_int, P, P, P, P, P, c_size_t, P]),
    "snb_chamfer_bwd": (c_int, [P, P, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "snb_emd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "snb_emd_fwd": (c_int, [P, P, c_int, c_int, c_float, c_int, P, P, P, c_size_t, P]),
    "snb_emd_fwd_scan": (c_int, [P, P, c_int, c_int, c_float, c_int, P, P, P, c_size_t, P]),
    "snb_emd_bwd": (c_int, [P, P, c_int, c_int, P, P, P, P]),
    "snb_expansion_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "snb_expansion_fwd": (c_int, [P, c_int, c_int, c_int, c_float, P, P, P, P, c_size_t, P]),
    "snb_expansion_bwd": (c_int, [P, c_int, c_int, P, P, P, P]),
    "snb_mds_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "snb_mds_sample": (c_int, [P, c_int, c_int, c_int, P, P, P, c_size_t, P]),
    "snb_gather_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "snb_gather_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "snb_p2i_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "snb_p2i_max_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int, P, P, P, c_size_t, P]),
    "snb_p2i_max_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int, P, P, P, P]),
    "snb_p2i_sum_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int, P, P]),
    "snb_p2i_sum_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int, P, P, P]),
    "snb_depthmaps_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "snb_depthmaps_fwd": (c_int, [P, c_int, c_int, P, c_int, c_int, c_double, P, P, P, c_size_t, P]),
    "snb_depthmaps_bwd": (c_int, [P, P, P, c_int, c_int, P, c_int, c_int, c_double, P, c_size_t, P, P]),
    "snb_knn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "snb_knn": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_size_t, P]),
    "snb_knn_pruned_workspace_bytes": (c_size_t, [c_int, c_int]),
    "snb_knn_pruned": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, c_size_t, P]),
    "snb_transpose_cn": (c_int, [P, c_int, c_int, c_int, P, P]),
    "snb_edge_reduce_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "snb_edge_reduce_bwd": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "snb_edge_reduce_sel_fwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "snb_edge_reduce_sel_bwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "snb_edge_reduce_sel_fwd_stacked": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "snb_edge_reduce_sel_bwd_stacked": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P]),
    "snb_row_stats": (c_int, [P, ctypes.c_longlong, c_int, P, P, P]),
    "snb_row_stats_bwd": (c_int, [P, P, P, P, ctypes.c_longlong, c_int, P, P]),
    "snb_row_affine_act_fwd": (c_int, [P, P, P, ctypes.c_longlong, c_int, c_int, c_float, P, P]),
    "snb_row_affine_act_bwd": (c_int, [P, P, P, P, ctypes.c_longlong, c_int, c_int, c_float, P, P, P, P]),
    "snb_row_minmax": (c_int, [P, ctypes.c_longlong, c_int, P, P, P, P, P]),
    "snb_row_stats_minmax": (c_int, [P, ctypes.c_longlong, c_int, P, P, P, P, P, P, P]),
    "snb_row_act_bwd_reduce": (c_int, [P, P, P, P, ctypes.c_longlong, c_int, c_float, P, P, P, P]),
    "snb_row_norm_act_bwd": (c_int, [P, P, P, P, P, P, P, ctypes.c_longlong, c_int, c_float, P, P, P]),
    "snb_gemm_tf32": (c_int, [ctypes.POINTER(GemmDesc), P]),
    "snb_gemm_tf32_tiles": (c_int, [c_int, c_int]),
    "snb_conv_extrema_bwd": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "snb_gemm_stats_merge": (c_int, [P, P, ctypes.c_longlong, c_int, c_int, P, P, P]),
    "snb_gemm_minmax_merge": (c_int, [P, P, P, P, ctypes.c_longlong, c_int, P, P, P, P, P]),
    "snb_gemm_tf32_block_n": (c_int, [c_int, c_int]),
    "snb_bn_max_tail_fwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_float, c_int, c_float, c_float, P, P, P, P, P, P]),
    "snb_bn_max_tail_bwd": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P, P, P]),
    "snb_adain_tail_save_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "snb_adain_tail_scratch_floats": (c_size_t, [c_int, c_int, c_int, c_int]),
    "snb_adain_tail_fwd": (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P, P, P, P]),
    "snb_adain_tail_bwd": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P, P, P, P, P, P, P, P, P]),
    "snb_adam_flat": (c_int, [P, P, P, P, c_size_t, c_float, c_float, c_float, c_float, c_float, c_int, P]),
    "snb_thin_expand": (c_int, [P, ctypes.c_longlong, P, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "snb_thin_reduce": (c_int, [P, ctypes.c_longlong, P, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, P, ctypes.c_longlong, P]),
    "snb_thin_wgrad": (c_int, [P, ctypes.c_longlong, P, ctypes.c_longlong, c_int, c_int, c_int, c_int, P, P]),
    "snb_linear_workspace_floats": (c_size_t, [c_int, c_int, c_int]),
    "snb_linear_fwd": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P]),
    "snb_linear_dgrad": (c_int, [P, P, c_int, c_int, c_int, P, P, P]),
    "snb_linear_wgrad": (c_int, [P, P, c_int, c_int, c_int, P, P, P]),
    "snb_multi_copy": (c_int, [P, P, P, c_int, P]),
    "snb_bn_se_tail_save_floats": (c_size_t, [c_int, c_int, c_int]),
    "snb_bn_se_tail_scratch_floats": (c_size_t, [c_int, c_int, c_int]),
    "snb_bn_se_tail_fwd": (c_int, [P, P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_float, c_int, c_float, c_float, P, P, P, P, P, P, P]),
    "snb_bn_se_tail_bwd": (c_int, [P, P, P, P, c_int, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P, P]),
    "snb_row_act_pool_fwd": (c_int, [P, P, P, ctypes.c_longlong, c_int, c_float, P, P, P, P]),
    "snb_row_act_pool_bwd_reduce": (c_int, [P, P, P, P, P, P, ctypes.c_longlong, c_int, c_float, P, P, P]),
    "snb_row_act_pool_bwd": (c_int, [P, P, P, P, P, P, P, P, P, ctypes.c_longlong, c_int, c_float, P, P]),
    "snb_gridding_fwd": (c_int, [P, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float, P, P, P, P]),
    "snb_gridding_bwd": (c_int, [P, P, P, c_int, c_int, ctypes.c_longlong, P, P]),
    "snb_gridding_rev_fwd": (c_int, [P, c_int, c_int, P, P]),
    "snb_gridding_rev_bwd": (c_int, [P, P, P, c_int, c_int, P, P]),
    "snb_gridding_dist_fwd": (c_int, [P, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float, P, P, P, P]),
    "snb_gridding_dist_bwd": (c_int, [P, P, P, c_int, c_int, ctypes.c_longlong, P, P]),
    "snb_cubic_sampling_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "snb_cubic_sampling_bwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P]),
}

_lib = None


class SnbError(RuntimeError):
    pass


def load():
    """Load the shared library (building nothing: run `python -m sparenet_b200.build` or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m sparenet_b200.build` (nvcc, sm_100a). "
                "sparenet_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().snb_strerror(rc).decode()
        raise SnbError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
