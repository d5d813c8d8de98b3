"""ctypes binding of libnts_b200.so (include/nts_b200.h). No compute happens in Python and
there is no fallback: if the CUDA library is missing or a call fails, an exception is raised."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NB_LIB_PATH") or os.path.join(_HERE, "lib", "libnts_b200.so")   # NB_LIB_PATH: A/B runs of two builds
HEADER = os.path.join(os.path.dirname(_HERE), "include", "nts_b200.h")

NB_WEIGHT_SUM, NB_WEIGHT_MEAN, NB_WEIGHT_NONE, NB_WEIGHT_MEAN_SAMPLED = 0, 1, 2, 3
NB_SAMPLER_MERGE_SRC_DST, NB_SAMPLER_UP_DEGREE, NB_SAMPLER_BUILD_CSR, NB_SAMPLER_NO_BOTTOM_CSR = 1, 2, 4, 8


class NtsError(RuntimeError):
    pass


class LayerView(C.Structure):
    _fields_ = [("n_dst", C.c_uint32), ("n_edges", C.c_uint32), ("n_src", C.c_uint32), ("reserved", C.c_uint32),
                ("destination", C.c_void_p), ("column_offset", C.c_void_p), ("sample_ans", C.c_void_p),
                ("row_indices", C.c_void_p), ("source", C.c_void_p), ("row_offset", C.c_void_p),
                ("column_indices", C.c_void_p), ("csr_to_csc", C.c_void_p), ("edge_weight_forward", C.c_void_p),
                ("edge_weight_backward", C.c_void_p), ("dst_local_id", C.c_void_p), ("src_to_dst", C.c_void_p),
                ("source_use_count", C.c_void_p), ("gather_index", C.c_void_p)]


def header_symbols():
    """Every function the header declares (used by the CPU test that the .so exports them all)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", text)))


_lib = None

P, U32, U64, I32, F32, SZ = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float, C.c_size_t
_SIGS = {
    "nb_abi_version": (I32, []),
    "nb_last_error": (C.c_char_p, []),
    "nb_device_count": (I32, [C.POINTER(I32)]),
    "nb_set_option": (I32, [C.c_char_p, I32]),
    "nb_mirror_invalidate": (I32, [P]),
    "nb_ctx_create": (I32, [I32, P, I32, C.POINTER(P)]),
    "nb_ctx_destroy": (I32, [P]),
    "nb_ctx_set_stream": (I32, [P, P]),
    "nb_ctx_stream": (P, [P]),
    "nb_ctx_device": (I32, [P]),
    "nb_ctx_synchronize": (I32, [P]),
    "nb_ctx_launch_count": (U64, [P]),
    "nb_malloc_pinned": (I32, [SZ, C.POINTER(P)]),
    "nb_free_host": (I32, [P]),
    "nb_device_pointer": (I32, [P, C.POINTER(P)]),
    "nb_malloc_device": (I32, [SZ, C.POINTER(P)]),
    "nb_free_device": (I32, [P]),
    "nb_memcpy_h2d": (I32, [P, P, P, SZ, I32]),
    "nb_memcpy_d2h": (I32, [P, P, P, SZ, I32]),
    "nb_memset_async": (I32, [P, P, I32, SZ]),
    "nb_read_feature_table": (I32, [C.c_char_p, U32, U32, U32, U32, P, I32, C.POINTER(I32)]),
    "nb_read_label_mask": (I32, [C.c_char_p, C.c_char_p, U32, U32, P, P]),
    "nb_graph_create": (I32, [P, U32, U64, P, P, P, P, C.POINTER(P)]),
    "nb_graph_create_from_pairs": (I32, [P, U32, U64, P, I32, C.POINTER(P)]),
    "nb_graph_create_from_device": (I32, [P, U32, U64, P, P, C.POINTER(P)]),
    "nb_graph_destroy": (I32, [P]),
    "nb_graph_info": (I32, [P, C.POINTER(U32), C.POINTER(U64), C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "nb_sampler_create": (I32, [P, P, I32, C.POINTER(I32), U32, U32, U64, C.POINTER(P)]),
    "nb_sampler_destroy": (I32, [P]),
    "nb_sampler_sample": (I32, [P, P, U32, I32, U64, U64, I32, P, U32, C.POINTER(LayerView), I32]),
    "nb_sampler_replay": (I32, [P, P, U32, C.POINTER(P), C.POINTER(U32), I32, C.POINTER(LayerView)]),
    "nb_sampler_wait": (I32, [P, C.POINTER(LayerView)]),
    "nb_sampler_layer": (I32, [P, I32, C.POINTER(LayerView)]),
    "nb_sampler_sizes_dev": (I32, [P, I32, C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(U32), C.POINTER(U32), C.POINTER(U32)]),
    "nb_sample_count": (I32, [P, P, P, P, U32, U32, P, U32, C.POINTER(U32)]),
    "nb_sample_traverse": (I32, [P, P, P, P, P, P, P, U32, U32, U32, P, P, U32, U32, I32, U64, U64]),
    "nb_sample_update_ri": (I32, [P, P, P, U32]),
    "nb_set_dst_local_index": (I32, [P, P, P, U32, P]),
    "nb_update_degree": (I32, [P, P, P, U32, U32, P, P, P, P, I32]),
    "nb_edge_weight": (I32, [P, P, P, P, U32, P, P, P, P, I32]),
    "nb_hotness": (I32, [P, P, P, U32, I32, I32, F32, U32, P, U32, C.POINTER(U32), P]),
    "nb_set_cache_index": (I32, [P, P, P, U32, P, U32]),
    "nb_gather_rows": (I32, [P, P, P, P, U32, U32, U32, U32]),
    "nb_gather_rows_dyn": (I32, [P, P, P, P, P, U32, U32, U32, U32]),
    "nb_aggregate_csc_fwd_dyn": (I32, [P, P, P, P, P, P, P, U32, U32, U32, U32]),
    "nb_aggregate_csr_bwd_dyn": (I32, [P, P, P, P, P, P, P, U32, U32, U32, U32]),
    "nb_gather_rows_cached": (I32, [P, P, P, U32, P, U32, P, P, U32, U32, U32, P]),
    "nb_gather_rows_indexed": (I32, [P, P, U32, P, U32, P, P, P, U32, U32]),
    "nb_gather_labels": (I32, [P, P, P, P, U32]),
    "nb_row_override": (I32, [P, P, P, P, P, P, U32, U32, U32]),
    "nb_row_override2": (I32, [P, P, P, P, P, P, P, P, U32, U32, U32, U32]),
    "nb_stage_create": (I32, [P, P, U32, U32, U32, C.POINTER(P)]),
    "nb_stage_destroy": (I32, [P]),
    "nb_stage_submit": (I32, [P, I32, P, U32, P]),
    "nb_stage_gather": (I32, [P, I32, P, U32, P, U32, P, P, C.POINTER(U32)]),
    "nb_stage_gather_table": (I32, [P, I32, P, U32, P, P, P, C.POINTER(U32)]),
    "nb_peer_comm_block_bytes": (SZ, [U64, U32]),
    "nb_peer_comm_create": (I32, [P, U32, U32, U64, C.POINTER(P), C.POINTER(P)]),
    "nb_peer_comm_destroy": (I32, [P]),
    "nb_peer_allreduce_sum": (I32, [P, P, U64]),
    "nb_peer_comm_check": (I32, [P, C.POINTER(I32)]),
    "nb_peer_allreduce_begin": (I32, [P, P, U64]),
    "nb_peer_allreduce_end": (I32, [P, P, U64]),
    "nb_peer_comm_stats": (I32, [P, C.POINTER(U64), C.POINTER(U64), C.POINTER(U64), I32]),
    "nb_trace_dump": (I32, []),
    "nb_trace_reset": (I32, []),
    "nb_table_create": (I32, [P, U32, C.POINTER(P), U32, U32, U64, C.POINTER(P)]),
    "nb_table_destroy": (I32, [P]),
    "nb_table_gather": (I32, [P, P, P, P, U32, U32]),
    "nb_ipc_get_handle": (I32, [P, P]),
    "nb_ipc_open_handle": (I32, [P, C.POINTER(P)]),
    "nb_ipc_close_handle": (I32, [P]),
    "nb_vmm_padded_size": (SZ, [P, SZ]),
    "nb_vmm_alloc": (I32, [P, SZ, C.POINTER(P), C.POINTER(I32)]),
    "nb_vmm_import": (I32, [P, I32, SZ, C.POINTER(P)]),
    "nb_vmm_grant": (I32, [P, I32]),
    "nb_vmm_free": (I32, [P]),
    "nb_aggregate_csc_fwd": (I32, [P, P, P, P, P, P, U32, U32, U32]),
    "nb_aggregate_csr_bwd": (I32, [P, P, P, P, P, P, U32, U32, U32]),
    "nb_aggregate_push_bwd": (I32, [P, P, P, P, P, P, U32, U32, U32]),
    "nb_aggregate_gathered_fwd_dyn": (I32, [P, P, U32, P, P, P, P, P, U32, U32, U32]),
    "nb_scatter_src_dst_to_msg": (I32, [P, P, P, P, P, U32, U32, P]),
    "nb_gather_msg_to_src_dst": (I32, [P, P, P, P, P, U32, U32, U32, P]),
    "nb_edge_softmax_fwd": (I32, [P, P, P, P, P, U32]),
    "nb_edge_softmax_bwd": (I32, [P, P, P, P, P, U32]),
    "nb_gather_msg_to_dst": (I32, [P, P, P, P, U32, U32]),
    "nb_scatter_dst_to_msg": (I32, [P, P, P, P, U32, U32]),
    "nb_gat_fwd": (I32, [P, P, P, F32, P, P, P, U32, U32, U32, P, P, P]),
    "nb_gat_bwd": (I32, [P, P, P, F32, P, P, P, P, P, P, P, P, P, P, U32, U32, U32, U32, P, P]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NtsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().nb_last_error()
        raise NtsError(f"libnts_b200 error {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """device/host address of a torch tensor, numpy array, int or None"""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    if hasattr(t, "ctypes"):
        return t.ctypes.data
    raise TypeError(type(t))
