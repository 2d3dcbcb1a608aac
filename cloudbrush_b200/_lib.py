"""ctypes binding of libcloudbrush_gpu.so (the C ABI declared in include/cloudbrush_gpu.h).

There is no CPU fallback: if the shared library is missing the import of this module fails, and
if no CUDA device is present cb_create returns CB_E_NODEVICE, which surfaces as BrushGpuError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB_LIB") or os.path.join(_HERE, "libcloudbrush_gpu.so")  # CB_LIB: an instrumented build

CB_ABI_VERSION = 1
CB_OK = 0
ERRORS = {-1: "CB_E_INVALID", -2: "CB_E_NODEVICE", -3: "CB_E_CUDA", -4: "CB_E_NOMEM", -5: "CB_E_IO",
          -6: "CB_E_NOREADS", -7: "CB_E_UNSUPPORTED", -8: "CB_E_CAPACITY"}
CK_PREPROCESS, CK_OVERLAP, CK_CHL2, CK_RE = 0, 1, 2, 3
FLAG_FLAT_PIPELINE = 1


class cb_config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("k", C.c_int32), ("readlen", C.c_int32),
                ("up_kmer", C.c_int64), ("low_kmer", C.c_int64), ("majority", C.c_float), ("pwm_n", C.c_float),
                ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("flags", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class cb_edge(C.Structure):
    _fields_ = [("node", C.c_uint32), ("nbr", C.c_uint32), ("ov", C.c_uint16), ("type", C.c_uint8),
                ("pad", C.c_uint8)]


class cb_node(C.Structure):
    _fields_ = [("line", C.c_uint32), ("cov", C.c_uint32)]


# every symbol include/cloudbrush_gpu.h declares
SYMBOLS = ["cb_default_config", "cb_create", "cb_destroy", "cb_last_error", "cb_version", "cb_load_sfa",
           "cb_load_sfa_buffer", "cb_load_sfa_device", "cb_preprocess", "cb_build_overlap", "cb_string_graph",
           "cb_get_counter", "cb_export_edges", "cb_export_nodes", "cb_write_nodes", "cb_get_time_ms",
           "cb_get_launch_count", "cb_borrow_sfa_device", "cb_set_profile", "cb_get_kernel_stat", "cb_set_exchange",
           "cb_nccl_get_unique_id", "cb_init_nccl", "cb_debug_sort", "cb_prefetch_sfa_buffer", "cb_nccl_last_error",
           "cb_debug_bucket_plan", "cb_debug_partition", "cb_write_fasta", "cb_export_nodes_begin",
           "cb_export_edges_begin", "cb_export_wait"]

_lib = None


def load():
    """Loads the shared library (no device needed for this) and sets the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C cloudbrush_b200/csrc`. There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    L.cb_default_config.argtypes = [C.POINTER(cb_config)]
    L.cb_default_config.restype = None
    L.cb_create.argtypes = [C.POINTER(cb_config), C.POINTER(P)]
    L.cb_destroy.argtypes = [P]
    L.cb_destroy.restype = None
    L.cb_last_error.argtypes = [P]
    L.cb_last_error.restype = C.c_char_p
    L.cb_version.restype = C.c_char_p
    L.cb_load_sfa.argtypes = [P, C.c_char_p]
    L.cb_load_sfa_buffer.argtypes = [P, C.c_void_p, C.c_size_t]
    L.cb_load_sfa_device.argtypes = [P, C.c_void_p, C.c_size_t]
    L.cb_prefetch_sfa_buffer.argtypes = [P, C.c_void_p, C.c_size_t]
    for f in (L.cb_preprocess, L.cb_build_overlap, L.cb_string_graph):
        f.argtypes = [P]
    L.cb_get_counter.argtypes = [P, C.c_char_p, C.POINTER(C.c_int64)]
    L.cb_export_edges.argtypes = [P, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.cb_export_nodes.argtypes = [P, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.cb_export_edges_begin.argtypes = [P, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.cb_export_nodes_begin.argtypes = [P, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.cb_export_wait.argtypes = [P]
    L.cb_write_nodes.argtypes = [P, C.c_int, C.c_char_p]
    L.cb_write_fasta.argtypes = [P, C.c_char_p]
    L.cb_get_time_ms.argtypes = [P, C.c_char_p, C.POINTER(C.c_double)]
    L.cb_get_launch_count.argtypes = [P, C.POINTER(C.c_int64)]
    L.cb_borrow_sfa_device.argtypes = [P, C.c_void_p, C.c_size_t]
    L.cb_set_exchange.argtypes = [P, C.c_void_p]
    L.cb_set_profile.argtypes = [P, C.c_int]
    L.cb_nccl_get_unique_id.argtypes = [C.c_void_p]
    L.cb_init_nccl.argtypes = [P, C.c_void_p]
    L.cb_nccl_last_error.restype = C.c_char_p
    L.cb_debug_sort.argtypes = [P, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.cb_debug_bucket_plan.argtypes = [C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_int32)]
    L.cb_debug_partition.argtypes = [P, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]
    L.cb_get_kernel_stat.argtypes = [P, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    _lib = L
    return L
