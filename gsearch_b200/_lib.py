"""ctypes binding of libgsearch_b200.so.  Fails loudly if the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "libgsearch_b200.so")


class GsbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"gsearch_b200 status {status}: {msg}")
        self.status = status


class SketchParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("kmer_size", "sketch_size", "algo", "data_t", "block_flag", "spec_flags")]


class IndexParams(C.Structure):
    _fields_ = [("max_nb_connection", C.c_uint32), ("capacity", C.c_uint64), ("max_layer", C.c_uint32),
                ("ef_construction", C.c_uint32), ("scale_modification", C.c_double),
                ("sig_type", C.c_uint32), ("sketch_size", C.c_uint32), ("extend_candidates", C.c_uint32),
                ("keep_pruned", C.c_uint32), ("level_seed", C.c_uint64)]


class NeighbourC(C.Structure):
    _fields_ = [("d_id", C.c_uint64), ("distance", C.c_float), ("layer", C.c_uint8),
                ("pad_", C.c_uint8 * 3), ("rank", C.c_int32)]


# every symbol include/gsearch_b200.h declares: name -> (restype, argtypes)
_vp, _u32, _u64, _int = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
SYMBOLS = {
    "gsb_last_error": (C.c_char_p, []),
    "gsb_version": (C.c_char_p, []),
    "gsb_device_count": (_int, []),
    "gsb_sketcher_create": (_int, [C.POINTER(SketchParams), _int, C.POINTER(_vp)]),
    "gsb_sketcher_destroy": (None, [_vp]),
    "gsb_sketcher_sig_type": (_int, [_vp]),
    "gsb_sketcher_elem_size": (_u32, [_vp]),
    "gsb_sketch_fasta_batch": (_int, [_vp, _vp, _vp, _u32, _vp, _vp]),
    "gsb_sketch_fasta_batch_dev": (_int, [_vp, _vp, _vp, _u32, _vp, _vp, _vp]),
    "gsb_sketch_fasta_batch_to_dev": (_int, [_vp, _vp, _vp, _u32, _vp, _vp]),
    "gsb_sketcher_launch_count": (_u64, [_vp]),
    "gsb_sketcher_retry_count": (_u64, [_vp]),
    "gsb_sketcher_fallback_count": (_u64, [_vp]),
    "gsb_sketcher_set_prob_path": (_int, [_vp, _int]),
    "gsb_sketcher_enable_timing": (None, [_vp, _int]),
    "gsb_sketcher_kernel_times": (None, [_vp, _vp, _vp]),
    "gsb_dist_hamming_u16": (C.c_float, [_vp, _vp, C.c_ulonglong]),
    "gsb_dist_hamming_u32": (C.c_float, [_vp, _vp, C.c_ulonglong]),
    "gsb_dist_hamming_u64": (C.c_float, [_vp, _vp, C.c_ulonglong]),
    "gsb_dist_hamming_f32": (C.c_float, [_vp, _vp, C.c_ulonglong]),
    "gsb_hamming_batch": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _int]),
    "gsb_hamming_matrix": (_int, [_vp, _u32, _vp, _u32, _u32, _u32, _vp, _int]),
    "gsb_hamming_matrix_dev": (_int, [_vp, _u32, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gsb_index_create": (_int, [C.POINTER(IndexParams), _int, C.POINTER(_vp)]),
    "gsb_index_destroy": (None, [_vp]),
    "gsb_index_insert_batch": (_int, [_vp, _vp, _vp, _u64]),
    "gsb_index_insert_batch_dev": (_int, [_vp, _vp, _vp, _u64]),
    "gsb_index_search_batch": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "gsb_index_search_batch_dev": (_int, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "gsb_index_insert_batch_sharded": (_int, [_vp, _vp, _vp, _vp, _u64]),
    "gsb_comm_unique_id": (_int, [_vp]),
    "gsb_comm_create": (_int, [_vp, _int, _int, _int, C.POINTER(_vp)]),
    "gsb_comm_destroy": (None, [_vp]),
    "gsb_comm_rank": (_int, [_vp]),
    "gsb_comm_size": (_int, [_vp]),
    "gsb_comm_all_gather": (_int, [_vp, _vp, _vp, _u64, _vp]),
    "gsb_comm_broadcast": (_int, [_vp, _vp, _u64, _int, _vp]),
    "gsb_comm_all_gather_rows": (_int, [_vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp]),
    "gsb_device_malloc": (_int, [_int, _u64, C.POINTER(_vp)]),
    "gsb_device_free": (None, [_int, _vp]),
    "gsb_memcpy_h2d": (_int, [_int, _vp, _vp, _u64]),
    "gsb_memcpy_d2h": (_int, [_int, _vp, _vp, _u64]),
    "gsb_host_alloc_pinned": (_int, [_u64, C.POINTER(_vp)]),
    "gsb_host_free_pinned": (None, [_vp]),
    "gsb_index_nb_point": (_u64, [_vp]),
    "gsb_index_load_graph": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _u64]),
    "gsb_index_graph_sizes": (_int, [_vp, _vp, _vp]),
    "gsb_index_export_graph": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gsb_index_export_signatures": (_int, [_vp, _vp]),
    "gsb_index_set_wave_max": (_int, [_vp, _u32]),
    "gsb_index_dump": (_int, [_vp, C.c_char_p, C.c_char_p]),
    "gsb_index_load": (_int, [_vp, C.c_char_p, C.c_char_p]),
}

_lib = None


def lib():
    """The loaded library.  No fallback: a missing build is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise GsbError(-1, f"{lib_path} is not built; run `python -c 'import __graft_entry__ as g; "
                               f"g.build()'` or `make -C gsearch_b200/csrc`")
        L = C.CDLL(lib_path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise GsbError(status, lib().gsb_last_error().decode("utf-8", "replace"))


def version():
    return lib().gsb_version().decode()


def device_count():
    return lib().gsb_device_count()
