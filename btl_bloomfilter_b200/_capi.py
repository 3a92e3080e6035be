"""ctypes declarations of the C ABI in include/btlbf.h (libbtlbf_cuda.so).

There is no Python / numpy / CPU implementation behind these calls: if the CUDA library is missing
the import of the product classes fails loudly."""
import ctypes as C
import os

from . import _build

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u64 = C.c_uint64
u32 = C.c_uint
vp = C.c_void_p
cpp = C.POINTER(C.c_char_p)

BTLBF_OK, BTLBF_ERR_ARG, BTLBF_ERR_CUDA, BTLBF_ERR_NOMEM, BTLBF_ERR_STATE = range(5)
BLOOM, COUNTING8 = 0, 1

# name -> argtypes; every function returns int status unless listed in _RESTYPES
SIGNATURES = {
    "btlbf_version": [],
    "btlbf_device_count": [C.POINTER(C.c_int)],
    "btlbf_ctx_create": [C.c_int, C.POINTER(vp)],
    "btlbf_ctx_destroy": [vp],
    "btlbf_ctx_set_stream": [vp, vp],
    "btlbf_ctx_sync": [vp],
    "btlbf_ctx_flush": [vp],
    "btlbf_ctx_aux_stream": [vp, C.POINTER(vp)],
    "btlbf_ctx_launch_count": [vp, u64p],
    "btlbf_ctx_counter": [vp, C.c_char_p, u64p],
    "btlbf_ctx_set_option": [vp, C.c_char_p, C.c_int64],
    "btlbf_filter_create": [vp, C.c_int, u64, u32, u32, u32, C.POINTER(vp)],
    "btlbf_filter_wrap": [vp, C.c_int, u64, u32, u32, u32, vp, u64, C.POINTER(vp)],
    "btlbf_filter_destroy": [vp],
    "btlbf_filter_clear": [vp],
    "btlbf_filter_info": [vp, C.POINTER(C.c_int), u64p, u64p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)],
    "btlbf_filter_set_threshold": [vp, u32],
    "btlbf_filter_upload": [vp, vp, u64],
    "btlbf_filter_download": [vp, vp, u64],
    "btlbf_filter_device_ptr": [vp, C.POINTER(vp), u64p],
    "btlbf_filter_popcount": [vp, u64p],
    "btlbf_filter_count_ge": [vp, u32, u64p],
    "btlbf_filter_set_seeds": [vp, cpp, u32, u32],
    "btlbf_filter_merge_from_device": [vp, vp, u64],
    "btlbf_merge_device_buffers": [vp, C.c_int, vp, vp, u64],
    "btlbf_merge_slice": [u64, C.c_int, C.c_int, u64p, u64p],
    "btlbf_ipc_export": [vp, vp, vp, u64p],
    "btlbf_ipc_open": [vp, vp, C.POINTER(vp)],
    "btlbf_ipc_close": [vp, vp],
    "btlbf_merge_peers": [vp, C.c_int, C.POINTER(vp), C.c_int, C.c_int, u64],
    "btlbf_merge_multimem": [vp, C.c_int, vp, C.c_int, C.c_int, u64],
    "btlbf_merge_hybrid": [vp, C.c_int, vp, C.POINTER(vp), C.c_int, C.c_int, u64, C.c_uint],
    "btlbf_filter_flush_parts": [vp, C.c_uint, C.c_uint, u64p, u64p],
    "btlbf_merge_peers_range": [vp, C.c_int, C.POINTER(vp), C.c_int, C.c_int, u64, u64, vp],
    "btlbf_filter_ordered_stats": [vp, u64p, u64p],
    "btlbf_insert_file": [vp, C.c_char_p, C.c_int, u64p, u64p],
    "btlbf_query_file": [vp, C.c_char_p, C.c_int, u64p, u64p, u64p],
    "btlbf_seqfile_open": [C.c_char_p, u32, C.c_int, C.c_int, C.POINTER(vp)],
    "btlbf_seqfile_next": [vp, vp, u64, u64p, u64, u64p, u64p, u64p, C.POINTER(C.c_int)],
    "btlbf_seqfile_close": [vp],
    "btlbf_ingest_release": [],
    "btlbf_filter_ctx": [vp, C.POINTER(vp)],
    "btlbf_insert_seqs": [vp, vp, u64p, u64, u64p],
    "btlbf_contains_seqs": [vp, vp, u64p, u64, vp, vp, u64p, u64p],
    "btlbf_insert_seqs_async": [vp, vp, u64p, u64, vp],
    "btlbf_contains_seqs_async": [vp, vp, u64p, u64, vp, vp, vp],
    "btlbf_pack_seqs": [vp, u64, vp, vp, C.c_int, u64p],
    "btlbf_insert_seqs_packed": [vp, vp, vp, u64p, u64, u64p],
    "btlbf_contains_seqs_packed": [vp, vp, vp, u64p, u64, vp, vp, u64p, u64p],
    "btlbf_insert_seqs_packed_async": [vp, vp, vp, u64p, u64, vp],
    "btlbf_contains_seqs_packed_async": [vp, vp, vp, u64p, u64, vp, vp, vp],
    "btlbf_insert_seqs_packed_dev": [vp, vp, vp, u64, vp, u64, vp],
    "btlbf_contains_seqs_packed_dev": [vp, vp, vp, u64, vp, u64, vp, vp, vp],
    "btlbf_insert_and_check_seqs": [vp, vp, u64p, u64, vp, vp, u64p],
    "btlbf_mincount_seqs": [vp, vp, u64p, u64, vp, vp, u64p],
    "btlbf_increment_all_seqs": [vp, vp, u64p, u64, u64p],
    "btlbf_hash_seqs": [vp, u32, u32, cpp, u32, u32, vp, u64p, u64, vp, vp, vp, u64p],
    "btlbf_insert_hashes": [vp, u64p, u64, vp],
    "btlbf_contains_hashes": [vp, u64p, u64, vp],
    "btlbf_mincount_hashes": [vp, u64p, u64, vp],
    "btlbf_increment_all_hashes": [vp, u64p, u64],
    "btlbf_filter_store": [vp, C.c_char_p, C.c_double, u64, u64],
    "btlbf_filter_load": [vp, C.c_char_p, C.c_int, u32, C.POINTER(vp), C.POINTER(C.c_double), u64p, u64p],
    "btlbf_format_header": [C.c_int, u64, u64, u32, u32, C.c_double, u64, u64, C.c_char_p, C.c_size_t,
                            C.POINTER(C.c_size_t)],
    "btlbf_parse_header": [C.c_int, C.c_char_p, C.c_size_t, u64p, u64p, C.POINTER(u32), C.POINTER(u32),
                           C.POINTER(C.c_double), u64p, u64p, C.POINTER(C.c_size_t)],
    "btlbf_insert_seqs_dev": [vp, vp, u64, vp, u64, vp],
    "btlbf_contains_seqs_dev": [vp, vp, u64, vp, u64, vp, vp, vp],
    "btlbf_mincount_seqs_dev": [vp, vp, u64, vp, u64, vp, vp, vp],
    "btlbf_synth_genome_dev": [vp, vp, u64, u64, u64],
    "btlbf_synth_reads_dev": [vp, vp, u64, u64, u32, u64, u64, u64, u64],
    "btlbf_random_access_probe": [vp, vp, u64, u64, C.c_int, C.POINTER(C.c_float)],
}

_lib = None


class BtlbfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("btlbf error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libbtlbf_cuda.so (building it with nvcc if it is missing).  Raises if that fails."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        _build.build_library()
    L = C.CDLL(path)
    L.btlbf_last_error.restype = C.c_char_p
    L.btlbf_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here means the library does not match include/btlbf.h
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != BTLBF_OK:
        raise BtlbfError(rc, lib().btlbf_last_error().decode(errors="replace"))
