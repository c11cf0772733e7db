"""ctypes loader for libbbmapcuda.so (C ABI in include/bbmap_cuda.h).

There is no CPU fallback anywhere in this package: if the CUDA extension is missing or no
device is visible, construction raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BBM_SO") or os.path.join(HERE, "libbbmapcuda.so")   # BBM_SO: A/B builds while tuning

BBM_OK, BBM_E_NODEVICE, BBM_E_CUDA, BBM_E_ARG, BBM_E_SHAPE, BBM_E_CAPACITY = 0, -1, -2, -3, -4, -5

# every symbol include/bbmap_cuda.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "bbm_init", "bbm_destroy", "bbm_set_band", "bbm_last_error", "bbm_device_count", "bbm_upload", "bbm_free_dev",
    "bbm_msa_batch_dev", "bbm_msa_batch_host", "bbm_launch_count", "bbm_set_option", "bbm_get_stat", "bbm_int_peak", "bbm_banded_batch_dev", "bbm_banded_batch_host", "bbm_seed_batch_dev", "bbm_seed_batch_host", "bbm_noindel_batch_dev", "bbm_noindel_batch_host", "bbm_index_build", "bbm_index_block_sites", "bbm_index_download", "bbm_index_save", "bbm_index_load", "bbm_wire_last_error", "bbm_wire_free", "bbm_wire_write_int_array", "bbm_wire_read_int_array", "bbm_wire_block_fname", "bbm_wire_write_block", "bbm_wire_read_block", "bbm_wire_write_chrom", "bbm_wire_read_chrom", "bbm_wire_write_summary", "bbm_wire_read_summary", "bbm_break_reads", "bbm_search_batch_dev", "bbm_search_batch_host", "bbm_msa_gapped_batch_dev", "bbm_msa_gapped_batch_host", "bbm_ingest_batch_dev", "bbm_ingest_batch_host", "bbm_sam_batch_dev", "bbm_sam_batch_host", "bbm_tipdel_batch_dev", "bbm_tipdel_batch_host", "bbm_rescue_batch_dev", "bbm_rescue_batch_host", "bbm_sitelist_from_search_dev", "bbm_sitelist_batch_dev", "bbm_sitelist_batch_host", "bbm_scoreslow_dev", "bbm_scoreslow_host", "bbm_sitelist_tipdel_dev", "bbm_sitelist_bounds_dev",
    "bbm_sitelist_clearzone3_dev", "bbm_sitelist_tip_penalty_dev", "bbm_sam_tasks_from_lists_dev",
    "bbm_map_set_scaffolds", "bbm_map_batch_dev", "bbm_map_batch_host", "bbm_index_share",
    "bbm_fillUnlimited", "bbm_fillLimitedX",
]


class BbmError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (building is __graft_entry__.build()'s job). Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise BbmError("libbbmapcuda.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                       "bbmap_b200 has no CPU fallback")
    L = C.CDLL(SO_PATH)
    L.bbm_last_error.restype = C.c_char_p
    L.bbm_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.bbm_destroy.argtypes = [C.c_void_p]
    L.bbm_destroy.restype = None
    L.bbm_set_band.argtypes = [C.c_void_p, C.c_int32, C.c_float]
    L.bbm_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
    L.bbm_free_dev.argtypes = [C.c_void_p, C.c_void_p]
    L.bbm_launch_count.argtypes = [C.c_void_p]
    L.bbm_launch_count.restype = C.c_int64
    L.bbm_msa_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_msa_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p]
    L.bbm_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.bbm_get_stat.argtypes = [C.c_void_p, C.c_char_p]
    L.bbm_get_stat.restype = C.c_int64
    L.bbm_int_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.bbm_banded_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
    L.bbm_banded_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_seed_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32] + [C.c_void_p] * 7
    L.bbm_seed_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.POINTER(C.c_float)]
    L.bbm_noindel_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.bbm_index_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.bbm_index_block_sites.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
    L.bbm_index_download.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bbm_index_save.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    L.bbm_index_load.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.bbm_wire_last_error.restype = C.c_char_p
    L.bbm_wire_free.argtypes = [C.c_void_p]; L.bbm_wire_free.restype = None
    L.bbm_wire_write_int_array.argtypes = [C.c_char_p, C.c_void_p, C.c_int64]
    L.bbm_wire_read_int_array.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.bbm_wire_block_fname.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.bbm_wire_write_block.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    L.bbm_wire_read_block.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.bbm_wire_write_chrom.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int8]
    L.bbm_wire_read_chrom.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int8)]
    L.bbm_wire_write_summary.argtypes = [C.c_char_p, C.c_void_p]
    L.bbm_wire_read_summary.argtypes = [C.c_char_p, C.c_void_p]
    L.bbm_break_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)] + [C.c_void_p] * 7
    L.bbm_search_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
    L.bbm_msa_gapped_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.bbm_ingest_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    L.bbm_search_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_noindel_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_ingest_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_msa_gapped_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_sam_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bbm_sam_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    for nm in ("tipdel", "rescue"):
        getattr(L, f"bbm_{nm}_batch_host").argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        getattr(L, f"bbm_{nm}_batch_dev").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_sitelist_from_search_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.bbm_sitelist_batch_dev.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 7 + [C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_sitelist_batch_host.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p, C.c_void_p]
    L.bbm_scoreslow_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bbm_scoreslow_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 8 + [C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    L.bbm_sitelist_tipdel_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 10 + [C.POINTER(C.c_float)]
    L.bbm_sitelist_bounds_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.bbm_sitelist_clearzone3_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.bbm_sitelist_tip_penalty_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bbm_sam_tasks_from_lists_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 5
    L.bbm_index_share.argtypes = [C.c_void_p, C.c_void_p]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise BbmError("%s failed (%d): %s" % (what, rc, load().bbm_last_error().decode()))
