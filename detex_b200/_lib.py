"""ctypes binding of the C ABI in include/detex_b200.h (libdetex_b200.so, in-tree).

There is no CPU fallback: if the shared library is missing it is built with nvcc, and if
that fails or no sm_100 GPU is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import _build

DTX_OK = 0
DTX_ERR_SHORT_CHUNK = 4
DTX_ERR_CAPACITY = 6
DTX_F64, DTX_F32 = 0, 1
ENGINE_TCGEN05, ENGINE_FP64, ENGINE_TCGEN05_X8, ENGINE_TCGEN05_AUTO = 0, 1, 2, 3
ENGINES = {"tcgen05": ENGINE_TCGEN05, "fp64": ENGINE_FP64, "tcgen05_x8": ENGINE_TCGEN05_X8,
           "tcgen05_auto": ENGINE_TCGEN05_AUTO}
HIST_BINS = 400

EXPORTS = [
    "dtx_version", "dtx_create", "dtx_destroy", "dtx_last_error", "dtx_sync", "dtx_set_bases",
    "dtx_load_chunks", "dtx_attach_device_chunks", "dtx_preprocess_chunks", "dtx_get_chunk", "dtx_detect_run", "dtx_num_lags", "dtx_get_ds",
    "dtx_get_ds64", "dtx_get_stalta", "dtx_get_rowstats", "dtx_get_hist", "dtx_get_fas", "dtx_get_candidates",
    "dtx_sta_lta_max", "dtx_set_events", "dtx_est_mags", "dtx_last_k1_ms", "dtx_launch_count", "dtx_ccx", "dtx_corr_zero_lag",
    "dtx_set_x8_tolerance", "dtx_get_chunk_modes", "dtx_set_trigger_sta", "dtx_preprocess_chunks_dec", "dtx_set_hist_bins",
    "dtx_accumulate_begin", "dtx_accumulate_end", "dtx_k1_ms_history", "dtx_ccx_device", "dtx_ccx_pack",
    "dtx_ccx_condensed", "dtx_set_ccx_batch", "dtx_host_alloc", "dtx_host_free", "dtx_set_core_lags", "dtx_set_ccx_passes", "dtx_set_fused",
    "dtx_ccx_pack_rows", "dtx_host_register", "dtx_host_unregister",
]


class Cand(C.Structure):
    _fields_ = [("row", C.c_int32), ("t", C.c_int32), ("ds", C.c_float), ("lta", C.c_float)]


CAND_DTYPE = np.dtype([("row", np.int32), ("t", np.int32), ("ds", np.float32), ("lta", np.float32)])

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building if needed) the shared library and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DETEX_B200_LIB", _build.LIB)   # experiment variants (see experiments/ab_issue.sh)
    if path == _build.LIB and not os.path.exists(path):
        _build.build()
    L = C.CDLL(path)
    p = C.c_void_p
    L.dtx_version.restype = C.c_int
    L.dtx_create.argtypes = [C.c_int, p, C.POINTER(p)]
    L.dtx_destroy.argtypes = [p]
    L.dtx_destroy.restype = None
    L.dtx_last_error.argtypes = [p]
    L.dtx_last_error.restype = C.c_char_p
    L.dtx_sync.argtypes = [p]
    L.dtx_set_x8_tolerance.argtypes = [p, C.c_double]
    L.dtx_get_chunk_modes.argtypes = [p, p]
    L.dtx_set_trigger_sta.argtypes = [p, C.c_int]
    L.dtx_set_hist_bins.argtypes = [p, C.c_int]
    L.dtx_set_bases.argtypes = [p, C.c_int, p, p, C.c_int, C.c_int, C.c_int, p]
    L.dtx_load_chunks.argtypes = [p, C.c_int, p, p, C.c_int]
    L.dtx_attach_device_chunks.argtypes = [p, C.c_int, p, p, p, C.c_int]
    L.dtx_preprocess_chunks.argtypes = [p, C.c_int, C.c_int, p, p, C.c_int, p, C.c_int, C.c_int, C.c_int]
    L.dtx_preprocess_chunks_dec.argtypes = [p, C.c_int, C.c_int, p, p, C.c_int, p, C.c_int, C.c_int, C.c_int,
                                            p, C.c_int, C.c_int]
    L.dtx_get_chunk.argtypes = [p, C.c_int, p, C.c_int64, C.POINTER(C.c_int64)]
    L.dtx_detect_run.argtypes = [p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                 C.c_int, C.c_int]
    L.dtx_num_lags.argtypes = [p, C.c_int, C.POINTER(C.c_int64)]
    L.dtx_get_ds.argtypes = [p, C.c_int, C.c_int, p, C.c_int64]
    L.dtx_get_ds64.argtypes = [p, C.c_int, C.c_int, p, C.c_int64]
    L.dtx_get_stalta.argtypes = [p, C.c_int, C.c_int, C.c_int, p, C.c_int64]
    L.dtx_get_rowstats.argtypes = [p, p, p, C.c_int64]
    L.dtx_get_hist.argtypes = [p, C.c_int, p, C.c_int64, C.c_int]
    L.dtx_get_fas.argtypes = [p, C.c_int, p, C.c_int64, C.c_int]
    L.dtx_get_candidates.argtypes = [p, p, C.c_int64, C.POINTER(C.c_int64)]
    L.dtx_last_k1_ms.argtypes = [p, C.POINTER(C.c_float)]
    L.dtx_sta_lta_max.argtypes = [p, C.c_int, C.c_int, C.c_int, C.c_int, p, C.c_int64]
    L.dtx_set_events.argtypes = [p, C.c_int, C.c_int, C.c_int, p, p, p, C.c_int]
    L.dtx_est_mags.argtypes = [p, C.c_int, C.c_int, p, p, p, p]
    L.dtx_launch_count.argtypes = [p, C.POINTER(C.c_int64)]
    L.dtx_corr_zero_lag.argtypes = [p, p, C.c_int, C.c_int, p]
    L.dtx_ccx.argtypes = [p, p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, p, p, p]
    L.dtx_set_core_lags.argtypes = [p, p, p]
    L.dtx_accumulate_begin.argtypes = [p, C.c_int64]
    L.dtx_accumulate_end.argtypes = [p]
    L.dtx_k1_ms_history.argtypes = [p, p, C.c_int64, C.POINTER(C.c_int64)]
    L.dtx_ccx_device.argtypes = [p, p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, p, C.c_int, C.c_int, p, p, p]
    L.dtx_ccx_pack.argtypes = [p, p, p, p, p, C.c_int, C.c_int, p, p, p]
    L.dtx_ccx_condensed.argtypes = [p, p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, p, p, p]
    L.dtx_set_ccx_batch.argtypes = [p, C.c_int, C.c_int64]
    L.dtx_set_ccx_passes.argtypes = [p, C.c_int]
    L.dtx_set_fused.argtypes = [p, C.c_int]
    L.dtx_host_alloc.argtypes = [C.POINTER(p), C.c_int64]
    L.dtx_host_free.argtypes = [p]
    L.dtx_host_register.argtypes = [p, C.c_int64]
    L.dtx_host_unregister.argtypes = [p]
    L.dtx_ccx_pack_rows.argtypes = [p, p, p, p, p, C.c_int, C.c_int, p, p, p]
    for name in EXPORTS:
        if name not in ("dtx_destroy", "dtx_last_error"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


class DtxError(Exception):
    def __init__(self, code, msg):
        Exception.__init__(self, "detex_b200 error %d: %s" % (code, msg))
        self.code = code
