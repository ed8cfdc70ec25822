"""ctypes binding of the C ABI declared in include/roadsurf_b200.h.

There is no CPU fallback: if libroadsurf_b200.so is missing or no sm_100 device is present,
loading / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

RS_OK = 0
STATUS = {
    0: "RS_OK", -1: "RS_ERR_INVALID_ARG", -2: "RS_ERR_CUDA", -3: "RS_ERR_CAPACITY",
    -4: "RS_ERR_ROTATED", -5: "RS_ERR_NO_DEVICE", -6: "RS_ERR_UNSUPPORTED", -7: "RS_ERR_NOT_PINNED",
    -8: "RS_ERR_NO_NCCL", -9: "RS_ERR_NCCL", -10: "RS_ERR_CODEC",
}
RS_COMM_ID_BYTES = 128
RS_ERR_NOT_PINNED = -7
RS_U8, RS_U16 = 0, 1
RS_HIST_BANDS, RS_HIST_CLASS_SCORE = 0, 1
RS_WINDOW_CROP, RS_WINDOW_FULL, RS_WINDOW_BOUNDLESS = 0, 1, 2
RS_NODATA_RAW, RS_NODATA_NONE, RS_NODATA_ZERO, RS_NODATA_ZERO_MASKED = 0, 1, 2, 3
RS_NSTAT = 9
STAT_COLS = ("count", "min", "max", "sum", "sumsq", "mean", "std", "median", "margin")
RS_NMETRIC = 12
METRIC_COLS = ("P0", "R0", "F0", "P1", "R1", "F1", "Pw", "Rw", "f1w", "Pb", "Rb", "f1b")
RS_VOTE_COUNT, RS_VOTE_SCORE = 0, 1

# every symbol include/roadsurf_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "rs_version", "rs_status_string", "rs_ctx_create", "rs_ctx_destroy", "rs_ctx_sync_status",
    "rs_ctx_last_cuda_error", "rs_ctx_launch_count", "rs_road_bbox_dev", "rs_zonal_hist_dev",
    "rs_zonal_hist_host", "rs_zonal_stats_host", "rs_zonal_stats_stream_host", "rs_zonal_stats_mapped_host", "rs_band_ratio_columns", "rs_band_ratios_dev", "rs_band_ratios_host", "rs_bin_counts_host", "rs_within_host", "rs_overlay_area_host", "rs_host_register", "rs_host_unregister", "rs_assemble_tiles_dev", "rs_assemble_tiles_host", "rs_rasterize_pairs_dev", "rs_rasterize_pairs_host", "rs_finalize_stats_dev",
    "rs_finalize_stats_host", "rs_vote_metrics_dev", "rs_vote_metrics_host", "rs_synth_tiles_dev",
    "rs_extract_pixels_host", "rs_group_hist_host", "rs_vote_table_host", "rs_confusion_metrics_host",
    "rs_pairs_bbox_host", "rs_rescale_u16_dev", "rs_rescale_u16_host", "rs_ks_hist_host",
    "rs_zonal_stats_f32_host", "rs_pairs_bbox_grid_host", "rs_pairs_intersect_host", "rs_clip_rings_host", "rs_decode_segments_dev", "rs_decode_segments_host", "rs_ingest_tiles_host", "rs_zonal_stats_compressed_host", "rs_comm_unique_id", "rs_comm_init", "rs_comm_destroy", "rs_comm_world", "rs_allreduce_accumulators_dev",
)


class RsRoads(C.Structure):
    _fields_ = [("xy", C.c_void_p), ("ring_off", C.c_void_p), ("road_ring_off", C.c_void_p),
                ("road_bbox", C.c_void_p), ("n_roads", C.c_int32), ("n_rings", C.c_int32), ("n_verts", C.c_int32)]


class RsTiles(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("gt", C.c_void_p), ("n_tiles", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("channels", C.c_int32), ("dtype", C.c_int32)]


class RsPairs(C.Structure):
    _fields_ = [("road_pair_off", C.c_void_p), ("pair_tile", C.c_void_p), ("n_pairs", C.c_int32)]


class RsZonalParams(C.Structure):
    _fields_ = [("hist_mode", C.c_int32), ("window_mode", C.c_int32), ("rescale", C.c_int32), ("border_px", C.c_int32),
                ("scale_k", C.c_double * 4), ("scale_off", C.c_double * 4), ("road_slot", C.c_void_p),
                ("min_zero", C.c_void_p)]


class RsLattice(C.Structure):
    _fields_ = [("x0", C.c_double), ("y0", C.c_double), ("tile_w", C.c_double), ("tile_h", C.c_double),
                ("nx", C.c_int32), ("ny", C.c_int32), ("lut", C.c_void_p)]


class NativeError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: {STATUS.get(status, status)}{(' - ' + detail) if detail else ''}")


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the sources are newer) the CUDA library.  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.is_stale():
        try:
            _build.build_native()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(path):
                raise RuntimeError(
                    f"libroadsurf_b200.so is missing and could not be built ({e}); the CUDA extension is "
                    "mandatory, there is no CPU path") from e
    L = C.CDLL(path)
    L.rs_version.restype = C.c_int
    L.rs_status_string.restype = C.c_char_p
    L.rs_status_string.argtypes = [C.c_int]
    L.rs_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.rs_ctx_destroy.argtypes = [C.c_void_p]
    L.rs_ctx_sync_status.argtypes = [C.c_void_p, C.c_void_p]
    L.rs_ctx_last_cuda_error.argtypes = [C.c_void_p]
    L.rs_ctx_launch_count.argtypes = [C.c_void_p]
    L.rs_ctx_launch_count.restype = C.c_int64
    P = C.c_void_p
    L.rs_road_bbox_dev.argtypes = [P, C.POINTER(RsRoads), P, P]
    L.rs_zonal_hist_dev.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsTiles), C.POINTER(RsPairs),
                                    C.POINTER(RsZonalParams), P, P, P]
    L.rs_zonal_hist_host.argtypes = L.rs_zonal_hist_dev.argtypes[:-1]
    L.rs_zonal_stats_host.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsTiles), C.POINTER(RsPairs),
                                      C.POINTER(RsZonalParams), C.c_int32, C.c_int32, P, C.c_int32, P, P, P]
    L.rs_zonal_stats_stream_host.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsTiles), C.POINTER(RsPairs),
                                             C.POINTER(RsZonalParams), C.c_int32, C.c_int32, P, C.c_int32, C.c_int32, P, P, P]
    L.rs_zonal_stats_mapped_host.argtypes = L.rs_zonal_stats_host.argtypes
    L.rs_rasterize_pairs_dev.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsTiles), C.POINTER(RsPairs), C.c_int, P, P]
    L.rs_rasterize_pairs_host.argtypes = L.rs_rasterize_pairs_dev.argtypes[:-1]
    L.rs_finalize_stats_dev.argtypes = [P, P, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P, C.c_int32, P, P]
    L.rs_finalize_stats_host.argtypes = L.rs_finalize_stats_dev.argtypes[:-1]
    L.rs_vote_metrics_dev.argtypes = [P, P, P, C.c_int32, P, C.c_int32, C.c_int32, C.c_double, P, P, P, P, P]
    L.rs_vote_metrics_host.argtypes = L.rs_vote_metrics_dev.argtypes[:-1]
    L.rs_extract_pixels_host.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsTiles), C.POINTER(RsPairs), C.c_int, P, P,
                                         C.c_int64, C.POINTER(C.c_int64)]
    L.rs_band_ratio_columns.argtypes = [C.c_int32]
    L.rs_band_ratios_dev.argtypes = [P, P, C.c_int64, C.c_int32, P, P]
    L.rs_band_ratios_host.argtypes = [P, P, C.c_int64, C.c_int32, P]
    L.rs_bin_counts_host.argtypes = [P, P, P, P, P, C.c_int32, C.c_int32, C.c_int32, P, P, C.c_int32, P]
    L.rs_within_host.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsRoads), P]
    L.rs_assemble_tiles_dev.argtypes = [P, P] + [C.c_int32] * 9 + [P, C.c_int32, P, P, P, P]
    L.rs_assemble_tiles_host.argtypes = L.rs_assemble_tiles_dev.argtypes[:-1]
    L.rs_host_register.argtypes = [P, P, C.c_size_t]
    L.rs_host_unregister.argtypes = [P, P]
    L.rs_overlay_area_host.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsRoads), P, P, C.c_int32, P, P]
    L.rs_group_hist_host.argtypes = [P, P, P, C.c_int64, C.c_int32, P]
    L.rs_vote_table_host.argtypes = [P, P, P, P, P, P, C.c_int32, P, C.c_int32, P, P]
    L.rs_confusion_metrics_host.argtypes = [P, P, P, C.c_int32, C.c_int32, P, P]
    L.rs_pairs_bbox_host.argtypes = [P, P, C.c_int32, P, C.c_int32, C.POINTER(RsLattice), P, P, C.c_int64, C.POINTER(C.c_int64)]
    L.rs_ks_hist_host.argtypes = [P, P, P, P, C.c_int32, C.c_int32, P, P]
    L.rs_rescale_u16_dev.argtypes = [P, P, C.c_int64, C.c_int32, C.c_int32, P, P, P, C.c_int32, P, P]
    L.rs_rescale_u16_host.argtypes = L.rs_rescale_u16_dev.argtypes[:-1]
    L.rs_zonal_stats_f32_host.argtypes = [P, C.POINTER(RsRoads), P, C.c_int32, C.c_int32, P, C.c_int32, C.c_double, C.c_int32, P,
                                          C.c_int32, P]
    L.rs_pairs_bbox_grid_host.argtypes = [P, P, C.c_int32, P, C.c_int32, P, P, C.c_int64, C.POINTER(C.c_int64)]
    L.rs_pairs_intersect_host.argtypes = [P, C.POINTER(RsRoads), P, C.c_int32, P, P, C.c_int32, P]
    L.rs_clip_rings_host.argtypes = [P, C.POINTER(RsRoads), P, P, C.c_int32, P, P, P, P]
    L.rs_decode_segments_dev.argtypes = [P, P, P, C.c_int32, C.c_int32, P, P, P]
    L.rs_decode_segments_host.argtypes = L.rs_decode_segments_dev.argtypes[:-1]
    L.rs_ingest_tiles_host.argtypes = [P, P, P, C.c_int32, C.c_int32, P] + [C.c_int32] * 9 + [P, C.c_int32, P, P, P, C.c_int32, P]
    L.rs_zonal_stats_compressed_host.argtypes = [P, C.POINTER(RsRoads), C.POINTER(RsTiles), C.POINTER(RsPairs), C.POINTER(RsZonalParams),
                                                 C.c_int32, C.c_int32, P, C.c_int32, P, P, C.c_int32, C.c_int32, P, C.c_int32, C.c_int32,
                                                 C.c_int32, P]
    L.rs_comm_unique_id.argtypes = [P]
    L.rs_comm_init.argtypes = [P, P, C.c_int32, C.c_int32]
    L.rs_comm_destroy.argtypes = [P]
    L.rs_comm_world.argtypes = [P]
    L.rs_allreduce_accumulators_dev.argtypes = [P, P, C.c_int64, P, P, C.c_int64, P]
    L.rs_synth_tiles_dev.argtypes = [P, P, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_uint64, P]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("rs_status_string", "rs_ctx_launch_count"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(status: int, where: str, ctx=None):
    if status == RS_OK:
        return
    detail = ""
    if status == -2 and ctx is not None:
        detail = f"cudaError {load().rs_ctx_last_cuda_error(ctx)}"
    raise NativeError(status, where, detail)
