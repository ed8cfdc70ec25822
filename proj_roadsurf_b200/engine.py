"""Batched overlay engine: the Python face of the C ABI (include/roadsurf_b200.h).

Two families of calls:
  * ``*_host``  numpy arrays in, numpy arrays out; host<->device copies happen inside the C call
                (what the reference-shaped helpers in functions/ and road_segmentation/ use);
  * ``*_dev``   torch CUDA tensors in, torch CUDA tensors out, asynchronous on the current torch
                stream (inputs resident in HBM; what bench.py times as ``value``).
torch is plumbing here (device memory, streams); every computation is a hand-written kernel in
libroadsurf_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .geometry import PairList, RoadSet, TileBatch

_HIST_MODES = {"bands": N.RS_HIST_BANDS, "class_score": N.RS_HIST_CLASS_SCORE}
_WINDOW_MODES = {"crop": N.RS_WINDOW_CROP, "full": N.RS_WINDOW_FULL, "boundless": N.RS_WINDOW_BOUNDLESS}
_NODATA_MODES = {"raw": N.RS_NODATA_RAW, "none": N.RS_NODATA_NONE, "zero": N.RS_NODATA_ZERO,
                 "zero_masked": N.RS_NODATA_ZERO_MASKED}
_RULES = {"count": N.RS_VOTE_COUNT, "score": N.RS_VOTE_SCORE}


def scale_params(smin: Sequence[float], smax: Sequence[float], f32: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """k, off of gdal.Translate scaleParams [smin, smax, 0, 255] (tif2cog.py:260-270), evaluated in
    the working precision GDAL is assumed to use (float64 by default)."""
    ft = np.float32 if f32 else np.float64
    lo, hi = np.asarray(smin, ft), np.asarray(smax, ft)
    k = ft(255.0) / (hi - lo)
    off = ft(0.0) - lo * k
    return k.astype(np.float64), off.astype(np.float64)


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


@dataclass
class DeviceRoads:
    xy: "object"
    ring_off: "object"
    road_ring_off: "object"
    bbox: "object"
    n_roads: int
    n_rings: int
    n_verts: int


@dataclass
class DevicePairs:
    road_pair_off: "object"
    pair_tile: "object"
    n_pairs: int


@dataclass
class DeviceTiles:
    pixels: "object"       # (T, H, W, C) uint8 | int16-viewed uint16
    gt: "object"
    n_tiles: int
    height: int
    width: int
    channels: int
    dtype: int


class Engine:
    """One context per device (rs_ctx_create)."""

    def __init__(self, device: int = 0):
        self.lib = N.load()
        self.device = int(device)
        h = C.c_void_p()
        N.check(self.lib.rs_ctx_create(self.device, C.byref(h)), "rs_ctx_create")
        self._ctx = h

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.rs_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.rs_ctx_launch_count(self._ctx))

    # ------------------------------------------------------------------ descriptors
    @staticmethod
    def _roads_desc(xy, ring_off, road_ring_off, bbox, n_roads, n_rings, n_verts) -> N.RsRoads:
        return N.RsRoads(xy, ring_off, road_ring_off, bbox, n_roads, n_rings, n_verts)

    @staticmethod
    def _params(hist_mode, window, rescale, road_slot_ptr, border_px: int = 0, min_zero_ptr=None) -> N.RsZonalParams:
        p = N.RsZonalParams()
        p.hist_mode = _HIST_MODES[hist_mode]
        p.window_mode = _WINDOW_MODES[window]
        p.border_px = int(border_px)
        p.rescale = 0
        if rescale is not None:
            k, off, f32 = rescale
            p.rescale = 2 if f32 else 1
            for i in range(4):
                p.scale_k[i] = float(k[i]) if i < len(k) else 1.0
                p.scale_off[i] = float(off[i]) if i < len(off) else 0.0
        p.road_slot = road_slot_ptr
        p.min_zero = min_zero_ptr
        return p

    # ------------------------------------------------------------------ host family
    def zonal_hist_host(self, roads: RoadSet, tiles: TileBatch, pairs: PairList, hist_mode: str = "bands",
                        window: str = "crop", rescale=None, road_slot: Optional[np.ndarray] = None,
                        n_slots: Optional[int] = None, border_px: int = 0, want_min_zero: bool = False):
        """Per-road histograms (n_slots, HC, 256) uint32 and all-zero-pixel counts (n_slots,) uint32; with want_min_zero
        also the (n_slots,) uint32 zero-padding term of the nodata == 0 convention (rs_zonal_params::min_zero)."""
        px = np.ascontiguousarray(tiles.pixels)
        dtype = N.RS_U16 if px.dtype == np.uint16 else N.RS_U8
        R = roads.n_roads
        HC = 3 if hist_mode == "class_score" else tiles.channels
        slot = None if road_slot is None else np.ascontiguousarray(road_slot, np.int32)
        S = R if slot is None else (int(slot.max()) + 1 if R else 0)     # rows the C call writes back
        if n_slots is not None and slot is not None:
            assert n_slots >= S, "n_slots smaller than the largest slot"
        S_out = S if (n_slots is None or slot is None) else int(n_slots)
        hist = np.zeros((S_out, HC, 256), np.uint32)
        nzero = np.zeros(S_out, np.uint32)
        xy = np.ascontiguousarray(roads.xy, np.float64)
        ro, rro = np.ascontiguousarray(roads.ring_off, np.int32), np.ascontiguousarray(roads.road_ring_off, np.int32)
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        rpo, pt = np.ascontiguousarray(pairs.road_pair_off, np.int32), np.ascontiguousarray(pairs.pair_tile, np.int32)
        gt = np.ascontiguousarray(tiles.gt, np.float64)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), R, roads.n_rings, roads.n_verts)
        td = N.RsTiles(_np_ptr(px), _np_ptr(gt), tiles.n_tiles, tiles.height, tiles.width, tiles.channels, dtype)
        pd_ = N.RsPairs(_np_ptr(rpo), _np_ptr(pt), pairs.n_pairs)
        minz = np.zeros(S_out, np.uint32) if want_min_zero else None
        prm = self._params(hist_mode, window, rescale, _np_ptr(slot), border_px, _np_ptr(minz))
        st = self.lib.rs_zonal_hist_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), C.byref(prm),
                                         _np_ptr(hist), _np_ptr(nzero))
        N.check(st, "rs_zonal_hist_host", self._ctx)
        return (hist, nzero, minz) if want_min_zero else (hist, nzero)

    def zonal_stats_host(self, roads: RoadSet, tiles: TileBatch, pairs: PairList, window: str = "crop", rescale=None,
                         nodata_mode: str = "none", ddof: int = 1, percentiles: Sequence[float] = (),
                         want_hist: bool = False, tiles_per_chunk: Optional[int] = None,
                         mapped: Optional[bool] = None):
        """Host buffers in, statistics table out, in ONE C call (rs_zonal_stats_host): the batched form of
        statistical_analysis.py:179-246.  Returns stats (R, C, RS_NSTAT + n_pct) [, hist, n_allzero].
        mapped: True = the kernel reads the page-locked tiles in place (rs_zonal_stats_mapped_host; RS_ERR_NOT_PINNED for
        pageable memory), False = the tiles are copied (whole, or streamed in chunks of tiles_per_chunk), None = in place
        when the buffer is page-locked (``pin_host`` / torch pin_memory), copied otherwise."""
        px = np.ascontiguousarray(tiles.pixels) if isinstance(tiles.pixels, np.ndarray) else tiles.pixels
        dtype = N.RS_U16 if px.dtype == np.uint16 else N.RS_U8
        R, Cc = roads.n_roads, tiles.channels
        pct = np.ascontiguousarray(percentiles, np.float64)
        stats = np.zeros((R, Cc, N.RS_NSTAT + len(pct)), np.float64)
        hist = np.zeros((R, Cc, 256), np.uint32) if want_hist else None
        nzero = np.zeros(R, np.uint32) if want_hist else None
        xy = np.ascontiguousarray(roads.xy, np.float64)
        ro, rro = np.ascontiguousarray(roads.ring_off, np.int32), np.ascontiguousarray(roads.road_ring_off, np.int32)
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        rpo, pt = np.ascontiguousarray(pairs.road_pair_off, np.int32), np.ascontiguousarray(pairs.pair_tile, np.int32)
        gt = np.ascontiguousarray(tiles.gt, np.float64)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), R, roads.n_rings, roads.n_verts)
        td = N.RsTiles(_np_ptr(px), _np_ptr(gt), tiles.n_tiles, tiles.height, tiles.width, tiles.channels, dtype)
        pd_ = N.RsPairs(_np_ptr(rpo), _np_ptr(pt), pairs.n_pairs)
        prm = self._params("bands", window, rescale, None)
        st = None
        if mapped or (mapped is None and tiles_per_chunk is None):
            # page-locked tiles read in place by the kernel (only the sectors under road pixels cross the host link)
            st = self.lib.rs_zonal_stats_mapped_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), C.byref(prm),
                                                     _NODATA_MODES[nodata_mode], int(ddof), _np_ptr(pct) if len(pct) else None,
                                                     len(pct), _np_ptr(stats), _np_ptr(hist), _np_ptr(nzero))
            if st == N.RS_ERR_NOT_PINNED and mapped is None:
                st = None                                   # pageable tiles: copy them
        if st is not None:
            pass
        elif tiles_per_chunk is None:
            st = self.lib.rs_zonal_stats_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), C.byref(prm),
                                              _NODATA_MODES[nodata_mode], int(ddof), _np_ptr(pct) if len(pct) else None,
                                              len(pct), _np_ptr(stats), _np_ptr(hist), _np_ptr(nzero))
        else:       # tiles streamed from host memory through two device buffers (copy overlaps compute)
            st = self.lib.rs_zonal_stats_stream_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), C.byref(prm),
                                                     _NODATA_MODES[nodata_mode], int(ddof), _np_ptr(pct) if len(pct) else None,
                                                     len(pct), int(tiles_per_chunk), _np_ptr(stats), _np_ptr(hist), _np_ptr(nzero))
        N.check(st, "rs_zonal_stats_host", self._ctx)
        return (stats, hist, nzero) if want_hist else stats

    def zonal_stats_f32_host(self, features: RoadSet, raster: np.ndarray, affine, nodata: Optional[float] = None, ddof: int = 0,
                             percentiles: Sequence[float] = ()) -> np.ndarray:
        """Statistics of every feature over ONE float32 raster (H, W) (rs_zonal_stats_f32_host: rasterstats.zonal_stats over a
        DEM, fct_rasters.py:147-163).  NaN and ``nodata`` pixels are masked.  Returns (n, RS_NSTAT + n_pct) float64."""
        arr = np.ascontiguousarray(raster, np.float32)
        assert arr.ndim == 2
        pct = np.ascontiguousarray(percentiles, np.float64)
        gt = np.ascontiguousarray(tuple(affine)[:6], np.float64)
        xy = np.ascontiguousarray(features.xy, np.float64)
        ro, rro = np.ascontiguousarray(features.ring_off, np.int32), np.ascontiguousarray(features.road_ring_off, np.int32)
        bb = np.ascontiguousarray(features.bbox, np.float64)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), features.n_roads, features.n_rings, features.n_verts)
        out = np.zeros((features.n_roads, N.RS_NSTAT + len(pct)), np.float64)
        st = self.lib.rs_zonal_stats_f32_host(self._ctx, C.byref(rd), _np_ptr(arr), arr.shape[0], arr.shape[1], _np_ptr(gt),
                                              0 if nodata is None else 1, 0.0 if nodata is None else float(nodata), int(ddof),
                                              _np_ptr(pct) if len(pct) else None, len(pct), _np_ptr(out))
        N.check(st, "rs_zonal_stats_f32_host", self._ctx)
        return out

    def zonal_stats_compressed_host(self, roads: RoadSet, gt: np.ndarray, height: int, width: int, channels: int, pairs: PairList,
                                    comp: np.ndarray, comp_off, raw_off, codec: int = 8, planar: int = 1, predictor: int = 1,
                                    nodata_mode: str = "none", ddof: int = 1, percentiles: Sequence[float] = ()) -> np.ndarray:
        """Statistics straight from compressed uint8 tile segments (rs_zonal_stats_compressed_host): decode + assemble + fused
        rasterize / histogram + statistics on the device, only the compressed bytes uploaded."""
        comp = np.ascontiguousarray(comp, np.uint8)
        co, ro = np.ascontiguousarray(comp_off, np.int64), np.ascontiguousarray(raw_off, np.int64)
        gt = np.ascontiguousarray(gt, np.float64).reshape(-1, 6)
        pct = np.ascontiguousarray(percentiles, np.float64)
        R = roads.n_roads
        stats = np.zeros((R, channels, N.RS_NSTAT + len(pct)), np.float64)
        xy = np.ascontiguousarray(roads.xy, np.float64)
        rof, rro = np.ascontiguousarray(roads.ring_off, np.int32), np.ascontiguousarray(roads.road_ring_off, np.int32)
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        rpo, pt = np.ascontiguousarray(pairs.road_pair_off, np.int32), np.ascontiguousarray(pairs.pair_tile, np.int32)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(rof), _np_ptr(rro), _np_ptr(bb), R, roads.n_rings, roads.n_verts)
        td = N.RsTiles(None, _np_ptr(gt), gt.shape[0], height, width, channels, N.RS_U8)
        pd_ = N.RsPairs(_np_ptr(rpo), _np_ptr(pt), pairs.n_pairs)
        prm = self._params("bands", "crop", None, None)
        st = self.lib.rs_zonal_stats_compressed_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), C.byref(prm),
                                                     _NODATA_MODES[nodata_mode], int(ddof), _np_ptr(pct) if len(pct) else None, len(pct),
                                                     _np_ptr(comp), _np_ptr(co), len(co) - 1, int(codec), _np_ptr(ro), int(planar),
                                                     int(predictor), 0, _np_ptr(stats))
        N.check(st, "rs_zonal_stats_compressed_host", self._ctx)
        return stats

    def pin_host(self, array: np.ndarray) -> np.ndarray:
        """Page-lock a numpy buffer in place (rs_host_register) so that zonal_stats_host reads it without a copy; call
        unpin_host before the array is freed.  Returns the array."""
        st = self.lib.rs_host_register(self._ctx, array.ctypes.data, array.nbytes)
        N.check(st, "rs_host_register", self._ctx)
        return array

    def unpin_host(self, array: np.ndarray) -> None:
        N.check(self.lib.rs_host_unregister(self._ctx, array.ctypes.data), "rs_host_unregister", self._ctx)

    def rasterize_pairs_host(self, roads: RoadSet, gt: np.ndarray, height: int, width: int, pairs: PairList,
                             window: str = "crop") -> np.ndarray:
        """uint8 masks (n_pairs, H, W), 1 = pixel selected for that (road, tile) pair."""
        masks = np.zeros((pairs.n_pairs, height, width), np.uint8)
        if pairs.n_pairs == 0:
            return masks
        xy = np.ascontiguousarray(roads.xy, np.float64)
        ro, rro = np.ascontiguousarray(roads.ring_off, np.int32), np.ascontiguousarray(roads.road_ring_off, np.int32)
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        rpo, pt = np.ascontiguousarray(pairs.road_pair_off, np.int32), np.ascontiguousarray(pairs.pair_tile, np.int32)
        gt = np.ascontiguousarray(gt, np.float64).reshape(-1, 6)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), roads.n_roads, roads.n_rings, roads.n_verts)
        td = N.RsTiles(None, _np_ptr(gt), gt.shape[0], height, width, 1, N.RS_U8)
        pd_ = N.RsPairs(_np_ptr(rpo), _np_ptr(pt), pairs.n_pairs)
        st = self.lib.rs_rasterize_pairs_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), _WINDOW_MODES[window],
                                              _np_ptr(masks))
        N.check(st, "rs_rasterize_pairs_host", self._ctx)
        return masks

    def finalize_stats_host(self, hist: np.ndarray, n_allzero: Optional[np.ndarray], nodata_mode: str = "none",
                            ddof: int = 1, percentiles: Sequence[float] = ()) -> np.ndarray:
        """(R, C, RS_NSTAT + len(percentiles)) float64; columns _native.STAT_COLS then the percentiles."""
        hist = np.ascontiguousarray(hist, np.uint32)
        R, Cc = hist.shape[0], hist.shape[1]
        nz = None if n_allzero is None else np.ascontiguousarray(n_allzero, np.uint32)
        pct = np.ascontiguousarray(percentiles, np.float64)
        out = np.zeros((R, Cc, N.RS_NSTAT + len(pct)), np.float64)
        st = self.lib.rs_finalize_stats_host(self._ctx, _np_ptr(hist), _np_ptr(nz), R, Cc, _NODATA_MODES[nodata_mode],
                                             int(ddof), _np_ptr(pct) if len(pct) else None, len(pct), _np_ptr(out))
        N.check(st, "rs_finalize_stats_host", self._ctx)
        return out

    def vote_metrics_host(self, joint_hist: np.ndarray, gt_class: Optional[np.ndarray], cutoffs: Sequence[int],
                          rule: str = "count", min_area_frac: float = 0.0):
        """cover (n_thr, R) int8, scores (n_thr, R, 3), confusion (n_thr, 2, 4) int64, metrics (n_thr, 12)."""
        jh = np.ascontiguousarray(joint_hist, np.uint32)
        R = jh.shape[0]
        gtc = None if gt_class is None else np.ascontiguousarray(gt_class, np.int8)
        cut = np.ascontiguousarray(cutoffs, np.int32)
        T = len(cut)
        cover = np.zeros((T, R), np.int8)
        scores = np.zeros((T, R, 3), np.float64)
        conf = np.zeros((T, 2, 4), np.int64)
        met = np.zeros((T, N.RS_NMETRIC), np.float64)
        st = self.lib.rs_vote_metrics_host(self._ctx, _np_ptr(jh), _np_ptr(gtc), R, _np_ptr(cut), T, _RULES[rule],
                                           float(min_area_frac), _np_ptr(cover), _np_ptr(scores), _np_ptr(conf), _np_ptr(met))
        N.check(st, "rs_vote_metrics_host", self._ctx)
        return cover, scores, conf, met

    def extract_pixels_host(self, roads: RoadSet, tiles: TileBatch, pairs: PairList, window: str = "crop"):
        """In-mask pixels of every pair, row-major inside a pair (rs_extract_pixels_host).
        Returns pair_off (n_pairs + 1,) int64 and values (total, C) of the tile dtype."""
        px = np.ascontiguousarray(tiles.pixels)
        dtype = N.RS_U16 if px.dtype == np.uint16 else N.RS_U8
        xy = np.ascontiguousarray(roads.xy, np.float64)
        ro, rro = np.ascontiguousarray(roads.ring_off, np.int32), np.ascontiguousarray(roads.road_ring_off, np.int32)
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        rpo, pt = np.ascontiguousarray(pairs.road_pair_off, np.int32), np.ascontiguousarray(pairs.pair_tile, np.int32)
        gt = np.ascontiguousarray(tiles.gt, np.float64)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), roads.n_roads, roads.n_rings, roads.n_verts)
        td = N.RsTiles(_np_ptr(px), _np_ptr(gt), tiles.n_tiles, tiles.height, tiles.width, tiles.channels, dtype)
        pd_ = N.RsPairs(_np_ptr(rpo), _np_ptr(pt), pairs.n_pairs)
        pair_off = np.zeros(pairs.n_pairs + 1, np.int64)
        total = C.c_int64(0)
        st = self.lib.rs_extract_pixels_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), _WINDOW_MODES[window],
                                             _np_ptr(pair_off), None, 0, C.byref(total))
        N.check(st, "rs_extract_pixels_host", self._ctx)
        values = np.zeros((int(total.value), tiles.channels), px.dtype)
        if total.value:
            st = self.lib.rs_extract_pixels_host(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), _WINDOW_MODES[window],
                                                 _np_ptr(pair_off), _np_ptr(values), int(total.value), C.byref(total))
            N.check(st, "rs_extract_pixels_host", self._ctx)
        return pair_off, values

    def pairs_bbox_host(self, roads: RoadSet, tiles: TileBatch) -> PairList:
        """GPU broad phase (rs_pairs_bbox_host): every (road, tile) whose bounding boxes overlap, tiles on a
        regular lattice -- the pair list gpd.sjoin(tiles, roads) gives, as a superset."""
        from .geometry import lattice_of
        lat = lattice_of(tiles)
        if lat is None:
            raise ValueError("tiles are not on a regular lattice: build the pair list with geometry.pairs_by_bbox")
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        ext = np.ascontiguousarray(tiles.extents(), np.float64)
        lut = np.ascontiguousarray(lat.lut, np.int32)
        ld = N.RsLattice(lat.x0, lat.y0, lat.tile_w, lat.tile_h, lat.nx, lat.ny, _np_ptr(lut))
        R = roads.n_roads
        off = np.zeros(R + 1, np.int32)
        total = C.c_int64(0)
        st = self.lib.rs_pairs_bbox_host(self._ctx, _np_ptr(bb), R, _np_ptr(ext), tiles.n_tiles, C.byref(ld), _np_ptr(off), None, 0,
                                         C.byref(total))
        N.check(st, "rs_pairs_bbox_host", self._ctx)
        pt = np.zeros(int(total.value), np.int32)
        if total.value:
            st = self.lib.rs_pairs_bbox_host(self._ctx, _np_ptr(bb), R, _np_ptr(ext), tiles.n_tiles, C.byref(ld), _np_ptr(off),
                                             _np_ptr(pt), int(total.value), C.byref(total))
            N.check(st, "rs_pairs_bbox_host", self._ctx)
        return PairList(off, pt)

    def pairs_bbox_grid_host(self, roads: RoadSet, tiles) -> PairList:
        """GPU broad phase for tiles that are not on a lattice (rs_pairs_bbox_grid_host): uniform-grid binning on the device.
        ``tiles``: a TileBatch, or an (T, 4) array of extents (xmin, ymin, xmax, ymax)."""
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        ext = np.ascontiguousarray(tiles.extents() if hasattr(tiles, "extents") else tiles, np.float64).reshape(-1, 4)
        n_tiles = len(ext)
        R = roads.n_roads
        off = np.zeros(R + 1, np.int32)
        total = C.c_int64(0)
        st = self.lib.rs_pairs_bbox_grid_host(self._ctx, _np_ptr(bb), R, _np_ptr(ext), n_tiles, _np_ptr(off), None, 0, C.byref(total))
        N.check(st, "rs_pairs_bbox_grid_host", self._ctx)
        pt = np.zeros(int(total.value), np.int32)
        if total.value:
            st = self.lib.rs_pairs_bbox_grid_host(self._ctx, _np_ptr(bb), R, _np_ptr(ext), n_tiles, _np_ptr(off), _np_ptr(pt),
                                                  int(total.value), C.byref(total))
            N.check(st, "rs_pairs_bbox_grid_host", self._ctx)
        return PairList(off, pt)

    def pairs_intersect_host(self, roads: RoadSet, tiles, pairs: PairList) -> PairList:
        """Exact reject (rs_pairs_intersect_host): the pairs whose road polygon intersects the tile rectangle -- the pair table
        gpd.sjoin(tiles, roads) gives (statistical_analysis.py:170-171).  ``tiles``: a TileBatch or (T, 4) extents."""
        if pairs.n_pairs == 0:
            return pairs
        xy = np.ascontiguousarray(roads.xy, np.float64)
        ro, rro = np.ascontiguousarray(roads.ring_off, np.int32), np.ascontiguousarray(roads.road_ring_off, np.int32)
        bb = np.ascontiguousarray(roads.bbox, np.float64)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), roads.n_roads, roads.n_rings, roads.n_verts)
        ext = np.ascontiguousarray(tiles.extents() if hasattr(tiles, "extents") else tiles, np.float64).reshape(-1, 4)
        rpo, pt = np.ascontiguousarray(pairs.road_pair_off, np.int32), np.ascontiguousarray(pairs.pair_tile, np.int32)
        keep = np.zeros(pairs.n_pairs, np.uint8)
        st = self.lib.rs_pairs_intersect_host(self._ctx, C.byref(rd), _np_ptr(ext), len(ext), _np_ptr(rpo), _np_ptr(pt), pairs.n_pairs,
                                              _np_ptr(keep))
        N.check(st, "rs_pairs_intersect_host", self._ctx)
        k = keep.astype(bool)
        csum = np.concatenate([[0], np.cumsum(k)])
        return PairList(csum[rpo.astype(np.int64)].astype(np.int32), np.ascontiguousarray(pt[k]))

    def clip_rings_host(self, labels: RoadSet, pair_label, rect):
        """The rings of label pair_label[p] clipped to the rectangle rect[p] (rs_clip_rings_host, determine_class.clip_labels).
        Returns a RoadSet whose 'road' p is the clipped label of pair p (rings that miss the rectangle are dropped)."""
        pl = np.ascontiguousarray(pair_label, np.int32)
        rc_ = np.ascontiguousarray(rect, np.float64).reshape(-1, 4)
        P = len(pl)
        nr = np.diff(labels.road_ring_off).astype(np.int64)[pl] if P else np.zeros(0, np.int64)
        pro = np.zeros(P + 1, np.int64)
        pro[1:] = np.cumsum(nr)
        nq = int(pro[-1])
        if nq == 0:
            return RoadSet.from_arrays(np.zeros((0, 2)), np.zeros(1, np.int32), np.zeros(P + 1, np.int32))
        xy = np.ascontiguousarray(labels.xy, np.float64)
        ro, rro = np.ascontiguousarray(labels.ring_off, np.int32), np.ascontiguousarray(labels.road_ring_off, np.int32)
        bb = np.ascontiguousarray(labels.bbox, np.float64)
        rd = self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), labels.n_roads, labels.n_rings, labels.n_verts)
        cnt = np.zeros(nq, np.int32)
        st = self.lib.rs_clip_rings_host(self._ctx, C.byref(rd), _np_ptr(pl), _np_ptr(rc_), P, _np_ptr(pro), _np_ptr(cnt), None, None)
        N.check(st, "rs_clip_rings_host", self._ctx)
        voff = np.zeros(nq, np.int64)
        voff[1:] = np.cumsum(cnt.astype(np.int64))[:-1]
        total = int(cnt.sum())
        out = np.zeros((total, 2), np.float64)
        if total:
            st = self.lib.rs_clip_rings_host(self._ctx, C.byref(rd), _np_ptr(pl), _np_ptr(rc_), P, _np_ptr(pro), _np_ptr(cnt), _np_ptr(voff),
                                             _np_ptr(out))
            N.check(st, "rs_clip_rings_host", self._ctx)
        keep = cnt > 0                                                 # rings that survive, per pair
        kcs = np.concatenate([[0], np.cumsum(keep)])
        ring_off = np.concatenate([[0], np.cumsum(cnt[keep].astype(np.int64))])
        return RoadSet.from_arrays(out, ring_off.astype(np.int32), kcs[pro].astype(np.int32))

    def rescale_u16_host(self, src: np.ndarray, smin: Sequence[float], smax: Sequence[float], bidx: Optional[Sequence[int]] = None,
                         f32: bool = False) -> np.ndarray:
        """gdal.Translate -scale smin smax 0 255 -ot Byte per OUTPUT band (tif2cog.py:260-270): src (..., C_in) uint16 ->
        (..., C_out) uint8, output band c read from source band bidx[c]."""
        src = np.ascontiguousarray(src, np.uint16)
        c_in = src.shape[-1]
        k, off = scale_params(smin, smax, f32)
        c_out = len(k)
        bi = None if bidx is None else np.ascontiguousarray(bidx, np.int32)
        out = np.zeros(src.shape[:-1] + (c_out,), np.uint8)
        n = int(np.prod(src.shape[:-1]))
        st = self.lib.rs_rescale_u16_host(self._ctx, _np_ptr(src), n, c_in, c_out, _np_ptr(bi), _np_ptr(k), _np_ptr(off), int(f32), _np_ptr(out))
        N.check(st, "rs_rescale_u16_host", self._ctx)
        return out

    def rescale_u16_dev(self, src, smin, smax, bidx=None, f32: bool = False, out=None):
        """device form: src int16-viewed uint16 CUDA tensor (..., C_in) -> uint8 CUDA tensor (..., C_out), asynchronous"""
        torch = self._torch()
        c_in = int(src.shape[-1])
        k, off = scale_params(smin, smax, f32)
        c_out = len(k)
        bi = None if bidx is None else np.ascontiguousarray(bidx, np.int32)
        if out is None:
            out = torch.empty(tuple(src.shape[:-1]) + (c_out,), dtype=torch.uint8, device=src.device)
        n = int(src.numel() // c_in)
        st = self.lib.rs_rescale_u16_dev(self._ctx, src.data_ptr(), n, c_in, c_out, _np_ptr(bi), _np_ptr(k), _np_ptr(off), int(f32),
                                         out.data_ptr(), self._stream())
        N.check(st, "rs_rescale_u16_dev", self._ctx)
        return out

    def ks_hist_host(self, hist: np.ndarray, ref_hist: np.ndarray, ref_of_road: Optional[np.ndarray] = None):
        """KS D of every road's (R, 256) histogram against ref_hist (n_refs, 256) (rs_ks_hist_host).  Returns D (R,), n (R,)."""
        h = np.ascontiguousarray(hist, np.uint32)
        ref = np.ascontiguousarray(np.atleast_2d(ref_hist), np.uint64)
        ror = None if ref_of_road is None else np.ascontiguousarray(ref_of_road, np.int32)
        R = h.shape[0]
        D, n = np.zeros(R, np.float64), np.zeros(R, np.float64)
        st = self.lib.rs_ks_hist_host(self._ctx, _np_ptr(h), _np_ptr(ror), _np_ptr(ref), R, ref.shape[0], _np_ptr(D), _np_ptr(n))
        N.check(st, "rs_ks_hist_host", self._ctx)
        return D, n

    def group_hist_host(self, values: np.ndarray, group: np.ndarray, n_groups: int) -> np.ndarray:
        """(n_groups, 256) uint32 histograms of a uint8 column by group index (rs_group_hist_host)."""
        v = np.ascontiguousarray(values, np.uint8)
        g = np.ascontiguousarray(group, np.int32)
        assert v.shape == g.shape and v.ndim == 1
        hist = np.zeros((int(n_groups), 256), np.uint32)
        st = self.lib.rs_group_hist_host(self._ctx, _np_ptr(v), _np_ptr(g), int(v.shape[0]), int(n_groups), _np_ptr(hist))
        N.check(st, "rs_group_hist_host", self._ctx)
        return hist

    def band_ratios_host(self, values: np.ndarray) -> np.ndarray:
        """(K, n) float64 derived columns of a uint8 pixel table (n, C): band ratios rounded to 3 decimals (+ VgNIR-BI for
        C = 4), statistical_analysis.py:279-293 (rs_band_ratios_host)."""
        v = np.ascontiguousarray(values, np.uint8)
        assert v.ndim == 2
        K = int(self.lib.rs_band_ratio_columns(int(v.shape[1])))
        if K == 0:
            raise ValueError("band ratios need 2 to 4 bands")
        out = np.zeros((K, v.shape[0]), np.float64)
        st = self.lib.rs_band_ratios_host(self._ctx, _np_ptr(v), int(v.shape[0]), int(v.shape[1]), _np_ptr(out))
        N.check(st, "rs_band_ratios_host", self._ctx)
        return out

    def bin_counts_host(self, values, sel, hit, group, n_groups: int, lo, hi) -> np.ndarray:
        """(n_groups, K, T, 2) int64: rows per (group, column, bin lo < v <= hi) and the hits among them (rs_bin_counts_host)."""
        v = np.ascontiguousarray(np.atleast_2d(values), np.float64)
        s_ = np.ascontiguousarray(np.atleast_2d(sel), np.int8)
        h = np.ascontiguousarray(np.atleast_2d(hit), np.int8)
        g = np.ascontiguousarray(group, np.int32)
        lo, hi = np.ascontiguousarray(lo, np.float64), np.ascontiguousarray(hi, np.float64)
        K, n = v.shape
        assert s_.shape == v.shape and h.shape == v.shape and g.shape == (n,) and lo.shape == hi.shape
        out = np.zeros((int(n_groups), K, len(lo), 2), np.int64)
        st = self.lib.rs_bin_counts_host(self._ctx, _np_ptr(v), _np_ptr(s_), _np_ptr(h), _np_ptr(g), n, K, int(n_groups),
                                         _np_ptr(lo), _np_ptr(hi), len(lo), _np_ptr(out))
        N.check(st, "rs_bin_counts_host", self._ctx)
        return out

    def within_host(self, a: RoadSet, b: RoadSet) -> np.ndarray:
        """(Ra, Rb) uint8, 1 where polygon i of ``a`` lies within polygon j of ``b`` (rs_within_host): the predicate of
        gpd.sjoin(roads, buffered_quarries, predicate='within'), determine_class.py:57."""
        out = np.zeros((a.n_roads, b.n_roads), np.uint8)
        if out.size == 0:
            return out
        keep = []
        def desc(r):
            xy = np.ascontiguousarray(r.xy, np.float64)
            ro, rro = np.ascontiguousarray(r.ring_off, np.int32), np.ascontiguousarray(r.road_ring_off, np.int32)
            bb = np.ascontiguousarray(r.bbox, np.float64)
            keep.extend([xy, ro, rro, bb])
            return self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), r.n_roads, r.n_rings, r.n_verts)
        da, db = desc(a), desc(b)
        st = self.lib.rs_within_host(self._ctx, C.byref(da), C.byref(db), _np_ptr(out))
        N.check(st, "rs_within_host", self._ctx)
        return out

    def assemble_tiles_host(self, raw: np.ndarray, info, bidx=None, rescale: Optional[dict] = None) -> np.ndarray:
        """(T, H, W, C_out) uint8 | uint16 tiles from decompressed TIFF samples (rs_assemble_tiles_host): predictor, byte
        order, planar -> interleaved, band selection (``bidx`` 1-based) and optional 16 -> 8 bit rescale in one kernel.
        ``info``: ingest.TiffInfo (height, width, channels, sample_bytes, planar, predictor, big_endian)."""
        raw = np.ascontiguousarray(raw, np.uint8)
        T = raw.shape[0]
        H, W, Cin, sb = info.height, info.width, info.channels, info.sample_bytes
        assert raw.size == T * H * W * Cin * sb, "raw sample buffer does not match the tile layout"
        bi = np.arange(Cin, dtype=np.int32) if bidx is None else np.ascontiguousarray(bidx, np.int32) - 1
        Cout = len(bi)
        mode, k, off = 0, None, None
        if rescale is not None:
            k, off = scale_params(rescale["smin"], rescale["smax"], bool(rescale.get("f32")))
            assert len(k) == Cout, "one scale range per output band"
            mode = 2 if rescale.get("f32") else 1
        out = np.zeros((T, H, W, Cout), np.uint8 if (sb == 1 or mode) else np.uint16)
        st = self.lib.rs_assemble_tiles_host(self._ctx, _np_ptr(raw), T, H, W, Cin, int(info.planar), int(info.predictor), sb,
                                             int(bool(info.big_endian)), Cout, _np_ptr(bi), mode, _np_ptr(k), _np_ptr(off),
                                             _np_ptr(out))
        N.check(st, "rs_assemble_tiles_host", self._ctx)
        return out

    def decode_segments_host(self, comp: np.ndarray, comp_off, codec: int, raw_off) -> np.ndarray:
        """Decompress TIFF segments on the device (rs_decode_segments_host): comp uint8 (all segments concatenated), comp_off /
        raw_off int64 (n + 1,), codec = TIFF Compression tag (1, 5, 8, 32946).  Returns the raw bytes (raw_off[-1],) uint8."""
        comp = np.ascontiguousarray(comp, np.uint8)
        co, ro = np.ascontiguousarray(comp_off, np.int64), np.ascontiguousarray(raw_off, np.int64)
        n = len(co) - 1
        raw = np.zeros(int(ro[-1]), np.uint8)
        st = self.lib.rs_decode_segments_host(self._ctx, _np_ptr(comp), _np_ptr(co), n, int(codec), _np_ptr(raw), _np_ptr(ro))
        N.check(st, "rs_decode_segments_host", self._ctx)
        return raw

    def ingest_tiles_host(self, comp: np.ndarray, comp_off, codec: int, raw_off, n_tiles: int, info, bidx=None,
                          rescale: Optional[dict] = None) -> np.ndarray:
        """Compressed TIFF segments -> (T, H, W, C_out) tiles in one call (rs_ingest_tiles_host): decompression, predictor, byte
        order, band selection and rescale all on the device; only the compressed bytes are uploaded."""
        comp = np.ascontiguousarray(comp, np.uint8)
        co, ro = np.ascontiguousarray(comp_off, np.int64), np.ascontiguousarray(raw_off, np.int64)
        H, W, Cin, sb = info.height, info.width, info.channels, info.sample_bytes
        bi = np.arange(Cin, dtype=np.int32) if bidx is None else np.ascontiguousarray(bidx, np.int32) - 1
        Cout = len(bi)
        mode, k, off = 0, None, None
        if rescale is not None:
            k, off = scale_params(rescale["smin"], rescale["smax"], bool(rescale.get("f32")))
            assert len(k) == Cout, "one scale range per output band"
            mode = 2 if rescale.get("f32") else 1
        out = np.zeros((n_tiles, H, W, Cout), np.uint8 if (sb == 1 or mode) else np.uint16)
        st = self.lib.rs_ingest_tiles_host(self._ctx, _np_ptr(comp), _np_ptr(co), len(co) - 1, int(codec), _np_ptr(ro), n_tiles, H, W, Cin,
                                           int(info.planar), int(info.predictor), sb, int(bool(info.big_endian)), Cout, _np_ptr(bi),
                                           mode, _np_ptr(k), _np_ptr(off), _np_ptr(out), 0, None)
        N.check(st, "rs_ingest_tiles_host", self._ctx)
        return out

    def overlay_area_host(self, a: RoadSet, b: RoadSet, pair_a, pair_b):
        """(area of a[pair_a[k]] intersected with b[pair_b[k]] for every k, area of every polygon of ``a``):
        gpd.overlay(...).area and GeoSeries.area of determine_class.py:107-114 (rs_overlay_area_host)."""
        pa, pb = np.ascontiguousarray(pair_a, np.int32), np.ascontiguousarray(pair_b, np.int32)
        assert pa.shape == pb.shape and pa.ndim == 1
        out = np.zeros(len(pa), np.float64)
        area_a = np.zeros(a.n_roads, np.float64)
        keep = []
        def desc(r):
            xy = np.ascontiguousarray(r.xy, np.float64)
            ro, rro = np.ascontiguousarray(r.ring_off, np.int32), np.ascontiguousarray(r.road_ring_off, np.int32)
            bb = np.ascontiguousarray(r.bbox, np.float64)
            keep.extend([xy, ro, rro, bb])
            return self._roads_desc(_np_ptr(xy), _np_ptr(ro), _np_ptr(rro), _np_ptr(bb), r.n_roads, r.n_rings, r.n_verts)
        da, db = desc(a), desc(b)
        st = self.lib.rs_overlay_area_host(self._ctx, C.byref(da), C.byref(db), _np_ptr(pa), _np_ptr(pb), len(pa), _np_ptr(out),
                                           _np_ptr(area_a))
        N.check(st, "rs_overlay_area_host", self._ctx)
        return out, area_a

    def vote_table_host(self, row_off, cls, score, weighted, area, thresholds):
        """determine_detected_class on a detection table sorted by road (rs_vote_table_host).
        Returns cover (T, R) int8 and scores (T, R, 3) = artificial index, natural index, diff."""
        row_off = np.ascontiguousarray(row_off, np.int32)
        R = len(row_off) - 1
        cls = np.ascontiguousarray(cls, np.int8)
        score, weighted, area = (np.ascontiguousarray(a, np.float64) for a in (score, weighted, area))
        thr = np.ascontiguousarray(thresholds, np.float64)
        T = len(thr)
        cover = np.zeros((T, R), np.int8)
        scores = np.zeros((T, R, 3), np.float64)
        st = self.lib.rs_vote_table_host(self._ctx, _np_ptr(row_off), _np_ptr(cls), _np_ptr(score), _np_ptr(weighted),
                                         _np_ptr(area), R, _np_ptr(thr), T, _np_ptr(cover), _np_ptr(scores))
        N.check(st, "rs_vote_table_host", self._ctx)
        return cover, scores

    def confusion_metrics_host(self, cover: np.ndarray, gt_class: np.ndarray):
        """cover (T, R) int8 codes vs gt_class (R,) -> confusion (T, 2, 4) int64, metrics (T, 12)."""
        cover = np.ascontiguousarray(np.atleast_2d(cover), np.int8)
        gt = np.ascontiguousarray(gt_class, np.int8)
        T, R = cover.shape
        conf = np.zeros((T, 2, 4), np.int64)
        met = np.zeros((T, N.RS_NMETRIC), np.float64)
        st = self.lib.rs_confusion_metrics_host(self._ctx, _np_ptr(cover), _np_ptr(gt), R, T, _np_ptr(conf), _np_ptr(met))
        N.check(st, "rs_confusion_metrics_host", self._ctx)
        return conf, met

    # ------------------------------------------------------------------ device family (torch tensors)
    def _torch(self):
        import torch
        return torch

    def _stream(self):
        torch = self._torch()
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def upload_roads(self, roads: RoadSet) -> DeviceRoads:
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        return DeviceRoads(torch.from_numpy(np.ascontiguousarray(roads.xy, np.float64)).to(dev),
                           torch.from_numpy(np.ascontiguousarray(roads.ring_off, np.int32)).to(dev),
                           torch.from_numpy(np.ascontiguousarray(roads.road_ring_off, np.int32)).to(dev),
                           torch.from_numpy(np.ascontiguousarray(roads.bbox, np.float64)).to(dev),
                           roads.n_roads, roads.n_rings, roads.n_verts)

    def upload_pairs(self, pairs: PairList) -> DevicePairs:
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        return DevicePairs(torch.from_numpy(np.ascontiguousarray(pairs.road_pair_off, np.int32)).to(dev),
                           torch.from_numpy(np.ascontiguousarray(pairs.pair_tile, np.int32)).to(dev), pairs.n_pairs)

    def upload_tiles(self, tiles: TileBatch) -> DeviceTiles:
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        px = np.ascontiguousarray(tiles.pixels)
        dtype = N.RS_U16 if px.dtype == np.uint16 else N.RS_U8
        t = torch.from_numpy(px.view(np.int16) if dtype == N.RS_U16 else px).to(dev)
        return DeviceTiles(t, torch.from_numpy(np.ascontiguousarray(tiles.gt, np.float64)).to(dev), tiles.n_tiles,
                           tiles.height, tiles.width, tiles.channels, dtype)

    def synth_tiles_dev(self, tile_key, height: int, width: int, channels: int, dtype: str = "u8", kind: int = 0,
                        seed: int = 20261018, gt=None) -> DeviceTiles:
        """Generate synthetic tiles on the device (rs_synth_tiles_dev).  tile_key: int64 per tile."""
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        key = torch.as_tensor(np.ascontiguousarray(tile_key, np.int64)).to(dev)
        T = int(key.numel())
        code = N.RS_U16 if dtype == "u16" else N.RS_U8
        px = torch.empty((T, height, width, channels), dtype=torch.int16 if code == N.RS_U16 else torch.uint8, device=dev)
        st = self.lib.rs_synth_tiles_dev(self._ctx, px.data_ptr(), key.data_ptr(), T, height, width, channels, code, kind,
                                         C.c_uint64(seed), self._stream())
        N.check(st, "rs_synth_tiles_dev", self._ctx)
        g = None if gt is None else torch.from_numpy(np.ascontiguousarray(gt, np.float64)).to(dev)
        return DeviceTiles(px, g, T, height, width, channels, code)

    def zonal_hist_dev(self, roads: DeviceRoads, tiles: DeviceTiles, pairs: DevicePairs, hist_mode: str = "bands",
                       window: str = "crop", rescale=None, road_slot=None, out=None, check: bool = True,
                       n_slots: Optional[int] = None, min_zero=None):
        """Asynchronous on the current torch stream.  Returns (hist, n_allzero) CUDA tensors (int32 storage
        of the uint32 counters).  ``check`` synchronises and raises on a kernel-side failure.  With road_slot,
        n_slots is the number of output rows (rows no local road maps to stay zero; the rows of a shard's boundary
        table must exist on every rank); min_zero: optional (n_slots,) int32 CUDA tensor, filled with the zero-padding
        term of the nodata == 0 convention (rs_zonal_params::min_zero)."""
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        HC = 3 if hist_mode == "class_score" else tiles.channels
        if out is None:
            if road_slot is None:
                S = roads.n_roads
            else:
                S = int(road_slot.max().item()) + 1 if roads.n_roads else 0
                if n_slots is not None:
                    assert n_slots >= S, "n_slots smaller than the largest slot"
                    S = int(n_slots)
            hist = torch.zeros((S, HC, 256), dtype=torch.int32, device=dev) if road_slot is not None else \
                torch.empty((S, HC, 256), dtype=torch.int32, device=dev)
            nzero = torch.zeros((S,), dtype=torch.int32, device=dev)
        else:
            hist, nzero = out
        rd = self._roads_desc(roads.xy.data_ptr(), roads.ring_off.data_ptr(), roads.road_ring_off.data_ptr(),
                              roads.bbox.data_ptr(), roads.n_roads, roads.n_rings, roads.n_verts)
        td = N.RsTiles(tiles.pixels.data_ptr(), tiles.gt.data_ptr(), tiles.n_tiles, tiles.height, tiles.width,
                       tiles.channels, tiles.dtype)
        pd_ = N.RsPairs(pairs.road_pair_off.data_ptr(), pairs.pair_tile.data_ptr() if pairs.n_pairs else None, pairs.n_pairs)
        prm = self._params(hist_mode, window, rescale, None if road_slot is None else road_slot.data_ptr(),
                           min_zero_ptr=None if min_zero is None else min_zero.data_ptr())
        st = self.lib.rs_zonal_hist_dev(self._ctx, C.byref(rd), C.byref(td), C.byref(pd_), C.byref(prm),
                                        hist.data_ptr(), nzero.data_ptr(), self._stream())
        N.check(st, "rs_zonal_hist_dev", self._ctx)
        if check:
            self.sync_status()
        return hist, nzero

    def finalize_stats_dev(self, hist, n_allzero, nodata_mode: str = "none", ddof: int = 1,
                           percentiles: Sequence[float] = (), out=None, check: bool = True):
        torch = self._torch()
        R, Cc = int(hist.shape[0]), int(hist.shape[1])
        pct = np.ascontiguousarray(percentiles, np.float64)
        if out is None:
            out = torch.empty((R, Cc, N.RS_NSTAT + len(pct)), dtype=torch.float64, device=hist.device)
        assert tuple(out.shape) == (R, Cc, N.RS_NSTAT + len(pct)) and out.dtype == torch.float64 and out.is_contiguous()
        st = self.lib.rs_finalize_stats_dev(self._ctx, hist.data_ptr(), None if n_allzero is None else n_allzero.data_ptr(),
                                            R, Cc, _NODATA_MODES[nodata_mode], int(ddof),
                                            _np_ptr(pct) if len(pct) else None, len(pct), out.data_ptr(), self._stream())
        N.check(st, "rs_finalize_stats_dev", self._ctx)
        if check:
            self.sync_status()
        return out

    def vote_metrics_dev(self, joint_hist, gt_class, cutoffs: Sequence[int], rule: str = "count",
                         min_area_frac: float = 0.0, want_scores: bool = True, check: bool = True):
        torch = self._torch()
        dev = joint_hist.device
        R = int(joint_hist.shape[0])
        cut = np.ascontiguousarray(cutoffs, np.int32)
        T = len(cut)
        cover = torch.empty((T, R), dtype=torch.int8, device=dev)
        scores = torch.empty((T, R, 3), dtype=torch.float64, device=dev) if want_scores else None
        conf = torch.empty((T, 2, 4), dtype=torch.int64, device=dev)
        met = torch.empty((T, N.RS_NMETRIC), dtype=torch.float64, device=dev)
        st = self.lib.rs_vote_metrics_dev(self._ctx, joint_hist.data_ptr(), None if gt_class is None else gt_class.data_ptr(),
                                          R, _np_ptr(cut), T, _RULES[rule], float(min_area_frac), cover.data_ptr(),
                                          None if scores is None else scores.data_ptr(), conf.data_ptr(), met.data_ptr(),
                                          self._stream())
        N.check(st, "rs_vote_metrics_dev", self._ctx)
        if check:
            self.sync_status()
        return cover, scores, conf, met

    # ------------------------------------------------------------------ multi-GPU merge (rs_comm.cu)
    def comm_unique_id(self) -> bytes:
        """rank 0: the 128-byte NCCL rendezvous id to hand to every rank (rs_comm_unique_id)."""
        buf = C.create_string_buffer(N.RS_COMM_ID_BYTES)
        N.check(self.lib.rs_comm_unique_id(buf), "rs_comm_unique_id", self._ctx)
        return buf.raw

    def comm_init(self, unique_id: bytes, world: int, rank: int) -> None:
        """collective: build this context's NCCL communicator (rs_comm_init)."""
        assert len(unique_id) == N.RS_COMM_ID_BYTES
        N.check(self.lib.rs_comm_init(self._ctx, C.c_char_p(unique_id), int(world), int(rank)), "rs_comm_init", self._ctx)

    def comm_init_from_torch(self) -> None:
        """comm_init with the id passed through the torch.distributed process group that is already up."""
        torch = self._torch()
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", self.device)
        t = torch.zeros(N.RS_COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        self.comm_init(bytes(t.cpu().numpy().tobytes()), world, rank)

    @property
    def comm_world(self) -> int:
        return int(self.lib.rs_comm_world(self._ctx))

    def allreduce_accumulators_dev(self, hist, n_allzero, n_own: int, min_zero=None) -> None:
        """In-place sum over ranks of the boundary rows [n_own:] of a shard's tables, one grouped NCCL launch on the
        current torch stream (rs_allreduce_accumulators_dev)."""
        n_b = int(hist.shape[0]) - int(n_own)
        if n_b <= 0:
            return
        assert hist.is_contiguous() and n_allzero.is_contiguous()
        row = int(hist.shape[1]) * int(hist.shape[2])
        st = self.lib.rs_allreduce_accumulators_dev(
            self._ctx, hist.data_ptr() + 4 * row * int(n_own), n_b * row, n_allzero.data_ptr() + 4 * int(n_own),
            None if min_zero is None else min_zero.data_ptr() + 4 * int(n_own), n_b, self._stream())
        N.check(st, "rs_allreduce_accumulators_dev", self._ctx)

    def sync_status(self):
        st = self.lib.rs_ctx_sync_status(self._ctx, self._stream())
        N.check(st, "kernel status", self._ctx)


_default: dict = {}


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (the reference-shaped helpers use it)."""
    e = _default.get(device)
    if e is None:
        e = _default[device] = Engine(device)
    return e
