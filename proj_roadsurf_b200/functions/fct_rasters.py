"""The two raster-vector call shapes the reference uses directly:
  rasterstats.zonal_stats(vectors, raster, affine=..., stats=[...], nodata=...)   scripts/functions/fct_rasters.py:162-163,
                                                                                 scripts/statistical_analysis/statistical_analysis.py:221-222
  rasterio.features.rasterize(shapes, out_shape, transform=...)                   scripts/sandbox/add_tile_mask.py:112-113
(the download / mosaic helpers of fct_rasters.py are network and IO utilities outside this path).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .._native import RS_NSTAT, STAT_COLS
from ..engine import default_engine
from ..geometry import PairList, RoadSet, TileBatch

IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
_COL = {k: i for i, k in enumerate(STAT_COLS)}
_VALID = {"min", "max", "mean", "count", "sum", "std", "median"}


def rasterize(shapes, out_shape, fill=0, transform=IDENTITY, all_touched=False, default_value=1, dtype=np.uint8):
    """Burn ``default_value`` where a pixel centre falls inside the shapes (GDAL even-odd scanline fill),
    ``fill`` elsewhere.  shapes: iterable of geometries or (geometry, value) pairs.  all_touched is not on the path."""
    if all_touched:
        raise NotImplementedError("all_touched=True is not used by the reference path")
    def is_geom(g):
        return isinstance(g, (dict, np.ndarray)) or hasattr(g, "__geo_interface__")
    geoms, values = [], []
    for sh in shapes:
        if isinstance(sh, (tuple, list)) and len(sh) == 2 and is_geom(sh[0]) and np.isscalar(sh[1]):
            geoms.append(sh[0]); values.append(sh[1])
        else:
            geoms.append(sh); values.append(default_value)
    H, W = int(out_shape[0]), int(out_shape[1])
    out = np.full((H, W), fill, dtype)
    if not geoms:
        return out
    roads = RoadSet.from_geometries(geoms)
    n = roads.n_roads
    pairs = PairList.from_pairs(n, np.arange(n), np.zeros(n, int))
    gt = np.asarray(tuple(transform)[:6], np.float64)[None]
    masks = default_engine().rasterize_pairs_host(roads, gt, H, W, pairs, window="full")
    for i in range(n):                       # merge_alg=replace: later shapes overwrite earlier ones
        out[masks[i] != 0] = values[i]
    return out


def zonal_stats(vectors, raster, affine=None, stats=None, band=1, nodata=None, layer=0, **kwargs) -> List[dict]:
    """rasterstats-shaped zonal statistics of ``vectors`` over a raster array (H, W) or (H, W, C) (``band`` is 1-based) with
    transform ``affine``.  Boundless window, pixels equal to ``nodata`` (and NaN) masked, std with ddof 0, features without
    valid pixels -> count 0 and None for the other statistics.  Percentiles are requested as 'percentile_<q>'.
    uint8 rasters go through the per-feature histograms; every other dtype is taken as float32, the way the reference's DEM
    call reads its raster (fct_rasters.py:147-163: src.read(1) of swissALTI3D, nodata=-9999): per-feature compaction + sort."""
    if stats is None:
        stats = ["count", "min", "max", "mean"]
    if isinstance(stats, str):
        stats = stats.split()
    arr = np.asarray(raster)
    if arr.ndim == 3:
        arr = arr[..., band - 1]
    if affine is None:
        raise ValueError("affine is required for array rasters")
    pct = [float(s.split("_", 1)[1]) for s in stats if s.startswith("percentile_")]
    for s in stats:
        if s not in _VALID and not s.startswith("percentile_"):
            raise ValueError(f"Stat `{s}` not valid")
    geoms = list(vectors)
    roads = RoadSet.from_geometries(geoms)
    n = roads.n_roads
    if n == 0:
        return []
    if arr.dtype != np.uint8:
        if arr.dtype.kind not in "fiu" or (arr.dtype.kind in "iu" and arr.dtype.itemsize > 2) or (arr.dtype.kind == "f" and arr.dtype.itemsize > 4):
            raise TypeError("the GPU zonal statistics path takes uint8 rasters, or rasters that float32 represents exactly")
        # rasterstats' Raster gives array inputs without nodata the value -999 (io.py, with a warning), as the oracle does
        table = default_engine().zonal_stats_f32_host(roads, arr.astype(np.float32), affine, nodata=-999 if nodata is None else nodata,
                                                      ddof=0, percentiles=pct)
        return _rows_to_dicts(table, stats, pct, n)
    tb = TileBatch.from_arrays(arr[None, :, :, None], np.asarray(tuple(affine)[:6], np.float64)[None], nodata)
    pairs = PairList.from_pairs(n, np.arange(n), np.zeros(n, int))
    eng = default_engine()
    hist, _ = eng.zonal_hist_host(roads, tb, pairs, window="boundless")
    if nodata is not None and 0 <= nodata <= 255 and float(nodata) == int(nodata):
        hist[:, :, int(nodata)] = 0           # masked where array == nodata
    table = eng.finalize_stats_host(hist, None, nodata_mode="raw", ddof=0, percentiles=pct)[:, 0, :]
    return _rows_to_dicts(table, stats, pct, n)


def _rows_to_dicts(table, stats, pct, n) -> List[dict]:
    out = []
    for r in range(n):
        row = table[r]
        cnt = int(row[_COL["count"]])
        d = {}
        for s in stats:
            if s == "count":
                d[s] = cnt
            elif cnt == 0:
                d[s] = None
            elif s.startswith("percentile_"):
                d[s] = float(row[RS_NSTAT + pct.index(float(s.split("_", 1)[1]))])
            else:
                d[s] = float(row[_COL[s]])
        out.append(d)
    return out
