"""Oracle: masked pixel extraction and per-road integer accumulators.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Rasterization underneath is
**parity unpinned** (oracle/gdal_fill.py); the wrapper logic restated here follows
  scripts/functions/fct_misc.py:57-123          get_pixel_values
  scripts/statistical_analysis/statistical_analysis.py:179-196   the (road, tile) double loop
and is pinned by running the reference's own get_pixel_values with rasterio
stubbed by this oracle (tests/golden/make_golden.py -> tests/golden/pixel_values.json).

A tile is a mapping with keys
  'data'      uint8/uint16 array, (H, W, C) pixel-interleaved
  'transform' affine (a, b, c, d, e, f), north-up
  'nodata'    None or a number (rasterio dataset.nodata)
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import gdal_fill


def mask_crop(tile: dict, rings: Sequence[np.ndarray]):
    """rasterio.mask.mask(src, [geom], crop=True) -> (count, h, w) filled array, or None
    where rasterio raises ValueError('Input shapes do not overlap raster.')."""
    data = np.asarray(tile["data"])
    H, W = data.shape[:2]
    inside, win = gdal_fill.raster_geometry_mask(tuple(tile["transform"]), rings, W, H)
    if inside is None:
        return None
    c0, r0, w, h = win
    img = np.moveaxis(data[r0:r0 + h, c0:c0 + w, :], 2, 0).copy()  # band-sequential read
    ds_nodata = tile.get("nodata")
    fill = ds_nodata if ds_nodata is not None else 0
    invalid = np.broadcast_to(inside == 0, img.shape).copy()
    if ds_nodata is not None:
        invalid |= (img == ds_nodata)
    img[invalid] = fill
    return img


def get_pixel_values(rings, tile: Optional[dict], BANDS=range(1, 4), pixel_values=None, **kwargs) -> pd.DataFrame:
    """Behavioural restatement of fct_misc.get_pixel_values (fct_misc.py:57-123).

    ``tile is None`` plays the missing-file branch (fct_misc.py:83-85): empty frame.
    """
    if pixel_values is None:
        pixel_values = pd.DataFrame()
    if tile is None:
        return pd.DataFrame()
    out_image = mask_crop(tile, rings)
    if out_image is None:
        raise ValueError("Input shapes do not overlap raster.")
    no_data = tile.get("nodata")
    bands = list(BANDS)
    columns: Dict[str, np.ndarray] = {}
    lengths = []
    for b in bands:
        plane = out_image[b - 1]
        if no_data is None:
            keep = np.ones(plane.shape, bool)      # `data != None` is all-True in numpy
        else:
            keep = plane != no_data
        vals = plane[keep]                          # row-major, like np.extract
        columns[f"band{b}"] = vals
        lengths.append(len(vals))
    longest = max(lengths)
    for b in bands:
        n = lengths[b - 1]                          # the reference indexes with band-1
        if n < longest:
            pad = np.full(longest - n, no_data)
            columns[f"band{b}"] = np.append(columns[f"band{b}"], pad)
    columns.update(kwargs)
    frame = pd.DataFrame(columns)
    if no_data is None:
        allzero = frame[[f"band{b}" for b in bands]].max(axis=1) == 0
        frame = frame.drop(frame[allzero].index)
    return pd.concat([pixel_values, frame], ignore_index=True)


# ----------------------------------------------------------------------------
# integer accumulators (what the CUDA path produces)
# ----------------------------------------------------------------------------
def pair_inside_mask(tile_transform, rings, W: int, H: int):
    """Full-tile uint8 (H, W) mask of the pixels rasterio.mask.mask(crop=True) selects."""
    full = np.zeros((H, W), np.uint8)
    inside, win = gdal_fill.raster_geometry_mask(tuple(tile_transform), rings, W, H)
    if inside is None:
        return full
    c0, r0, w, h = win
    full[r0:r0 + h, c0:c0 + w] = inside
    return full


def zonal_accumulate(tiles: np.ndarray, transforms: np.ndarray, roads: Sequence[Sequence[np.ndarray]],
                     pairs: Iterable[Tuple[int, int]], rescale=None, want_min_zero: bool = False):
    """Per-road histograms over the pair list.

    tiles (T, H, W, C) uint8 (or uint16 with ``rescale``); transforms (T, 6);
    roads[r] = list of rings; pairs = iterable of (tile_idx, road_idx).
    Returns hist uint64 (R, C, 256) and n_allzero uint64 (R,) = in-mask pixels
    whose bands are all 0 (fct_misc.py:117-119 drops exactly those rows when the
    tile has no nodata value).  want_min_zero: third result min_zero uint64 (R,) = sum over the road's
    (road, tile) calls of min over bands of the call's zero-valued in-mask pixels: with tile nodata 0 every
    call drops the zeros per band and pads each band with zeros up to the call's longest band (fct_misc.py:95-111),
    so over all calls band b keeps  zeros(b) - min_zero  zeros.
    """
    T, H, W, C = tiles.shape
    R = len(roads)
    hist = np.zeros((R, C, 256), np.uint64)
    nzero = np.zeros(R, np.uint64)
    minz = np.zeros(R, np.uint64)
    for t, r in pairs:
        m = pair_inside_mask(transforms[t], roads[r], W, H).astype(bool)
        if not m.any():
            continue
        px = tiles[t][m]                             # (n, C)
        if rescale is not None:
            px = rescale(px)
        for c in range(C):
            hist[r, c] += np.bincount(px[:, c], minlength=256).astype(np.uint64)
        nzero[r] += np.uint64(np.count_nonzero(px.max(axis=1) == 0))
        minz[r] += np.uint64(min(int(np.count_nonzero(px[:, c] == 0)) for c in range(C)))
    return (hist, nzero, minz) if want_min_zero else (hist, nzero)


def rescale_u16_to_u8(px: np.ndarray, smin: Sequence[float], smax: Sequence[float], f32: bool = False) -> np.ndarray:
    """gdal.Translate(outputType=GDT_Byte, scaleParams=[[smin,smax,0,255]...]) per band
    (scripts/preprocessing/tif2cog.py:260-270).  **Parity unpinned**:
    GDAL's working precision for UInt16->Byte is not pinned by the reference, hence the flag.
    """
    ft = np.float32 if f32 else np.float64
    out = np.empty(px.shape, np.uint8)
    for c in range(px.shape[-1]):
        lo, hi = ft(smin[c]), ft(smax[c])
        k = ft(255.0) / (hi - lo)
        off = ft(0.0) - lo * k
        v = px[..., c].astype(ft) * k + off
        v = np.clip(v, ft(0.0), ft(255.0))
        out[..., c] = (v + ft(0.5)).astype(np.int32).astype(np.uint8)
    return out


def tiff_assemble(raw: np.ndarray, height: int, width: int, channels: int, sample_bytes: int, planar: int, predictor: int,
                  big_endian: bool, bidx: Optional[Sequence[int]] = None) -> np.ndarray:
    """What a TIFF decoder (libtiff under GDAL under rasterio, fct_misc.py:76-77) does after decompression, for one tile:
    samples in file byte order, TIFF 6.0 section 14 horizontal differencing (predictor 2: every sample is the difference to
    the same sample of the pixel on its left, modulo the sample width), PlanarConfiguration 1 (chunky) or 2 (planes).
    raw: uint8 buffer of the decompressed segments ([H][W][C] or [C][H][W] samples).  bidx: 1-based band selection.
    Returns (H, W, C_out) uint8 | uint16."""
    dt = np.dtype(np.uint8) if sample_bytes == 1 else np.dtype(">u2" if big_endian else "<u2")
    a = np.frombuffer(np.ascontiguousarray(raw, np.uint8).tobytes(), dt)
    a = a.reshape(height, width, channels) if planar == 1 else np.moveaxis(a.reshape(channels, height, width), 0, 2)
    a = a.astype(np.uint8 if sample_bytes == 1 else np.uint16)
    if predictor == 2:
        a = np.cumsum(a.astype(np.uint64), axis=1).astype(a.dtype)          # wraps modulo 2^bits
    if bidx is not None:
        a = a[..., [b - 1 for b in bidx]]
    return np.ascontiguousarray(a)
