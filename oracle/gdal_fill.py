"""Oracle: GDAL 3.0.x scanline polygon fill and the rasterio 1.3.2 wrappers around it.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned**: GDAL /
rasterio are third-party dependencies of the reference (requirements.txt:77,211)
that are not vendored and not importable here; this file restates their
published algorithm and is anchored on the reference call sites
  scripts/functions/fct_misc.py:77            rasterio.mask.mask(src, geoms, crop=True)
  scripts/sandbox/add_tile_mask.py:112-113    rasterio.features.rasterize(shapes, out_shape)
  scripts/functions/fct_rasters.py:162-163    rasterstats.zonal_stats (rasterize_geom)

Restated upstream routines:
  GDAL alg/llrasterize.cpp      GDALdllImageFilledPolygon + gvBurnScanline
  GDAL gcore/gdal_misc.cpp      GDALInvGeoTransform (north-up special case)
  GDAL alg/gdaltransformer.cpp  GDALGenImgProjTransform (dst geotransform only)
  affine 2.3.1                  Affine.__invert__, __mul__ (operation order)
  rasterio features.py/mask.py  bounds, geometry_window, raster_geometry_mask

Every float operation below is an IEEE-754 binary64 operation in the same order
as upstream (no fused multiply-add): that order is what decides single pixels.

Geometry model: a geometry is a list of rings; a ring is an (n, 2) float64 array
of vertices exactly as given (GeoJSON rings are closed: last == first).  All
rings of all parts of a (Multi)Polygon are concatenated, which is what
GDALCollectRingsFromGeometry hands to the fill in GDAL 3.0.x (even-odd rule over
all rings: holes and overlapping parts cancel).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

Ring = np.ndarray
Affine6 = Tuple[float, float, float, float, float, float]  # (a, b, c, d, e, f)


# ----------------------------------------------------------------------------
# geometry helpers
# ----------------------------------------------------------------------------
def rings_from_geojson(geom) -> List[Ring]:
    """All rings of a GeoJSON-like Polygon / MultiPolygon, in GDAL collection order.

    Mirrors shapely.geometry.mapping() output consumed at fct_misc.py:72.
    """
    geom = getattr(geom, "__geo_interface__", geom)
    if "geometry" in geom and "coordinates" not in geom:
        geom = geom["geometry"]
    t = geom["type"]
    if t == "Polygon":
        polys = [geom["coordinates"]]
    elif t == "MultiPolygon":
        polys = geom["coordinates"]
    elif t == "GeometryCollection":
        out: List[Ring] = []
        for g in geom["geometries"]:
            out.extend(rings_from_geojson(g))
        return out
    else:
        raise ValueError(f"unsupported geometry type {t!r}")
    rings = []
    for poly in polys:
        for ring in poly:
            arr = np.asarray(ring, dtype=np.float64)
            if arr.ndim != 2 or arr.shape[0] == 0:
                continue
            rings.append(np.ascontiguousarray(arr[:, :2]))
    return rings


# ----------------------------------------------------------------------------
# affine 2.3.1 arithmetic (operation order matters)
# ----------------------------------------------------------------------------
def affine_invert(t: Affine6) -> Affine6:
    """Affine.__invert__ (affine/__init__.py)."""
    sa, sb, sc, sd, se, sf = t
    det = sa * se - sb * sd
    idet = 1.0 / det
    ra = se * idet
    rb = -sb * idet
    rd = -sd * idet
    re = sa * idet
    return (ra, rb, -sc * ra - sf * rb, rd, re, -sc * rd - sf * re)


def affine_apply(t: Affine6, x, y):
    """Affine.__mul__ on a point: (vx*sa + vy*sb + sc, vx*sd + vy*se + sf)."""
    sa, sb, sc, sd, se, sf = t
    return (x * sa + y * sb + sc, x * sd + y * se + sf)


def affine_mul_translation(t: Affine6, xoff: float, yoff: float) -> Affine6:
    """t * Affine.translation(xoff, yoff) -- rasterio.windows.transform()."""
    sa, sb, sc, sd, se, sf = t
    oa, ob, oc, od, oe, of = 1.0, 0.0, float(xoff), 0.0, 1.0, float(yoff)
    return (
        sa * oa + sb * od,
        sa * ob + sb * oe,
        sa * oc + sb * of + sc,
        sd * oa + se * od,
        sd * ob + se * oe,
        sd * oc + se * of + sf,
    )


IDENTITY: Affine6 = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)


# ----------------------------------------------------------------------------
# GDAL geotransform inverse + vertex transform
# ----------------------------------------------------------------------------
def gdal_inv_geotransform(t: Affine6):
    """GDALInvGeoTransform on gt = (c, a, b, f, d, e).  Returns inv[0..5]."""
    a, b, c, d, e, f = t
    gt = (c, a, b, f, d, e)
    if gt[2] == 0.0 and gt[4] == 0.0 and gt[1] != 0.0 and gt[5] != 0.0:
        return (-gt[0] / gt[1], 1.0 / gt[1], 0.0, -gt[3] / gt[5], 0.0, 1.0 / gt[5])
    det = gt[1] * gt[5] - gt[2] * gt[4]
    if abs(det) < 1e-15:
        raise ValueError("non-invertible geotransform")
    inv_det = 1.0 / det
    o1 = gt[5] * inv_det
    o4 = -gt[4] * inv_det
    o2 = -gt[2] * inv_det
    o5 = gt[1] * inv_det
    o0 = (gt[2] * gt[3] - gt[0] * gt[5]) * inv_det
    o3 = (-gt[1] * gt[3] + gt[0] * gt[4]) * inv_det
    return (o0, o1, o2, o3, o4, o5)


def world_to_pixel(inv, xy: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """GDALGenImgProjTransform, dst side: px = inv0 + X*inv1 + Y*inv2 (left to right)."""
    X = xy[:, 0]
    Y = xy[:, 1]
    px = inv[0] + X * inv[1] + Y * inv[2]
    py = inv[3] + X * inv[4] + Y * inv[5]
    return px, py


# ----------------------------------------------------------------------------
# GDALdllImageFilledPolygon
# ----------------------------------------------------------------------------
def _c_int(v: float) -> int:
    """C (int) cast: truncation toward zero."""
    return int(v)


def filled_polygon_spans(rings_px: Sequence[Tuple[np.ndarray, np.ndarray]], W: int, H: int):
    """Yield the gvBurnScanline calls (y, xs, xe) of GDALdllImageFilledPolygon, unclamped.

    rings_px: sequence of (X, Y) arrays in pixel/line space, one per ring.
    """
    rings_px = [(np.asarray(x, np.float64), np.asarray(y, np.float64)) for x, y in rings_px if len(x)]
    if not rings_px:
        return
    ally = np.concatenate([y for _, y in rings_px])
    miny = _c_int(float(ally.min()))
    maxy = _c_int(float(ally.max()))
    if miny < 0:
        miny = 0
    if maxy >= H:
        maxy = H - 1
    minx, maxx = 0, W - 1

    # edge (ind1 -> ind2): first index of a ring pairs with the ring's last vertex
    x1 = np.concatenate([np.roll(x, 1) for x, _ in rings_px])
    y1 = np.concatenate([np.roll(y, 1) for _, y in rings_px])
    x2 = np.concatenate([x for x, _ in rings_px])
    y2 = np.concatenate([y for _, y in rings_px])

    for y in range(miny, maxy + 1):
        dy = y + 0.5
        skip = ((y1 < dy) & (y2 < dy)) | ((y1 > dy) & (y2 > dy))
        cand = np.nonzero(~skip)[0]
        ints: List[int] = []
        for i in cand:  # upstream vertex order (matters only for the order of burn calls)
            a1, a2, b1, b2 = y1[i], y2[i], x1[i], x2[i]
            if a1 < a2:
                dy1, dy2, dx1, dx2 = a1, a2, b1, b2
            elif a1 > a2:
                dy2, dy1, dx2, dx1 = a1, a2, b1, b2
            else:
                # horizontal edge lying exactly on the scanline: filled separately,
                # only for the direction X[ind1] > X[ind2]
                if b1 > b2:
                    hx1 = int(math.floor(b2 + 0.5))
                    hx2 = int(math.floor(b1 + 0.5))
                    if hx1 > maxx or hx2 <= minx:
                        continue
                    yield (y, hx1, hx2 - 1)
                continue
            if dy < dy2 and dy >= dy1:
                intersect = (dy - dy1) * (dx2 - dx1) / (dy2 - dy1) + dx1
                ints.append(int(math.floor(intersect + 0.5)))  # ROUND BEFORE SORT
        ints.sort()
        i = 0
        while i + 1 < len(ints):
            if ints[i] <= maxx and ints[i + 1] > minx:
                yield (y, ints[i], ints[i + 1] - 1)
            i += 2


def burn_spans(spans: Iterable[Tuple[int, int, int]], W: int, H: int) -> np.ndarray:
    """gvBurnScanline for a uint8 buffer, burn value 1, replace mode."""
    out = np.zeros((H, W), dtype=np.uint8)
    for y, xs, xe in spans:
        if xs > xe:
            continue
        if xs < 0:
            xs = 0
        if xe >= W:
            xe = W - 1
        if y < 0 or y >= H or xs > xe:
            continue
        out[y, xs:xe + 1] = 1
    return out


# ----------------------------------------------------------------------------
# rasterio wrappers
# ----------------------------------------------------------------------------
def rasterize(rings: Sequence[Ring], out_shape: Tuple[int, int], transform: Affine6 = IDENTITY) -> np.ndarray:
    """rasterio.features.rasterize(shapes, out_shape, transform), all_touched=False,
    default_value=1, fill=0 (add_tile_mask.py:112-113).  Returns uint8 (H, W)."""
    H, W = int(out_shape[0]), int(out_shape[1])
    if H <= 0 or W <= 0:
        return np.zeros((max(H, 0), max(W, 0)), np.uint8)
    if transform[1] != 0.0 or transform[3] != 0.0:
        raise ValueError("rotated transforms are outside the hot path")
    inv = gdal_inv_geotransform(transform)
    rings_px = [world_to_pixel(inv, np.asarray(r, np.float64)) for r in rings if len(r)]
    return burn_spans(filled_polygon_spans(rings_px, W, H), W, H)


def geometry_window(transform: Affine6, rings: Sequence[Ring], width: int, height: int
                    ) -> Optional[Tuple[int, int, int, int]]:
    """rasterio.features.geometry_window(dataset, shapes) with pad 0, boundless=False.

    Returns (col_off, row_off, w, h), or None where rasterio raises WindowError
    ("windows do not intersect").  Bounds are taken vertex by vertex in pixel space
    through ~dataset.transform (rasterio _features._bounds).
    """
    inv = affine_invert(transform)
    xs = []
    ys = []
    for r in rings:
        r = np.asarray(r, np.float64)
        if len(r) == 0:
            continue
        px, py = affine_apply(inv, r[:, 0], r[:, 1])
        xs.append(px)
        ys.append(py)
    if not xs:
        return None
    allx = np.concatenate(xs)
    ally = np.concatenate(ys)
    left, right = float(allx.min()), float(allx.max())
    top, bottom = float(ally.min()), float(ally.max())
    row_start, row_stop = int(math.floor(top)), int(math.ceil(bottom))
    col_start, col_stop = int(math.floor(left)), int(math.ceil(right))
    w = max(col_stop - col_start, 0)
    h = max(row_stop - row_start, 0)
    # rasterio.windows.intersect(): ranges that only touch do not intersect
    r0, r1 = row_start, row_start + h
    c0, c1 = col_start, col_start + w
    if r0 >= height or r1 <= 0 or c0 >= width or c1 <= 0:
        return None
    rr0, rr1 = max(r0, 0), min(r1, height)
    cc0, cc1 = max(c0, 0), min(c1, width)
    return (cc0, rr0, cc1 - cc0, rr1 - rr0)


def raster_geometry_mask(transform: Affine6, rings: Sequence[Ring], width: int, height: int):
    """rasterio.mask.raster_geometry_mask(dataset, shapes, crop=True).

    Returns (inside, window) where inside is a uint8 (h, w) array with 1 = pixel
    selected by the shapes (i.e. ~shape_mask), or (None, None) where rasterio raises
    ValueError('Input shapes do not overlap raster.').
    """
    win = geometry_window(transform, rings, width, height)
    if win is None:
        return None, None
    col_off, row_off, w, h = win
    wt = affine_mul_translation(transform, col_off, row_off)
    inside = rasterize(rings, (h, w), wt)
    return inside, win
