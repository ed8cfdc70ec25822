"""Host-side data model of the overlay path: road polygon soup, tile batches, pair lists.

Flat struct-of-arrays layouts, exactly what include/roadsurf_b200.h takes.  The geometry
objects the reference passes around (shapely (Multi)Polygons, scripts/functions/fct_misc.py:72
``mapping(geoms)``) only need ``__geo_interface__`` / GeoJSON-like dicts here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

WEB_MERCATOR_R = 6378137.0


def rings_of(geom) -> List[np.ndarray]:
    """All rings (exterior + holes of every part) of a GeoJSON-like Polygon / MultiPolygon,
    or of a plain list of (n, 2) arrays.  Rings are kept as given (closed rings stay closed)."""
    if isinstance(geom, np.ndarray):
        return [np.ascontiguousarray(geom[:, :2], np.float64)]
    gi = getattr(geom, "__geo_interface__", None)
    if gi is not None:
        geom = gi
    if isinstance(geom, dict):
        if "geometry" in geom and "coordinates" not in geom and "geometries" not in geom:
            geom = geom["geometry"]
        t = geom["type"]
        if t == "Polygon":
            polys = [geom["coordinates"]]
        elif t == "MultiPolygon":
            polys = geom["coordinates"]
        elif t == "GeometryCollection":
            out: List[np.ndarray] = []
            for g in geom["geometries"]:
                out.extend(rings_of(g))
            return out
        else:
            raise ValueError(f"unsupported geometry type {t!r} (the overlay path takes polygons)")
        rings = []
        for poly in polys:
            for ring in poly:
                a = np.asarray(ring, np.float64)
                if a.ndim == 2 and a.shape[0] > 0:
                    rings.append(np.ascontiguousarray(a[:, :2]))
        return rings
    return [np.ascontiguousarray(np.asarray(r, np.float64)[:, :2]) for r in geom]


@dataclass
class RoadSet:
    """Polygon soup of R roads.  xy (V, 2) float64; ring_off (NR+1); road_ring_off (R+1); bbox (R, 4)."""
    xy: np.ndarray
    ring_off: np.ndarray
    road_ring_off: np.ndarray
    bbox: np.ndarray
    ids: Optional[np.ndarray] = None

    @property
    def n_roads(self) -> int:
        return len(self.road_ring_off) - 1

    @property
    def n_rings(self) -> int:
        return len(self.ring_off) - 1

    @property
    def n_verts(self) -> int:
        return int(self.xy.shape[0])

    @staticmethod
    def from_geometries(geoms: Iterable, ids: Optional[Sequence] = None) -> "RoadSet":
        xs: List[np.ndarray] = []
        ring_off = [0]
        road_ring_off = [0]
        bbox = []
        for g in geoms:
            rings = rings_of(g)
            for r in rings:
                xs.append(r)
                ring_off.append(ring_off[-1] + len(r))
            road_ring_off.append(len(ring_off) - 1)
            if rings:
                allv = np.concatenate(rings)
                bbox.append([allv[:, 0].min(), allv[:, 1].min(), allv[:, 0].max(), allv[:, 1].max()])
            else:
                bbox.append([np.inf, np.inf, -np.inf, -np.inf])
        xy = np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros((0, 2)), np.float64)
        if xy.shape[0] >= 2 ** 31 - 1:
            raise ValueError("too many vertices for int32 offsets")
        return RoadSet(xy, np.asarray(ring_off, np.int32), np.asarray(road_ring_off, np.int32),
                       np.asarray(bbox, np.float64).reshape(-1, 4), None if ids is None else np.asarray(ids))

    @staticmethod
    def from_arrays(xy, ring_off, road_ring_off, ids=None) -> "RoadSet":
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        ring_off = np.ascontiguousarray(ring_off, np.int32)
        road_ring_off = np.ascontiguousarray(road_ring_off, np.int32)
        R = len(road_ring_off) - 1
        v0 = ring_off[road_ring_off[:-1]].astype(np.int64)
        v1 = ring_off[road_ring_off[1:]].astype(np.int64)
        bbox = np.empty((R, 4), np.float64)
        nonempty = v1 > v0
        bbox[~nonempty] = [np.inf, np.inf, -np.inf, -np.inf]
        if nonempty.any():
            # reduceat over contiguous vertex ranges: roads are stored in vertex order, so the range of a non-empty
            # road ends where the next non-empty road starts
            idx = v0[nonempty].astype(np.intp)
            assert np.all(np.diff(idx) >= 0), "roads must be stored in vertex order"
            bbox[nonempty] = np.stack([np.minimum.reduceat(xy[:, 0], idx), np.minimum.reduceat(xy[:, 1], idx),
                                       np.maximum.reduceat(xy[:, 0], idx), np.maximum.reduceat(xy[:, 1], idx)], 1)
        return RoadSet(xy, ring_off, road_ring_off, bbox, None if ids is None else np.asarray(ids))

    def rings(self, r: int) -> List[np.ndarray]:
        g0, g1 = self.road_ring_off[r], self.road_ring_off[r + 1]
        return [self.xy[self.ring_off[g]:self.ring_off[g + 1]] for g in range(g0, g1)]

    def subset(self, road_idx: np.ndarray) -> "RoadSet":
        """The roads `road_idx` (any order, repeats allowed) as a new compact soup; vectorised gathers."""
        road_idx = np.asarray(road_idx, np.int64)
        g0 = self.road_ring_off[road_idx].astype(np.int64)
        ng = self.road_ring_off[road_idx + 1].astype(np.int64) - g0
        new_rro = np.zeros(len(road_idx) + 1, np.int64)
        new_rro[1:] = np.cumsum(ng)
        ring_src = np.repeat(g0 - new_rro[:-1], ng) + np.arange(new_rro[-1])           # old ring of each new ring
        v0 = self.ring_off[ring_src].astype(np.int64)
        nv = self.ring_off[ring_src + 1].astype(np.int64) - v0
        new_ro = np.zeros(len(ring_src) + 1, np.int64)
        new_ro[1:] = np.cumsum(nv)
        vert_src = np.repeat(v0 - new_ro[:-1], nv) + np.arange(new_ro[-1])
        return RoadSet(np.ascontiguousarray(self.xy[vert_src]), new_ro.astype(np.int32), new_rro.astype(np.int32),
                       np.ascontiguousarray(self.bbox[road_idx]), None if self.ids is None else self.ids[road_idx])


def xyz_tile_transform(tx: int, ty: int, z: int, size: int = 256) -> Tuple[float, ...]:
    """Affine (a, b, c, d, e, f) of the XYZ tile (tx, ty, z) in EPSG:3857."""
    res = 2.0 * np.pi * WEB_MERCATOR_R / (size * 2.0 ** z)
    x_min = -np.pi * WEB_MERCATOR_R + tx * size * res
    y_max = np.pi * WEB_MERCATOR_R - ty * size * res
    return (res, 0.0, x_min, 0.0, -res, y_max)


@dataclass
class TileBatch:
    """T tiles of identical shape.  pixels (T, H, W, C) uint8|uint16 interleaved; gt (T, 6)."""
    pixels: Optional[np.ndarray]
    gt: np.ndarray
    height: int
    width: int
    channels: int
    nodata: Optional[float] = None
    ids: Optional[Sequence] = None

    @property
    def n_tiles(self) -> int:
        return int(self.gt.shape[0])

    @staticmethod
    def from_arrays(pixels, gt, nodata=None, ids=None, layout: str = "HWC") -> "TileBatch":
        px = np.asarray(pixels)
        if px.ndim == 3:
            px = px[..., None] if layout == "HWC" else px[:, None]
        if layout == "CHW":      # rasterio band-sequential reads
            px = np.moveaxis(px, 1, 3)
        px = np.ascontiguousarray(px)
        if px.dtype not in (np.uint8, np.uint16):
            raise ValueError("tiles must be uint8 or uint16")
        gt = np.ascontiguousarray(gt, np.float64).reshape(-1, 6)
        if gt.shape[0] != px.shape[0]:
            raise ValueError("one transform per tile expected")
        return TileBatch(px, gt, px.shape[1], px.shape[2], px.shape[3], nodata, ids)

    def extents(self) -> np.ndarray:
        """(T, 4) xmin, ymin, xmax, ymax of every tile (north-up)."""
        a, c, e, f = self.gt[:, 0], self.gt[:, 2], self.gt[:, 4], self.gt[:, 5]
        x0, x1 = c, c + a * self.width
        y0, y1 = f, f + e * self.height
        return np.stack([np.minimum(x0, x1), np.minimum(y0, y1), np.maximum(x0, x1), np.maximum(y0, y1)], 1)


@dataclass
class PairList:
    """Road-major CSR of (road, tile) pairs: the tiles each road is masked against
    (scripts/statistical_analysis/statistical_analysis.py:170-171,183-191)."""
    road_pair_off: np.ndarray
    pair_tile: np.ndarray

    @property
    def n_pairs(self) -> int:
        return int(self.pair_tile.shape[0])

    @staticmethod
    def from_pairs(n_roads: int, road_idx, tile_idx) -> "PairList":
        road_idx = np.asarray(road_idx, np.int64)
        tile_idx = np.asarray(tile_idx, np.int64)
        order = np.lexsort((tile_idx, road_idx))
        road_idx, tile_idx = road_idx[order], tile_idx[order]
        if len(road_idx):                                   # drop_duplicates(['id', 'OBJECTID'])
            keep = np.ones(len(road_idx), bool)
            keep[1:] = (road_idx[1:] != road_idx[:-1]) | (tile_idx[1:] != tile_idx[:-1])
            road_idx, tile_idx = road_idx[keep], tile_idx[keep]
        off = np.zeros(n_roads + 1, np.int64)
        np.add.at(off, road_idx + 1, 1)
        off = np.cumsum(off)
        if off[-1] >= 2 ** 31 - 1:
            raise ValueError("too many pairs for int32 offsets")
        return PairList(off.astype(np.int32), tile_idx.astype(np.int32))

    def restrict_tiles(self, tile_lo: int, tile_hi: int) -> "PairList":
        """Pairs whose tile index lies in [tile_lo, tile_hi), tile indices rebased to tile_lo
        (what a tile shard keeps); road numbering unchanged."""
        keep = (self.pair_tile >= tile_lo) & (self.pair_tile < tile_hi)
        csum = np.concatenate([[0], np.cumsum(keep)])
        off = csum[self.road_pair_off.astype(np.int64)]
        return PairList(off.astype(np.int32), (self.pair_tile[keep] - tile_lo).astype(np.int32))

    def take_roads(self, road_idx: np.ndarray) -> "PairList":
        """Pair list of the roads `road_idx`, renumbered 0..len-1 in that order."""
        road_idx = np.asarray(road_idx, np.int64)
        p0 = self.road_pair_off[road_idx].astype(np.int64)
        n = self.road_pair_off[road_idx + 1].astype(np.int64) - p0
        off = np.zeros(len(road_idx) + 1, np.int64)
        off[1:] = np.cumsum(n)
        src = np.repeat(p0 - off[:-1], n) + np.arange(off[-1])
        return PairList(off.astype(np.int32), np.ascontiguousarray(self.pair_tile[src]))

    def road_of_pair(self) -> np.ndarray:
        return np.repeat(np.arange(len(self.road_pair_off) - 1, dtype=np.int32), np.diff(self.road_pair_off))


@dataclass
class Lattice:
    """Regular lattice a tile set sits on: cell (ix, iy) = [x0 + ix*tile_w, +tile_w] x [y0 + iy*tile_h, +tile_h],
    lut[iy, ix] = tile index or -1."""
    x0: float
    y0: float
    tile_w: float
    tile_h: float
    lut: np.ndarray

    @property
    def nx(self) -> int:
        return int(self.lut.shape[1])

    @property
    def ny(self) -> int:
        return int(self.lut.shape[0])


def lattice_of(tiles: TileBatch, max_cells_factor: int = 4) -> Optional[Lattice]:
    """The lattice of an XYZ-style tile set (equal, axis-aligned, snapped tiles), or None."""
    T = tiles.n_tiles
    if T == 0:
        return None
    ext = tiles.extents()
    tw, th = ext[:, 2] - ext[:, 0], ext[:, 3] - ext[:, 1]
    if not (np.allclose(tw, tw[0], rtol=1e-9, atol=0) and np.allclose(th, th[0], rtol=1e-9, atol=0)):
        return None
    X0, Y0 = float(ext[:, 0].min()), float(ext[:, 1].min())
    ix = np.rint((ext[:, 0] - X0) / tw[0]).astype(np.int64)
    iy = np.rint((ext[:, 1] - Y0) / th[0]).astype(np.int64)
    if not (np.allclose(X0 + ix * tw[0], ext[:, 0], rtol=0, atol=1e-6 * tw[0]) and
            np.allclose(Y0 + iy * th[0], ext[:, 1], rtol=0, atol=1e-6 * th[0])):
        return None
    nx, ny = int(ix.max()) + 1, int(iy.max()) + 1
    if nx * ny > max(max_cells_factor * T, 1 << 22):
        return None
    lut = np.full((ny, nx), -1, np.int32)
    lut[iy, ix] = np.arange(T, dtype=np.int32)
    return Lattice(X0, Y0, float(tw[0]), float(th[0]), lut)


def pairs_by_bbox(roads: RoadSet, tiles: TileBatch, chunk: int = 4096) -> PairList:
    """Host broad phase: every (road, tile) whose bounding boxes overlap (closed comparison).

    A superset of the reference's ``gpd.sjoin(tiles, roads)`` (statistical_analysis.py:170):
    pairs whose polygon misses the tile contribute no pixel, so results are identical.
    Tiles on a regular lattice (XYZ grids) are found by index arithmetic; anything else by
    chunked brute force.
    """
    ext = tiles.extents()
    R, T = roads.n_roads, tiles.n_tiles
    if R == 0 or T == 0:
        return PairList.from_pairs(R, [], [])
    bb = roads.bbox
    tw = ext[:, 2] - ext[:, 0]
    th = ext[:, 3] - ext[:, 1]
    regular = np.allclose(tw, tw[0], rtol=1e-9, atol=0) and np.allclose(th, th[0], rtol=1e-9, atol=0)
    if regular:
        X0, Y0 = ext[:, 0].min(), ext[:, 1].min()
        ix = np.rint((ext[:, 0] - X0) / tw[0]).astype(np.int64)
        iy = np.rint((ext[:, 1] - Y0) / th[0]).astype(np.int64)
        snapped = np.allclose(X0 + ix * tw[0], ext[:, 0], rtol=0, atol=1e-6 * tw[0]) and \
            np.allclose(Y0 + iy * th[0], ext[:, 1], rtol=0, atol=1e-6 * th[0])
        nx, ny = int(ix.max()) + 1, int(iy.max()) + 1
        if snapped and nx * ny <= max(4 * T, 1 << 22):
            lut = np.full((ny, nx), -1, np.int64)
            lut[iy, ix] = np.arange(T)
            valid = bb[:, 0] <= bb[:, 2]
            # candidate index ranges with one cell of slack, exact overlap test afterwards
            ix0 = np.clip(np.floor((bb[:, 0] - X0) / tw[0]).astype(np.int64) - 1, 0, nx - 1)
            ix1 = np.clip(np.floor((bb[:, 2] - X0) / tw[0]).astype(np.int64) + 1, 0, nx - 1)
            iy0 = np.clip(np.floor((bb[:, 1] - Y0) / th[0]).astype(np.int64) - 1, 0, ny - 1)
            iy1 = np.clip(np.floor((bb[:, 3] - Y0) / th[0]).astype(np.int64) + 1, 0, ny - 1)
            cx = np.where(valid, ix1 - ix0 + 1, 0)
            cy = np.where(valid, iy1 - iy0 + 1, 0)
            cnt = cx * cy
            road_rep = np.repeat(np.arange(R), cnt)
            start = np.cumsum(cnt) - cnt
            k = np.arange(cnt.sum()) - np.repeat(start, cnt)
            cxr = np.repeat(cx, cnt)
            gx = np.repeat(ix0, cnt) + k % np.maximum(cxr, 1)
            gy = np.repeat(iy0, cnt) + k // np.maximum(cxr, 1)
            tidx = lut[gy, gx]
            ok = tidx >= 0
            road_rep, tidx = road_rep[ok], tidx[ok]
            e = ext[tidx]
            b = bb[road_rep]
            hit = (b[:, 0] <= e[:, 2]) & (b[:, 2] >= e[:, 0]) & (b[:, 1] <= e[:, 3]) & (b[:, 3] >= e[:, 1])
            return PairList.from_pairs(R, road_rep[hit], tidx[hit])
    rr, tt = [], []
    for s in range(0, R, chunk):
        b = bb[s:s + chunk]
        hit = (b[:, None, 0] <= ext[None, :, 2]) & (b[:, None, 2] >= ext[None, :, 0]) & \
              (b[:, None, 1] <= ext[None, :, 3]) & (b[:, None, 3] >= ext[None, :, 1])
        r, t = np.nonzero(hit)
        rr.append(r + s)
        tt.append(t)
    return PairList.from_pairs(R, np.concatenate(rr), np.concatenate(tt))


def rasterio_window(transform, bbox, width: int, height: int):
    """rasterio.features.geometry_window (pad 0, not boundless) of a geometry's bounds on a north-up raster,
    with affine's arithmetic order: (col_off, row_off, w, h), or None where rasterio raises WindowError
    ("Input shapes do not overlap raster." from rasterio.mask.mask, fct_misc.py:77)."""
    import math
    sa, sb, sc, sd, se, sf = (float(v) for v in transform)
    det = sa * se - sb * sd
    if det == 0.0:
        return None
    idet = 1.0 / det
    ra, rb, rd, re = se * idet, -sb * idet, -sd * idet, sa * idet
    rc, rf = -sc * ra - sf * rb, -sc * rd - sf * re
    xmin, ymin, xmax, ymax = (float(v) for v in bbox)
    if not all(math.isfinite(v) for v in (xmin, ymin, xmax, ymax)):
        return None
    xs = [vx * ra + vy * rb + rc for vx in (xmin, xmax) for vy in (ymin, ymax)]
    ys = [vx * rd + vy * re + rf for vx in (xmin, xmax) for vy in (ymin, ymax)]
    c0, r0 = math.floor(min(xs)), math.floor(min(ys))
    w, h = max(math.ceil(max(xs)) - c0, 0), max(math.ceil(max(ys)) - r0, 0)
    if r0 >= height or r0 + h <= 0 or c0 >= width or c0 + w <= 0:
        return None
    cc0, rr0 = max(c0, 0), max(r0, 0)
    return (cc0, rr0, min(c0 + w, width) - cc0, min(r0 + h, height) - rr0)


def bbox_pairs(a_bbox: np.ndarray, b_bbox: np.ndarray, chunk: int = 512):
    """(ia, ib) of every pair of boxes (xmin, ymin, xmax, ymax) that overlap (closed), ordered by ia then ib: the candidate
    pairs of gpd.overlay / sjoin before the exact test."""
    a = np.asarray(a_bbox, np.float64).reshape(-1, 4)
    b = np.asarray(b_bbox, np.float64).reshape(-1, 4)
    ia_all, ib_all = [], []
    for lo in range(0, len(a), chunk):
        c = a[lo:lo + chunk]
        hit = ((c[:, None, 0] <= b[None, :, 2]) & (c[:, None, 2] >= b[None, :, 0]) &
               (c[:, None, 1] <= b[None, :, 3]) & (c[:, None, 3] >= b[None, :, 1]))
        i, j = np.nonzero(hit)
        ia_all.append(i + lo)
        ib_all.append(j)
    if not ia_all:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(ia_all).astype(np.int64), np.concatenate(ib_all).astype(np.int64)
