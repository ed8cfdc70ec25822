"""Deterministic synthetic inputs of the overlay path.

The reference ships neither imagery (data/readme.md:20-21) nor the road lines its configs
point at (.MISSING_LARGE_BLOBS), so tests and bench.py run on inputs of the named shapes:
  * an XYZ tile lattice anchored at the AOI's north-west zoom-18 tile (136678, 92197);
  * "ribbon" road polygons: a centreline random walk whose vertex count follows the
    distribution recovered from data/swissTLM3D/roads_lines.shx (median 11, p90 43, p99 133,
    max 1042), buffered by Width/2 with flat caps and mitre joins -- the shape
    prepare_data_obj_detec.py:125-126 produces with ``buffer(Width/2, cap_style=2)`` -- with
    optional rectangular holes (forest difference, :186-191) and two-part MultiPolygons;
  * the (road, tile) pair list the reference gets from gpd.sjoin (statistical_analysis.py:170-171).
Everything is plain numpy, vectorised so that the canton-scale case (1 M roads) is generated
in seconds.  Pixel values are generated on the device (rs_synth_tiles_dev) or by numpy here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .geometry import WEB_MERCATOR_R, PairList, RoadSet

SEED = 20261018
AOI_X0, AOI_Y0, AOI_Z = 136678, 92197, 18

# Width column of data/roads_parameters.xlsx (GDB-Code -> metres); 3.4 m is the only class kept
# by the reference's filter, the others exercise wider / narrower ribbons.
ROAD_WIDTHS_M = np.array([3.4, 6.0, 4.0, 10.2, 8.0, 5.21, 1.81, 1.0, 10.0, 7.0, 9.0])
ROAD_WIDTH_P = np.array([0.60, 0.12, 0.06, 0.03, 0.03, 0.04, 0.04, 0.02, 0.02, 0.02, 0.02])
COS_LAT = 0.6845          # cos(46.8 deg): EPSG:3857 units per ground metre at the AOI = 1 / COS_LAT


@dataclass(frozen=True)
class Grid:
    """nx x ny XYZ tiles of `size` pixels at zoom z; tile index = iy * nx + ix."""
    nx: int
    ny: int
    x0: int = AOI_X0
    y0: int = AOI_Y0
    z: int = AOI_Z
    size: int = 256                # 1024: finer pixels on the same zoom-z footprint (config 5)

    @property
    def n_tiles(self) -> int:
        return self.nx * self.ny

    @property
    def span(self) -> float:
        """tile edge length in EPSG:3857 units"""
        return 2.0 * np.pi * WEB_MERCATOR_R / (2.0 ** self.z)

    @property
    def res(self) -> float:
        return 2.0 * np.pi * WEB_MERCATOR_R / (self.size * 2.0 ** self.z)

    @property
    def origin(self) -> Tuple[float, float]:
        """(x_min, y_max) of the lattice"""
        res256 = self.span
        return (-np.pi * WEB_MERCATOR_R + self.x0 * res256, np.pi * WEB_MERCATOR_R - self.y0 * res256)

    def transforms(self, tile_idx: Optional[np.ndarray] = None) -> np.ndarray:
        """(T, 6) affine (a, b, c, d, e, f) per tile, Affine(res, 0, x_min, 0, -res, y_max)."""
        idx = np.arange(self.n_tiles) if tile_idx is None else np.asarray(tile_idx)
        ix, iy = idx % self.nx, idx // self.nx
        res = self.res
        gt = np.zeros((len(idx), 6), np.float64)
        gt[:, 0] = res
        gt[:, 2] = -np.pi * WEB_MERCATOR_R + (self.x0 + ix) * self.size * res
        gt[:, 4] = -res
        gt[:, 5] = np.pi * WEB_MERCATOR_R - (self.y0 + iy) * self.size * res
        return gt

    def keys(self, tile_idx: Optional[np.ndarray] = None) -> np.ndarray:
        """int64 key of each tile (x, y, z packed): seeds the on-device pixel generator."""
        idx = np.arange(self.n_tiles) if tile_idx is None else np.asarray(tile_idx)
        ix, iy = idx % self.nx, idx // self.nx
        return ((self.x0 + ix).astype(np.int64) << 32) | ((self.y0 + iy).astype(np.int64) << 8) | np.int64(self.z)

    def tile_ids(self, tile_idx: Optional[np.ndarray] = None):
        """'(x, y, z)' strings (prepare_data_obj_detec.py:275-280)."""
        idx = np.arange(self.n_tiles) if tile_idx is None else np.asarray(tile_idx)
        return [f"({self.x0 + i % self.nx}, {self.y0 + i // self.nx}, {self.z})" for i in idx]


def _seg_cumsum(v: np.ndarray, start_idx: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """inclusive cumulative sum restarting at every segment start"""
    c = np.cumsum(v, axis=0)
    base = c[start_idx] - v[start_idx]
    return c - np.repeat(base, counts, axis=0)


def _morton(ix: np.ndarray, iy: np.ndarray) -> np.ndarray:
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFFFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v
    return spread(ix) | (spread(iy) << np.uint64(1))


@dataclass
class RibbonRoads:
    roads: RoadSet
    pairs: PairList
    width_m: np.ndarray          # (R,) ground width
    n_centre: np.ndarray         # (R,) centreline vertex count
    gt_class: np.ndarray         # (R,) int8 0 artificial / 1 natural (config 2 ground truth)


def ribbon_roads(grid: Grid, n_roads: int, seed: int = SEED, hole_frac: float = 0.10, multi_frac: float = 0.05,
                 max_vertices: int = 1042, width_scale: float = 1.0, order: str = "rowband",
                 n_vertices: Optional[Tuple[int, int]] = None, step: Tuple[float, float] = (20.0, 60.0),
                 jitter_deg: float = 15.0, max_holes: int = 3, width_range: Optional[Tuple[float, float]] = None) -> RibbonRoads:
    """n_roads ribbon polygons scattered over `grid`, with their (road, tile) pair list.

    order: 'rowband' sorts roads by (tile row of the first vertex, x) so that a row-band shard of the
    tile lattice owns a contiguous road range; 'morton' sorts by the Morton code of the start tile.
    """
    rng = np.random.default_rng(seed)
    R = int(n_roads)
    span = grid.span
    X0, Y1 = grid.origin
    ext_w, ext_h = grid.nx * span, grid.ny * span

    if n_vertices is None:
        n = np.clip(np.rint(rng.lognormal(np.log(11.0), 1.065, R)), 2, max_vertices).astype(np.int64)
    else:                                                         # config 5: long edge lists, log-uniform vertex counts
        n = np.rint(np.exp(rng.uniform(np.log(n_vertices[0]), np.log(n_vertices[1]), R))).astype(np.int64)
    start = np.stack([X0 + rng.random(R) * ext_w, Y1 - rng.random(R) * ext_h], 1)
    sx = np.floor((start[:, 0] - X0) / span).astype(np.int64)
    sy = np.floor((Y1 - start[:, 1]) / span).astype(np.int64)
    if order == "morton":
        perm = np.argsort(_morton(sx, sy), kind="stable")
    else:
        perm = np.lexsort((start[:, 0], sy))
    n, start = n[perm], start[perm]

    V = int(n.sum())
    first = np.cumsum(n) - n                                  # flat index of each road's first centre vertex
    road_of = np.repeat(np.arange(R), n)
    is_first = np.zeros(V, bool)
    is_first[first] = True

    width = rng.choice(ROAD_WIDTHS_M, R, p=ROAD_WIDTH_P / ROAD_WIDTH_P.sum()) * width_scale
    if width_range is not None:
        width = width_range[0] + (width_range[1] - width_range[0]) * rng.random(R)
    hw = width / (2.0 * COS_LAT)                              # half width in EPSG:3857 units

    dhead = rng.normal(0.0, np.deg2rad(jitter_deg), V)
    dhead[first] = rng.random(R) * 2.0 * np.pi
    head = _seg_cumsum(dhead, first, n)                       # heading of the segment ARRIVING at each vertex
    step_len = step[0] + (step[1] - step[0]) * rng.random(V)
    dxy = np.stack([np.cos(head), np.sin(head)], 1) * step_len[:, None]
    dxy[first] = 0.0
    pos = _seg_cumsum(dxy, first, n) + np.repeat(start, n, axis=0)

    # unit direction of the segment arriving at / leaving each vertex
    d_in = np.stack([np.cos(head), np.sin(head)], 1)
    d_out = np.empty_like(d_in)
    d_out[:-1] = d_in[1:]
    last = first + n - 1
    d_out[last] = d_in[last]
    d_in[first] = d_out[first]
    n_in = np.stack([-d_in[:, 1], d_in[:, 0]], 1)
    n_out = np.stack([-d_out[:, 1], d_out[:, 0]], 1)
    dot = np.maximum(1.0 + (n_in * n_out).sum(1), 0.25)
    mitre = (n_in + n_out) / dot[:, None]
    off = mitre * np.repeat(hw, n)[:, None]
    left, right = pos + off, pos - off

    # ring of road r: left[0..n-1], right[n-1..0], left[0]  -> 2n + 1 vertices
    ring_n = 2 * n + 1
    ring_first = np.cumsum(ring_n) - ring_n
    VV = int(ring_n.sum())
    xy = np.empty((VV, 2), np.float64)
    k = np.arange(V) - np.repeat(first, n)                    # index along the centreline
    xy[np.repeat(ring_first, n) + k] = left
    xy[np.repeat(ring_first, n) + (2 * np.repeat(n, n) - 1 - k)] = right
    xy[ring_first + 2 * n] = left[first]

    # ---- holes and second parts: extra rings appended per road ----
    n_holes = np.where(rng.random(R) < hole_frac, rng.integers(1, max_holes + 1, R), 0)
    is_multi = rng.random(R) < multi_frac
    extra_rings = n_holes + is_multi
    rings_per_road = 1 + extra_rings
    E = int(extra_rings.sum())
    if E:
        e_road = np.repeat(np.arange(R), extra_rings)
        e_k = np.arange(E) - np.repeat(np.cumsum(extra_rings) - extra_rings, extra_rings)
        e_is_part = is_multi[e_road] & (e_k == n_holes[e_road])
        # a hole: small axis-aligned rectangle centred on a centreline vertex (cuts the ribbon like a
        # forest edge); a second part: a detached rectangle beside the first vertex
        pick = first[e_road] + (rng.random(E) * n[e_road]).astype(np.int64)
        c = pos[pick].copy()
        half = np.stack([hw[e_road] * (0.3 + rng.random(E)), hw[e_road] * (0.3 + rng.random(E))], 1) * 1.5
        c[e_is_part] = pos[first[e_road[e_is_part]]] + np.stack([6.0 * hw[e_road[e_is_part]] + 5.0,
                                                                  6.0 * hw[e_road[e_is_part]] + 5.0], 1)
        half[e_is_part] = np.stack([hw[e_road[e_is_part]] * 2.0, hw[e_road[e_is_part]] * 3.0], 1)
        rect = np.stack([c + half * [-1, -1], c + half * [-1, 1], c + half * [1, 1], c + half * [1, -1],
                         c + half * [-1, -1]], 1)            # (E, 5, 2), closed
    # interleave: per road its ribbon ring then its extra rings
    total_rings = int(rings_per_road.sum())
    ring_sizes = np.empty(total_rings, np.int64)
    road_ring_off = np.zeros(R + 1, np.int64)
    road_ring_off[1:] = np.cumsum(rings_per_road)
    ring_sizes[road_ring_off[:-1]] = ring_n
    if E:
        e_ring = road_ring_off[e_road] + 1 + e_k
        ring_sizes[e_ring] = 5
    ring_off = np.zeros(total_rings + 1, np.int64)
    ring_off[1:] = np.cumsum(ring_sizes)
    out_xy = np.empty((int(ring_off[-1]), 2), np.float64)
    dst0 = ring_off[road_ring_off[:-1]]
    kk = np.arange(VV) - np.repeat(ring_first, ring_n)
    out_xy[np.repeat(dst0, ring_n) + kk] = xy
    if E:
        dst = ring_off[e_ring]
        out_xy[(dst[:, None] + np.arange(5)[None, :]).ravel()] = rect.reshape(-1, 2)

    roads = RoadSet.from_arrays(out_xy, ring_off.astype(np.int32), road_ring_off.astype(np.int32),
                                ids=np.arange(R, dtype=np.int64))

    # ---- pair list: tiles under the bbox of every ribbon quad / extra ring ----
    seg = ~is_first
    seg_idx = np.nonzero(seg)[0]
    qx = np.stack([left[seg_idx, 0], left[seg_idx - 1, 0], right[seg_idx, 0], right[seg_idx - 1, 0]], 1)
    qy = np.stack([left[seg_idx, 1], left[seg_idx - 1, 1], right[seg_idx, 1], right[seg_idx - 1, 1]], 1)
    bx0, bx1, by0, by1 = qx.min(1), qx.max(1), qy.min(1), qy.max(1)
    b_road = road_of[seg_idx]
    if E:
        bx0 = np.concatenate([bx0, rect[:, :, 0].min(1)]); bx1 = np.concatenate([bx1, rect[:, :, 0].max(1)])
        by0 = np.concatenate([by0, rect[:, :, 1].min(1)]); by1 = np.concatenate([by1, rect[:, :, 1].max(1)])
        b_road = np.concatenate([b_road, e_road])
    pairs = pairs_from_boxes(grid, R, b_road, bx0, by0, bx1, by1)

    gt_class = (rng.random(R) >= 0.8).astype(np.int8)         # 80 % artificial
    return RibbonRoads(roads, pairs, width, n, gt_class)


def pairs_from_boxes(grid: Grid, n_roads: int, box_road: np.ndarray, bx0, by0, bx1, by1, slack: float = 1e-3) -> PairList:
    """(road, tile) pairs for every lattice tile overlapped by any of a road's boxes."""
    span = grid.span
    X0, Y1 = grid.origin
    ix0 = np.floor((bx0 - slack - X0) / span).astype(np.int64)
    ix1 = np.floor((bx1 + slack - X0) / span).astype(np.int64)
    iy0 = np.floor((Y1 - (by1 + slack)) / span).astype(np.int64)
    iy1 = np.floor((Y1 - (by0 - slack)) / span).astype(np.int64)
    ix0c, ix1c = np.clip(ix0, 0, grid.nx - 1), np.clip(ix1, 0, grid.nx - 1)
    iy0c, iy1c = np.clip(iy0, 0, grid.ny - 1), np.clip(iy1, 0, grid.ny - 1)
    inside = (ix1 >= 0) & (ix0 <= grid.nx - 1) & (iy1 >= 0) & (iy0 <= grid.ny - 1)
    cx = np.where(inside, ix1c - ix0c + 1, 0)
    cy = np.where(inside, iy1c - iy0c + 1, 0)
    cnt = cx * cy
    tot = int(cnt.sum())
    rr = np.repeat(np.asarray(box_road, np.int64), cnt)
    st = np.cumsum(cnt) - cnt
    k = np.arange(tot) - np.repeat(st, cnt)
    cxr = np.maximum(np.repeat(cx, cnt), 1)
    gx = np.repeat(ix0c, cnt) + k % cxr
    gy = np.repeat(iy0c, cnt) + k // cxr
    return PairList.from_pairs(n_roads, rr, gy * grid.nx + gx)


def host_tiles(grid: Grid, channels: int = 3, kind: str = "uniform", seed: int = SEED, dtype=np.uint8,
               tile_idx: Optional[np.ndarray] = None) -> np.ndarray:
    """(T, H, W, C) tiles generated with numpy (tests and the e2e leg of bench.py).
    kind: 'uniform' iid 0..255 | 'asphalt' N(110, 6) clipped | 'class_score' (C == 2: class 0/1/2 in
    16 px blocks + uniform score).  1 % of the pixels have every band 0 (fct_misc.py:117-119)."""
    idx = np.arange(grid.n_tiles) if tile_idx is None else np.asarray(tile_idx)
    T, S = len(idx), grid.size
    rng = np.random.default_rng([seed, 77])
    if kind == "class_score":
        assert channels == 2
        blocks = rng.integers(0, 4, (T, (S + 15) // 16, (S + 15) // 16), dtype=np.uint8)
        blocks[blocks == 3] = 0
        cls = np.repeat(np.repeat(blocks, 16, 1), 16, 2)[:, :S, :S]
        score = rng.integers(0, 256, (T, S, S), dtype=np.uint8)
        return np.ascontiguousarray(np.stack([cls, score], 3))
    if np.dtype(dtype) == np.uint16:
        px = np.clip(rng.lognormal(7.5, 0.8, (T, S, S, channels)), 0, 65535).astype(np.uint16)
    elif kind == "asphalt":
        px = np.clip(np.rint(rng.normal(110.0, 6.0, (T, S, S, channels))), 0, 255).astype(np.uint8)
    else:
        px = rng.integers(0, 256, (T, S, S, channels), dtype=np.uint8)
    holes = rng.random((T, S, S)) < 0.01
    px[holes] = 0
    return px


def wide_polygons(grid: Grid, n_polys: int, seed: int = SEED) -> RibbonRoads:
    """BASELINE.json configs[4]: wide multi-lane polygons with holes and long edge lists for 1024 px tiles --
    15-45 m wide (100-300 px at 0.15 m pixels), 1 k-10 k ring vertices (centreline vertices every 1-3 m),
    1-8 rectangular holes each."""
    return ribbon_roads(grid, n_polys, seed=seed, hole_frac=1.0, multi_frac=0.0, n_vertices=(500, 5000), step=(1.0, 3.0),
                        jitter_deg=1.0, max_holes=8, width_range=(15.0, 45.0))
