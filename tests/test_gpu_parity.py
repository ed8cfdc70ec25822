"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: -m gpu.

Integer outputs (masks, histograms, counts, cover codes, confusion counts) must be bit-exact;
float outputs (mean, std, margin, percentiles, P/R/F1) within 1e-6 relative (north star), and
in practice they are compared at 1e-12 because both sides evaluate the same float64 formula on
identical integers.
"""
import numpy as np
import pytest

from oracle import cport, gdal_fill, raster as oraster, stats as ostats, vote as ovote
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.geometry import PairList, RoadSet, TileBatch, pairs_by_bbox

pytestmark = pytest.mark.gpu

RTOL = 1e-6       # north-star tolerance for float outputs


@pytest.fixture(scope="module")
def eng():
    from proj_roadsurf_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def ring(*pts):
    pts = list(pts)
    if pts[0] != pts[-1]:
        pts.append(pts[0])
    return np.array(pts, np.float64)


def oracle_hist(rr_roads: RoadSet, pairs: PairList, tiles, gt, **kw):
    return cport.zonal_accumulate(rr_roads.xy, rr_roads.ring_off, rr_roads.road_ring_off, pairs.road_pair_off,
                                  pairs.pair_tile, tiles, gt, **kw)


# ------------------------------------------------------------------------------------------
# masks: rasterio.features.rasterize (identity / north-up transform) and mask.mask(crop=True)
# ------------------------------------------------------------------------------------------
KAT_SHAPES = [
    ("integer_rect", [ring((2, 1), (6, 1), (6, 4), (2, 4))], (8, 8)),
    ("on_centres_cw", [ring((1.5, 1.5), (4.5, 1.5), (4.5, 3.5), (1.5, 3.5))], (6, 8)),
    ("on_centres_ccw", [ring((1.5, 1.5), (1.5, 3.5), (4.5, 3.5), (4.5, 1.5))], (6, 8)),
    ("triangle_vertex_on_scanline", [ring((1, 0), (7, 0), (4, 2.5))], (4, 8)),
    ("hole", [ring((0, 0), (10, 0), (10, 10), (0, 10)), ring((3, 3), (3, 7), (7, 7), (7, 3))], (10, 10)),
    ("even_odd_parts", [ring((0, 0), (6, 0), (6, 4), (0, 4)), ring((4, 0), (10, 0), (10, 4), (4, 4))], (4, 10)),
    ("partly_outside", [ring((-5, -5), (3, -5), (3, 2), (-5, 2))], (8, 8)),
    ("fully_outside", [ring((20, 20), (30, 20), (30, 30), (20, 30))], (8, 8)),
    ("covers_all", [ring((-100, -100), (100, -100), (100, 100), (-100, 100))], (8, 8)),
    ("sliver_empty", [ring((2.6, 0), (2.9, 0), (2.9, 5), (2.6, 5))], (5, 8)),
    ("sliver_one_col", [ring((2.4, 0), (2.6, 0), (2.6, 5), (2.4, 5))], (5, 8)),
    ("unclosed", [ring((1, 1), (6, 1), (6, 5), (1, 5))[:-1]], (8, 8)),
]


@pytest.mark.parametrize("name,rings,shape", KAT_SHAPES, ids=[k[0] for k in KAT_SHAPES])
def test_rasterize_kat(eng, name, rings, shape):
    H, W = shape
    roads = RoadSet.from_geometries([rings])
    pairs = PairList.from_pairs(1, [0], [0])
    gt = np.array([[1.0, 0, 0, 0, 1.0, 0]])
    m = eng.rasterize_pairs_host(roads, gt, H, W, pairs, window="full")[0]
    assert np.array_equal(m, gdal_fill.rasterize(rings, (H, W))), name
    # crop=True window: same selected pixels (expanded to the full raster by the oracle)
    m2 = eng.rasterize_pairs_host(roads, gt, H, W, pairs, window="crop")[0]
    assert np.array_equal(m2, oraster.pair_inside_mask(gt[0], rings, W, H)), name


def _random_polygon(rng, W, H, n):
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(0.2, 1.0, n) * min(W, H) * 0.6
    cx, cy = rng.uniform(-0.2 * W, 1.2 * W), rng.uniform(-0.2 * H, 1.2 * H)
    pts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)
    if rng.random() < 0.3:   # half-pixel lattice: ties, horizontal edges on scanlines
        pts = np.round(pts * 2) / 2
    return np.concatenate([pts, pts[:1]])


@pytest.mark.parametrize("window", ["full", "crop"])
def test_rasterize_random_polygons_bit_exact(eng, window):
    rng = np.random.default_rng(20261018)
    H, W = 48, 40
    geoms = [[_random_polygon(rng, W, H, int(rng.integers(3, 14))) for _ in range(int(rng.integers(1, 4)))]
             for _ in range(400)]
    roads = RoadSet.from_geometries(geoms)
    n = roads.n_roads
    pairs = PairList.from_pairs(n, np.arange(n), np.zeros(n, int))
    gt = np.array([[1.0, 0, 0, 0, 1.0, 0]])
    masks = eng.rasterize_pairs_host(roads, gt, H, W, pairs, window=window)
    for i, rings in enumerate(geoms):
        exp = cport.rasterize(rings, (H, W)) if window == "full" else cport.pair_mask_full(gt[0], rings, W, H)
        assert np.array_equal(masks[i], exp), i


def test_rasterize_world_transform_tiles(eng):
    """north-up EPSG:3857 zoom-18 transforms: inverse geotransform arithmetic."""
    g = synth.Grid(4, 4)
    rr = synth.ribbon_roads(g, 24, seed=5)
    gt = g.transforms()
    masks = eng.rasterize_pairs_host(rr.roads, gt, 256, 256, rr.pairs, window="crop")
    road_of = rr.pairs.road_of_pair()
    assert rr.pairs.n_pairs > 0
    total = 0
    for p in range(rr.pairs.n_pairs):
        exp = cport.pair_mask_full(gt[rr.pairs.pair_tile[p]], rr.roads.rings(int(road_of[p])), 256, 256)
        assert np.array_equal(masks[p], exp), p
        total += int(exp.sum())
    assert total > 10000


# ------------------------------------------------------------------------------------------
# fused rasterize + zonal histograms
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("channels", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["uniform", "asphalt"])
def test_zonal_hist_bands_u8(eng, channels, kind):
    g = synth.Grid(8, 8)
    rr = synth.ribbon_roads(g, 48, seed=11 + channels)
    tiles = synth.host_tiles(g, channels, kind)
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    hist, nz = eng.zonal_hist_host(rr.roads, tb, rr.pairs)
    oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, gt)
    assert hist.shape == (48, channels, 256)
    assert oh.sum() > 50000
    assert np.array_equal(hist.astype(np.uint64), oh)
    assert np.array_equal(nz.astype(np.uint64), onz)
    assert onz.sum() > 0          # the all-bands-zero pixels are exercised


def test_zonal_hist_python_oracle_cross_check(eng):
    """the slow numpy/pure-Python oracle (gdal_fill + raster) on a small case"""
    g = synth.Grid(3, 3)
    rr = synth.ribbon_roads(g, 6, seed=3)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    road_of = rr.pairs.road_of_pair()
    pl = [(int(t), int(r)) for t, r in zip(rr.pairs.pair_tile, road_of)]
    oh, onz = oraster.zonal_accumulate(tiles, gt, [rr.roads.rings(r) for r in range(6)], pl)
    hist, nz = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs)
    assert np.array_equal(hist.astype(np.uint64), oh)
    assert np.array_equal(nz.astype(np.uint64), onz)


def test_zonal_hist_class_score(eng):
    g = synth.Grid(6, 6)
    rr = synth.ribbon_roads(g, 32, seed=21)
    tiles = synth.host_tiles(g, 2, "class_score")
    gt = g.transforms()
    hist, nz = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs, hist_mode="class_score")
    oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, gt, joint=True)
    assert hist.shape == (32, 3, 256)
    assert np.array_equal(hist.astype(np.uint64), oh)


@pytest.mark.parametrize("f32", [False, True])
def test_zonal_hist_u16_rescale(eng, f32):
    from proj_roadsurf_b200.engine import scale_params
    g = synth.Grid(4, 4)
    rr = synth.ribbon_roads(g, 16, seed=31)
    tiles = synth.host_tiles(g, 4, dtype=np.uint16)
    gt = g.transforms()
    smin, smax = [150.0, 300.0, 300.0, 300.0], [9000.0, 6000.0, 6000.0, 6000.0]   # NIR range, then RGB range x3
    k, off = scale_params(smin, smax, f32)
    hist, nz = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs, rescale=(k, off, f32))
    oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, gt, scale_k=k, scale_off=off, rescale_f32=f32)
    assert np.array_equal(hist.astype(np.uint64), oh)
    assert np.array_equal(nz.astype(np.uint64), onz)
    # the numpy restatement of gdal.Translate scaleParams agrees with the C one
    px = tiles.reshape(-1, 4)[:5000]
    a = oraster.rescale_u16_to_u8(px, smin, smax, f32)
    ft = np.float32 if f32 else np.float64
    v = np.clip(px.astype(ft) * k.astype(ft) + off.astype(ft), 0, 255)
    assert np.array_equal(a, (v + ft(0.5)).astype(np.int32).astype(np.uint8))


def test_zonal_long_edge_lists_and_big_tiles(eng):
    """config 5 shape: 1024 px tiles, wide polygons with thousands of vertices and holes."""
    g = synth.Grid(2, 2, size=1024)
    rng = np.random.default_rng(9)
    X0, Y1 = g.origin
    geoms = []
    for i in range(6):
        n = [300, 1200, 2600, 5000, 130, 9000][i]
        ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
        rad = g.span * (0.35 + 0.25 * rng.random()) * (1 + 0.15 * np.sin(ang * (7 + i)) + 0.02 * rng.random(n))
        c = np.array([X0 + g.span * (0.6 + 0.8 * rng.random()), Y1 - g.span * (0.6 + 0.8 * rng.random())])
        ext = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        ext = np.concatenate([ext, ext[:1]])
        rings = [ext]
        for hq in range(i % 4):
            hc = c + g.span * 0.12 * np.array([np.cos(hq * 2.1), np.sin(hq * 2.1)])
            hr = g.span * 0.04
            rings.append(ring((hc[0] - hr, hc[1] - hr), (hc[0] - hr, hc[1] + hr), (hc[0] + hr, hc[1] + hr), (hc[0] + hr, hc[1] - hr)))
        geoms.append(rings)
    roads = RoadSet.from_geometries(geoms)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    pairs = pairs_by_bbox(roads, tb)
    assert pairs.n_pairs >= 12
    hist, nz = eng.zonal_hist_host(roads, tb, pairs)
    oh, onz = oracle_hist(roads, pairs, tiles, gt)
    assert oh[:, 0].sum() > 500000
    assert np.array_equal(hist.astype(np.uint64), oh)
    assert np.array_equal(nz.astype(np.uint64), onz)


def test_zonal_edge_cases(eng):
    g = synth.Grid(2, 2)
    gt = g.transforms()
    tiles = synth.host_tiles(g, 3)
    tb = TileBatch.from_arrays(tiles, gt)
    X0, Y1 = g.origin
    far = ring((X0 - 900, Y1 + 900), (X0 - 800, Y1 + 900), (X0 - 800, Y1 + 800), (X0 - 900, Y1 + 800))
    inside = ring((X0 + 10, Y1 - 10), (X0 + 60, Y1 - 10), (X0 + 60, Y1 - 50), (X0 + 10, Y1 - 50))
    roads = RoadSet.from_geometries([[far], [inside], [inside], []])
    # road 0 paired with tiles it misses; road 1 paired; road 2 has no pair; road 3 has no ring
    pairs = PairList.from_pairs(4, [0, 0, 1, 1, 3], [0, 3, 0, 1, 0])
    hist, nz = eng.zonal_hist_host(roads, tb, pairs)
    oh, onz = oracle_hist(roads, pairs, tiles, gt)
    assert np.array_equal(hist.astype(np.uint64), oh)
    assert hist[0].sum() == 0 and hist[2].sum() == 0 and hist[3].sum() == 0 and hist[1].sum() > 0
    # no roads at all / no pairs at all
    empty = RoadSet.from_geometries([])
    h0, z0 = eng.zonal_hist_host(empty, tb, PairList.from_pairs(0, [], []))
    assert h0.shape == (0, 3, 256)
    h1, z1 = eng.zonal_hist_host(roads, tb, PairList.from_pairs(4, [], []))
    assert h1.sum() == 0 and z1.sum() == 0


def test_zonal_road_slots(eng):
    """road_slot scatters the roads' rows into a larger table (the boundary table of a tile shard)"""
    g = synth.Grid(4, 4)
    rr = synth.ribbon_roads(g, 20, seed=77)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    slot = (np.arange(20) + 5).astype(np.int32)          # a permutation-free shift into a larger table
    hist, nz = eng.zonal_hist_host(rr.roads, tb, rr.pairs, road_slot=slot, n_slots=32)
    oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, gt)
    assert hist.shape == (32, 3, 256)
    assert np.array_equal(hist[5:25].astype(np.uint64), oh)
    assert hist[:5].sum() == 0 and hist[25:].sum() == 0


def test_zonal_rotated_transform_is_rejected(eng):
    from proj_roadsurf_b200._native import NativeError
    roads = RoadSet.from_geometries([[ring((1, 1), (5, 1), (5, 5), (1, 5))]])
    pairs = PairList.from_pairs(1, [0], [0])
    tiles = np.zeros((1, 8, 8, 3), np.uint8)
    gt = np.array([[1.0, 0.2, 0, 0.1, 1.0, 0]])
    with pytest.raises(NativeError) as ei:
        eng.zonal_hist_host(roads, TileBatch.from_arrays(tiles, gt), pairs)
    assert ei.value.status == -4


def test_device_family_matches_host_family(eng):
    import torch
    g = synth.Grid(8, 8)
    rr = synth.ribbon_roads(g, 48, seed=5)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    hist, nz = eng.zonal_hist_host(rr.roads, tb, rr.pairs)
    dr, dp, dt = eng.upload_roads(rr.roads), eng.upload_pairs(rr.pairs), eng.upload_tiles(tb)
    dh, dz = eng.zonal_hist_dev(dr, dt, dp)
    torch.cuda.synchronize()
    assert np.array_equal(dh.cpu().numpy().view(np.uint32), hist)
    assert np.array_equal(dz.cpu().numpy().view(np.uint32), nz)
    # determinism: integer outputs identical run to run
    dh2, dz2 = eng.zonal_hist_dev(dr, dt, dp)
    assert torch.equal(dh, dh2) and torch.equal(dz, dz2)


def test_synth_tiles_device_generator_feeds_both_sides(eng):
    """on-device synthetic tiles copied back are what the oracle sees: full-size-style check"""
    import torch
    g = synth.Grid(16, 8)
    rr = synth.ribbon_roads(g, 64, seed=13)
    gt = g.transforms()
    dt = eng.synth_tiles_dev(g.keys(), 256, 256, 3, kind=1, gt=gt)
    dh, dz = eng.zonal_hist_dev(eng.upload_roads(rr.roads), dt, eng.upload_pairs(rr.pairs))
    torch.cuda.synchronize()
    tiles = dt.pixels.cpu().numpy()
    oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, gt, threads=4)
    assert np.array_equal(dh.cpu().numpy().view(np.uint32).astype(np.uint64), oh)
    assert np.array_equal(dz.cpu().numpy().view(np.uint32).astype(np.uint64), onz)


# ------------------------------------------------------------------------------------------
# statistics from histograms
# ------------------------------------------------------------------------------------------
def _rand_hists(rng, R, C):
    h = np.zeros((R, C, 256), np.uint32)
    for r in range(R):
        for c in range(C):
            mode = rng.integers(0, 5)
            if mode == 0:
                continue                                   # empty
            if mode == 1:
                h[r, c, rng.integers(0, 256)] = 1          # n = 1: std NaN with ddof 1
            elif mode == 2:
                h[r, c, rng.integers(0, 256, 2)] += 1
            elif mode == 3:
                h[r, c] = rng.integers(0, 50, 256)
            else:
                k = rng.integers(100, 120)
                h[r, c, k - 5:k + 5] = rng.integers(0, 100000, 10)
    return h


@pytest.mark.parametrize("ddof", [0, 1])
def test_finalize_stats_raw(eng, ddof):
    rng = np.random.default_rng(1)
    h = _rand_hists(rng, 40, 3)
    pct = [0.0, 5.0, 25.0, 50.0, 90.0, 99.5, 100.0]
    out = eng.finalize_stats_host(h, None, nodata_mode="raw", ddof=ddof, percentiles=pct)
    from proj_roadsurf_b200._native import STAT_COLS
    col = {k: i for i, k in enumerate(STAT_COLS)}
    for r in range(40):
        for c in range(3):
            s = ostats.stats_from_hist(h[r, c], ddof=ddof, percentiles=pct)
            o = out[r, c]
            assert o[col["count"]] == s["count"]
            if s["count"] == 0:
                assert np.isnan(o[1:]).all()
                continue
            assert o[col["min"]] == s["min"] and o[col["max"]] == s["max"]          # bit-exact integers
            assert o[col["median"]] == s["median"]
            v = ostats.expand_hist(h[r, c])
            assert o[col["sum"]] == v.sum() and o[col["sumsq"]] == (v * v).sum()
            np.testing.assert_allclose(o[col["mean"]], s["mean"], rtol=RTOL)
            if np.isnan(s["std"]):
                assert np.isnan(o[col["std"]])
            else:
                np.testing.assert_allclose(o[col["std"]], s["std"], rtol=RTOL, atol=1e-9)
                np.testing.assert_allclose(o[col["margin"]], 2 * s["std"] / np.sqrt(s["count"]), rtol=RTOL, atol=1e-9)
            for i, q in enumerate(pct):
                np.testing.assert_allclose(o[len(STAT_COLS) + i], s[f"percentile_{q:g}"], rtol=1e-12)


@pytest.mark.parametrize("mode,omode", [("none", "N"), ("zero", "Z")])
def test_finalize_stats_nodata_conventions(eng, mode, omode):
    g = synth.Grid(4, 4)
    rr = synth.ribbon_roads(g, 24, seed=8)
    tiles = synth.host_tiles(g, 3)
    tiles[..., 1][tiles[..., 0] < 40] = 0          # bands with different numbers of zeros
    gt = g.transforms()
    for t in range(g.n_tiles):                     # ... and different ones per tile: the zero padding is per (road, tile) call
        tiles[t, :, :, t % 3][tiles[t, :, :, (t + 1) % 3] > 200 - 10 * t] = 0
    hist, nz, mz = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs, want_min_zero=True)
    oh, onz, omz = cport.zonal_accumulate(rr.roads.xy, rr.roads.ring_off, rr.roads.road_ring_off, rr.pairs.road_pair_off,
                                          rr.pairs.pair_tile, tiles, gt, want_min_zero=True)
    assert np.array_equal(hist.astype(np.uint64), oh) and np.array_equal(nz.astype(np.uint64), onz)
    assert np.array_equal(mz.astype(np.uint64), omz)
    assert (omz != onz).any()                      # the two corrections differ on this input
    out = eng.finalize_stats_host(hist, mz if mode == "zero" else nz, nodata_mode=mode, ddof=1)
    adj = ostats.apply_nodata_convention(hist, nz, omode, min_zero=omz)
    for r in range(24):
        for c in range(3):
            s = ostats.stats_from_hist(adj[r, c], ddof=1)
            assert out[r, c, 0] == s["count"]
            if s["count"] > 1:
                assert out[r, c, 1] == s["min"] and out[r, c, 2] == s["max"] and out[r, c, 7] == s["median"]
                np.testing.assert_allclose(out[r, c, 5], s["mean"], rtol=RTOL)
                np.testing.assert_allclose(out[r, c, 6], s["std"], rtol=RTOL)


# ------------------------------------------------------------------------------------------
# vote + confusion + F1
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rule", ["count", "score"])
def test_vote_cutoffs_in_any_order(eng, rule):
    """vote_kernel visits the cut-offs in descending order while it walks the score bins down: unsorted, repeated and
    out-of-range cut-offs (above 255: nothing counts; 0 and below: everything counts) land in the caller's positions"""
    g = synth.Grid(6, 6)
    rr = synth.ribbon_roads(g, 150, seed=43)             # more roads than a block of the kernel: a partial last block
    tiles = synth.host_tiles(g, 2, "class_score")
    jh, _ = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, g.transforms()), rr.pairs, hist_mode="class_score")
    cuts = np.array([128, 3, 255, 0, 300, 128, -5, 254, 17, 256, 1], np.int32)
    cover, scores, conf, _ = eng.vote_metrics_host(jh, rr.gt_class, cuts, rule=rule, min_area_frac=0.05)
    for i, c in enumerate(cuts):
        ocov, ia, inn, _ = ovote.raster_vote(jh, int(c), rule, 0.05)
        assert np.array_equal(cover[i], ocov.astype(np.int8)), (i, c)
        assert np.array_equal(conf[i], ovote.confusion(ocov, rr.gt_class)), (i, c)
        np.testing.assert_allclose(scores[i, :, 0], ia, rtol=1e-12)
        np.testing.assert_allclose(scores[i, :, 1], inn, rtol=1e-12)


@pytest.mark.parametrize("rule", ["count", "score"])
@pytest.mark.parametrize("min_area_frac", [0.0, 0.05])
def test_vote_metrics_sweep(eng, rule, min_area_frac):
    g = synth.Grid(8, 8)
    rr = synth.ribbon_roads(g, 96, seed=41)
    tiles = synth.host_tiles(g, 2, "class_score")
    gt = g.transforms()
    jh, _ = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs, hist_mode="class_score")
    gtc = rr.gt_class.copy()
    gtc[::17] = -1                                  # roads outside the ground truth are skipped
    cuts = ovote.score_cutoffs()
    cover, scores, conf, met = eng.vote_metrics_host(jh, gtc, cuts, rule=rule, min_area_frac=min_area_frac)
    keep = gtc >= 0
    from proj_roadsurf_b200._native import METRIC_COLS
    for i, c in enumerate(cuts):
        ocov, ia, inn, diff = ovote.raster_vote(jh, int(c), rule, min_area_frac)
        assert np.array_equal(cover[i], ocov.astype(np.int8)), (i, c)
        m = ovote.confusion(ocov[keep], gtc[keep])
        assert np.array_equal(conf[i], m)
        om = ovote.metrics_from_confusion(m)
        np.testing.assert_allclose(scores[i, :, 0], ia, rtol=1e-12)
        np.testing.assert_allclose(scores[i, :, 1], inn, rtol=1e-12)
        for j, name in enumerate(METRIC_COLS):
            key = {"P0": "P_0", "R0": "R_0", "F0": "f1_0", "P1": "P_1", "R1": "R_1", "F1": "f1_1"}.get(name, name)
            np.testing.assert_allclose(met[i, j], om[key], rtol=RTOL, atol=1e-15)
    assert len(set(cover[0].tolist())) >= 2


# ------------------------------------------------------------------------------------------
# tile sharding: the ranks of a multi-GPU run, emulated one after the other on this GPU
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4])
def test_tile_sharded_accumulation_equals_single_shot(eng, world):
    from proj_roadsurf_b200.distributed import global_rows, plan_shards
    g = synth.Grid(8, 12)
    rr = synth.ribbon_roads(g, 90, seed=61)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    full_h, full_z = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs)
    shards = plan_shards(rr.roads, rr.pairs, g.n_tiles, world)
    nb = shards[0].n_boundary
    assert nb > 0
    boundary_h = np.zeros((nb, 3, 256), np.uint32)
    boundary_z = np.zeros(nb, np.uint32)
    outs = []
    for sh in shards:
        tb = TileBatch.from_arrays(tiles[sh.tile_lo:sh.tile_hi], gt[sh.tile_lo:sh.tile_hi])
        h, z = eng.zonal_hist_host(sh.roads, tb, sh.pairs, road_slot=sh.slot, n_slots=sh.n_rows)
        boundary_h += h[sh.n_own:]                         # what the all-reduce(SUM) of the boundary table does
        boundary_z += z[sh.n_own:]
        outs.append((h, z))
    seen = np.zeros(90, int)
    for sh, (h, z) in zip(shards, outs):
        rows = global_rows(sh)
        assert np.array_equal(h[:sh.n_own], full_h[rows[:sh.n_own]])
        assert np.array_equal(z[:sh.n_own], full_z[rows[:sh.n_own]])
        seen[rows[:sh.n_own]] += 1
    assert np.array_equal(boundary_h, full_h[shards[0].boundary_global])
    assert np.array_equal(boundary_z, full_z[shards[0].boundary_global])
    seen[shards[0].boundary_global] += 1
    assert np.array_equal(seen > 0, np.diff(rr.pairs.road_pair_off) > 0) and seen.max() == 1


# ------------------------------------------------------------------------------------------
# GPU broad phase
# ------------------------------------------------------------------------------------------
def test_pairs_bbox_gpu_equals_host_broad_phase(eng):
    g = synth.Grid(24, 17)
    rr = synth.ribbon_roads(g, 700, seed=71)
    gt = g.transforms()
    keep = np.ones(g.n_tiles, bool)
    keep[::7] = False                                       # a lattice with missing tiles
    tb = TileBatch(None, gt[keep], 256, 256, 3)
    host = pairs_by_bbox(rr.roads, tb)
    dev = eng.pairs_bbox_host(rr.roads, tb)
    assert np.array_equal(dev.road_pair_off, host.road_pair_off)
    assert np.array_equal(dev.pair_tile, host.pair_tile)
    assert dev.n_pairs > 3000
    # and it loses no pixel-carrying pair of the (tighter) generator list
    full = TileBatch(None, gt, 256, 256, 3)
    dev_full = eng.pairs_bbox_host(rr.roads, full)
    have = set(zip(dev_full.road_of_pair().tolist(), dev_full.pair_tile.tolist()))
    assert set(zip(rr.pairs.road_of_pair().tolist(), rr.pairs.pair_tile.tolist())) <= have
    # results through either pair list are identical
    tiles = synth.host_tiles(synth.Grid(6, 6), 3)
    g2 = synth.Grid(6, 6)
    rr2 = synth.ribbon_roads(g2, 40, seed=72)
    tb2 = TileBatch.from_arrays(tiles, g2.transforms())
    h1, z1 = eng.zonal_hist_host(rr2.roads, tb2, rr2.pairs)
    h2, z2 = eng.zonal_hist_host(rr2.roads, tb2, eng.pairs_bbox_host(rr2.roads, tb2))
    assert np.array_equal(h1, h2) and np.array_equal(z1, z2)


# ------------------------------------------------------------------------------------------
# BASELINE.json's full single-GPU size: size-independent properties + oracle on a sample of roads
# ------------------------------------------------------------------------------------------
def test_full_size_properties(eng):
    import torch
    free, _ = torch.cuda.mem_get_info()
    nx = ny = 512 if free > 70e9 else 128                    # 262 144 tiles (51.5 GB) when the GPU is free
    g = synth.Grid(nx, ny)
    R = nx * ny // 2
    rr = synth.ribbon_roads(g, R)
    gt = g.transforms()
    dt = eng.synth_tiles_dev(g.keys(), 256, 256, 3, kind=0, gt=gt)
    dr, dp = eng.upload_roads(rr.roads), eng.upload_pairs(rr.pairs)
    dh, dz = eng.zonal_hist_dev(dr, dt, dp)
    torch.cuda.synchronize()
    h = dh.view(torch.int32)
    # (1) every band histogram of a road sums to the same pixel count
    cnt = h.sum(dim=2, dtype=torch.int64)
    assert bool((cnt[:, 0] == cnt[:, 1]).all()) and bool((cnt[:, 0] == cnt[:, 2]).all())
    total_px = int(cnt[:, 0].sum().item())
    assert 0.03 < total_px / (g.n_tiles * 65536.0) < 0.15        # ribbons cover a few per cent of the pixels
    assert int(dz.sum().item()) > 0 and bool((dz.to(torch.int64) <= cnt[:, 0]).all())
    # (2) determinism: bit-identical accumulators run to run (scheduling differs, integers do not)
    dh2, dz2 = eng.zonal_hist_dev(dr, dt, dp)
    assert torch.equal(dh, dh2) and torch.equal(dz, dz2)
    # (3) invariance to road order: reversing the roads reverses the rows
    rev = np.arange(R)[::-1].copy()
    roads_r, pairs_r = rr.roads.subset(rev), rr.pairs.take_roads(rev)
    dh3, dz3 = eng.zonal_hist_dev(eng.upload_roads(roads_r), dt, eng.upload_pairs(pairs_r))
    assert torch.equal(dh3.flip(0), dh) and torch.equal(dz3.flip(0), dz)
    # (4) statistics consistent with the histograms: count, sum and min/max of band 0
    st = eng.finalize_stats_dev(dh, dz, nodata_mode="raw", ddof=1)
    vals = torch.arange(256, device=h.device, dtype=torch.int64)
    s1 = (h[:, 0].to(torch.int64) * vals).sum(dim=1)
    assert torch.equal(st[:, 0, 0].to(torch.int64), cnt[:, 0]) and torch.equal(st[:, 0, 3].to(torch.int64), s1)
    # (5) the CPU oracle on a sample of roads (their tiles are copied back)
    rng = np.random.default_rng(5)
    sample = np.sort(rng.choice(R, 48, replace=False))
    sub_pairs = rr.pairs.take_roads(sample)
    tiles_needed, inv = np.unique(sub_pairs.pair_tile, return_inverse=True)
    host_tiles = dt.pixels[torch.from_numpy(tiles_needed.astype(np.int64)).to(h.device)].cpu().numpy()
    sub_roads = rr.roads.subset(sample)
    oh, onz = cport.zonal_accumulate(sub_roads.xy, sub_roads.ring_off, sub_roads.road_ring_off, sub_pairs.road_pair_off,
                                     inv.astype(np.int32), host_tiles, gt[tiles_needed], threads=4)
    got = dh[torch.from_numpy(sample).to(h.device)].cpu().numpy().view(np.uint32)
    assert np.array_equal(got.astype(np.uint64), oh)
    assert np.array_equal(dz[torch.from_numpy(sample).to(h.device)].cpu().numpy().view(np.uint32).astype(np.uint64), onz)
    assert oh.sum() > 100000


# ------------------------------------------------------------------------------------------
# 16 -> 8 bit rescale as a materialising pass (tif2cog.py:260-270) + band selection (config_stats.yaml:39)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("f32", [False, True])
def test_rescale_u16_matches_gdal_translate_restatement(eng, f32):
    import torch
    rng = np.random.default_rng(12)
    src = np.clip(rng.lognormal(7.5, 0.9, (3, 70, 65, 4)), 0, 65535).astype(np.uint16)      # NIR, R, G, B
    src[0, :3, :3] = [0, 65535, 150, 9000]
    nir, rgb = (150.0, 9000.0), (300.0, 6000.0)
    # output bands R, G, B, NIR = source bands 1, 2, 3, 0 (bidx=2,3,4,1), each with its own range
    bidx = [1, 2, 3, 0]
    smin, smax = [rgb[0]] * 3 + [nir[0]], [rgb[1]] * 3 + [nir[1]]
    got = eng.rescale_u16_host(src, smin, smax, bidx=bidx, f32=f32)
    exp = oraster.rescale_u16_to_u8(src[..., bidx], smin, smax, f32)
    assert got.shape == (3, 70, 65, 4) and np.array_equal(got, exp)
    assert len(np.unique(got)) > 200
    # device form, identity band order, and a 3-band output (generic kernel)
    d = torch.from_numpy(src.view(np.int16)).cuda()
    out = eng.rescale_u16_dev(d, [nir[0]] + [rgb[0]] * 3, [nir[1]] + [rgb[1]] * 3, f32=f32)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), oraster.rescale_u16_to_u8(src, [nir[0]] + [rgb[0]] * 3, [nir[1]] + [rgb[1]] * 3, f32))
    got3 = eng.rescale_u16_host(src, smin[:3], smax[:3], bidx=[1, 2, 3], f32=f32)
    assert np.array_equal(got3, exp[..., :3])
    # rescale then 8-bit statistics == fused RS_U16 statistics
    g = synth.Grid(2, 2)
    rr = synth.ribbon_roads(g, 6, seed=2)
    t16 = synth.host_tiles(g, 4, dtype=np.uint16)
    from proj_roadsurf_b200.engine import scale_params
    k, off = scale_params([nir[0]] + [rgb[0]] * 3, [nir[1]] + [rgb[1]] * 3, f32)
    h_fused, _ = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(t16, g.transforms()), rr.pairs, rescale=(k, off, f32))
    t8 = eng.rescale_u16_host(t16, [nir[0]] + [rgb[0]] * 3, [nir[1]] + [rgb[1]] * 3, f32=f32)
    h_two, _ = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(t8, g.transforms()), rr.pairs)
    assert np.array_equal(h_fused, h_two)


# ------------------------------------------------------------------------------------------
# degenerate geometry, extreme shapes, size limits
# ------------------------------------------------------------------------------------------
def test_degenerate_and_self_intersecting_geometry(eng):
    H, W = 33, 45                                    # odd sizes: the generic (unaligned) pixel path
    shapes = [
        [ring((3, 3), (20, 3))[:2]],                                  # 2 vertices: no area
        [np.array([[5.0, 5.0]])],                                     # a single vertex
        [ring((2, 2), (30, 2), (30, 2), (30, 20), (2, 20), (2, 20))], # repeated vertices
        [ring((2, 2), (40, 30), (40, 2), (2, 30))],                   # bow-tie (self-intersecting): even-odd
        [ring((1, 1), (10, 1), (20, 1), (20, 10), (20, 25), (1, 25))],            # collinear vertices
        [ring((0.5, 0.5), (44.5, 0.5), (44.5, 32.5), (0.5, 32.5))],  # edges through pixel centres on the raster border
        [ring((-1e7, -1e7), (1e7, -1e7), (1e7, 1e7), (-1e7, 1e7))],  # enormous
        [ring((10, 10), (10.2, 10), (10.2, 30), (10, 30))],          # sliver
        [ring((5, 5), (25, 5), (25, 25), (5, 25)), ring((5, 5), (25, 5), (25, 25), (5, 25))],   # identical rings cancel
        [ring(*[(22 + 18 * np.cos(a), 16 + 14 * np.sin(a)) for a in np.linspace(0, 4 * np.pi, 41)[:-1]])],   # winds twice
    ]
    roads = RoadSet.from_geometries(shapes)
    n = roads.n_roads
    pairs = PairList.from_pairs(n, np.arange(n), np.zeros(n, int))
    gt = np.array([[1.0, 0, 0, 0, 1.0, 0]])
    for window in ("full", "crop"):
        masks = eng.rasterize_pairs_host(roads, gt, H, W, pairs, window=window)
        for i, rings in enumerate(shapes):
            exp = cport.rasterize(rings, (H, W)) if window == "full" else cport.pair_mask_full(gt[0], rings, W, H)
            assert np.array_equal(masks[i], exp), (window, i)
    rng = np.random.default_rng(3)
    tiles = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
    h, z = eng.zonal_hist_host(roads, TileBatch.from_arrays(tiles, gt), pairs)
    oh, oz = oracle_hist(roads, pairs, tiles, gt)
    assert np.array_equal(h.astype(np.uint64), oh) and np.array_equal(z.astype(np.uint64), oz)


def test_nan_coordinates_do_not_fault(eng):
    bad = ring((2, 2), (float("nan"), 5), (20, 20), (2, 20))
    roads = RoadSet.from_geometries([[bad], [ring((1, 1), (9, 1), (9, 9), (1, 9))]])
    pairs = PairList.from_pairs(2, [0, 1], [0, 0])
    tiles = np.full((1, 16, 16, 3), 7, np.uint8)
    h, _ = eng.zonal_hist_host(roads, TileBatch.from_arrays(tiles, np.array([[1.0, 0, 0, 0, 1.0, 0]])), pairs)
    assert h[1, 0, 7] == 64                          # the valid road is unaffected; the NaN road may hold anything finite
    assert h[0].sum() <= 3 * 256


def test_width_limits(eng):
    from proj_roadsurf_b200._native import NativeError
    roads = RoadSet.from_geometries([[ring((100, 1), (1900, 1), (1900, 3), (100, 3))]])
    pairs = PairList.from_pairs(1, [0], [0])
    wide = np.zeros((1, 4, 2048, 3), np.uint8)
    wide[0, :, :, 1] = (np.arange(2048) % 251).astype(np.uint8)[None, :]
    gt = np.array([[1.0, 0, 0, 0, 1.0, 0]])
    h, _ = eng.zonal_hist_host(roads, TileBatch.from_arrays(wide, gt), pairs)          # the widest supported tile
    oh, _ = oracle_hist(roads, pairs, wide, gt)
    assert np.array_equal(h.astype(np.uint64), oh) and oh[0, 0].sum() == 2 * 1800
    # the raster may be wider than 2048 px as long as every road's window is not
    big = np.zeros((1, 4, 6000, 3), np.uint8)
    big[0, :, :, 2] = (np.arange(6000) % 199).astype(np.uint8)[None, :]
    r3 = RoadSet.from_geometries([[ring((3100, 1), (4900, 1), (4900, 3), (3100, 3))], [ring((10.5, 0.2), (600.2, 0.2), (600.2, 3.9), (10.5, 3.9))]])
    p3 = PairList.from_pairs(2, [0, 1], [0, 0])
    h3, _ = eng.zonal_hist_host(r3, TileBatch.from_arrays(big, gt), p3)
    oh3, _ = oracle_hist(r3, p3, big, gt)
    assert np.array_equal(h3.astype(np.uint64), oh3) and oh3[0, 0].sum() == 2 * 1800
    r4 = RoadSet.from_geometries([[ring((100, 1), (2300, 1), (2300, 3), (100, 3))]])
    with pytest.raises(NativeError) as ei:
        eng.zonal_hist_host(r4, TileBatch.from_arrays(big, gt), pairs)
    assert ei.value.status == -6                     # RS_ERR_UNSUPPORTED (a 2200 px wide window), never a silent wrong answer
    tall = np.zeros((1, 5000, 8, 1), np.uint8)       # tall rasters are fine
    tall[0, :, :, 0] = (np.arange(5000) % 200).astype(np.uint8)[:, None]
    r2 = RoadSet.from_geometries([[ring((1, 10), (7, 10), (7, 4990), (1, 4990))]])
    h2, _ = eng.zonal_hist_host(r2, TileBatch.from_arrays(tall, gt), pairs)
    oh2, _ = oracle_hist(r2, pairs, tall, gt)
    assert np.array_equal(h2.astype(np.uint64), oh2) and oh2.sum() == 6 * 4980


def test_wide_polygons_config5(eng):
    """BASELINE configs[4]: 1024 px tiles, wide polygons with up to ~10 k vertices and 1-8 holes"""
    g = synth.Grid(4, 4, size=1024)
    wp = synth.wide_polygons(g, 10, seed=9)
    nv = np.diff(wp.roads.ring_off[wp.roads.road_ring_off])
    assert nv.max() > 3000 and wp.roads.n_rings > wp.roads.n_roads
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    h, z = eng.zonal_hist_host(wp.roads, TileBatch.from_arrays(tiles, gt), wp.pairs)
    oh, oz = oracle_hist(wp.roads, wp.pairs, tiles, gt, threads=4)
    assert oh[:, 0].sum() > 2_000_000
    assert np.array_equal(h.astype(np.uint64), oh) and np.array_equal(z.astype(np.uint64), oz)


def test_streaming_from_host_equals_resident(eng):
    """rs_zonal_stats_stream_host: tiles streamed through two device buffers, histograms accumulated on the device"""
    g = synth.Grid(7, 5)
    rr = synth.ribbon_roads(g, 40, seed=81)
    tiles = synth.host_tiles(g, 3)
    tb = TileBatch.from_arrays(tiles, g.transforms())
    ref = eng.zonal_stats_host(rr.roads, tb, rr.pairs, nodata_mode="none", ddof=1, percentiles=(10.0, 90.0), want_hist=True)
    for per in (1, 4, 9, 35, 100):
        got = eng.zonal_stats_host(rr.roads, tb, rr.pairs, nodata_mode="none", ddof=1, percentiles=(10.0, 90.0), want_hist=True,
                                   tiles_per_chunk=per)
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]), per
        assert np.array_equal(got[0], ref[0], equal_nan=True), per
    oh, _ = oracle_hist(rr.roads, rr.pairs, tiles, g.transforms())
    assert np.array_equal(ref[1].astype(np.uint64), oh)


def test_pageable_sources_through_the_slot_ring(monkeypatch):
    """rs_hostcopy.cu: copies of >= 8 MiB out of pageable memory go through page-locked slots filled by several host threads;
    results equal the plain cudaMemcpyAsync path (RS_STAGE_COPY=0) for the whole-batch copy and for streamed chunks whose sizes
    are not multiples of the slot size, with 1, 3 and the default number of copy threads"""
    from proj_roadsurf_b200.engine import Engine
    g = synth.Grid(16, 12)
    rr = synth.ribbon_roads(g, 60, seed=83)
    tiles = synth.host_tiles(g, 3)                         # 37.7 MB of ordinary numpy memory
    tb = TileBatch.from_arrays(tiles, g.transforms())
    monkeypatch.setenv("RS_STAGE_COPY", "0")
    e0 = Engine(0)
    ref = e0.zonal_stats_host(rr.roads, tb, rr.pairs, want_hist=True, mapped=False)
    monkeypatch.delenv("RS_STAGE_COPY")
    for threads in ("1", "3", None):
        if threads is None:
            monkeypatch.delenv("RS_STAGE_THREADS", raising=False)
        else:
            monkeypatch.setenv("RS_STAGE_THREADS", threads)
        e = Engine(0)                                      # the copy threads are created with the context's first staged copy
        for per in (None, 100, 50, 192):
            got = e.zonal_stats_host(rr.roads, tb, rr.pairs, want_hist=True, mapped=False, tiles_per_chunk=per)
            assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]), (threads, per)
            assert np.array_equal(got[0], ref[0], equal_nan=True), (threads, per)
        del e
    oh, _ = oracle_hist(rr.roads, rr.pairs, tiles, g.transforms(), threads=4)
    assert np.array_equal(ref[1].astype(np.uint64), oh)


def test_pairs_per_item_do_not_change_results(eng, monkeypatch):
    """small launches get shorter work items (rs_zonal.cu, launch_impl: pairs per item from the number of pairs per team); the
    rows of roads split over several items are accumulated with atomics -- same histograms, n_allzero and min_zero"""
    g = synth.Grid(6, 6)
    rr = synth.ribbon_roads(g, 50, seed=84)
    tiles = synth.host_tiles(g, 3)
    tiles[tiles < 40] = 0                                  # plenty of zeros: n_allzero and min_zero are exercised
    tb = TileBatch.from_arrays(tiles, g.transforms())
    res = {}
    for ppi in ("8", "3", "1", None):
        if ppi is None:
            monkeypatch.delenv("RS_ZONAL_PPI", raising=False)
        else:
            monkeypatch.setenv("RS_ZONAL_PPI", ppi)
        res[ppi] = eng.zonal_hist_host(rr.roads, tb, rr.pairs, want_min_zero=True)
    for ppi in ("3", "1", None):
        for a, b in zip(res[ppi], res["8"]):
            assert np.array_equal(a, b), ppi
    oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, g.transforms())
    assert np.array_equal(res["1"][0].astype(np.uint64), oh) and np.array_equal(res["1"][1].astype(np.uint64), onz)


def test_mapped_host_tiles_equal_resident(eng):
    """rs_zonal_stats_mapped_host: the kernel reads page-locked host tiles in place; pageable memory is refused"""
    import torch
    from proj_roadsurf_b200 import _native as N
    g = synth.Grid(7, 5)
    rr = synth.ribbon_roads(g, 40, seed=82)
    tiles = synth.host_tiles(g, 3)
    ref = eng.zonal_stats_host(rr.roads, TileBatch.from_arrays(tiles, g.transforms()), rr.pairs, percentiles=(25.0,), want_hist=True)
    pinned = torch.empty(tiles.shape, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(torch.from_numpy(tiles))
    tb = TileBatch(pinned.numpy(), g.transforms(), tiles.shape[1], tiles.shape[2], tiles.shape[3])
    got = eng.zonal_stats_host(rr.roads, tb, rr.pairs, percentiles=(25.0,), want_hist=True, mapped=True)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    assert np.array_equal(got[0], ref[0], equal_nan=True)
    with pytest.raises(Exception, match="RS_ERR_NOT_PINNED"):
        eng.zonal_stats_host(rr.roads, TileBatch.from_arrays(tiles.copy(), g.transforms()), rr.pairs, mapped=True)
    for ch in (1, 2, 4):                     # the other band counts read host memory through the resident load path
        t2 = synth.host_tiles(g, ch)
        ref2 = eng.zonal_stats_host(rr.roads, TileBatch.from_arrays(t2, g.transforms()), rr.pairs, want_hist=True, mapped=False)
        pin2 = torch.empty(t2.shape, dtype=torch.uint8, pin_memory=True)
        pin2.copy_(torch.from_numpy(t2))
        got2 = eng.zonal_stats_host(rr.roads, TileBatch(pin2.numpy(), g.transforms(), t2.shape[1], t2.shape[2], ch), rr.pairs,
                                    want_hist=True, mapped=True)
        assert np.array_equal(got2[1], ref2[1]) and np.array_equal(got2[2], ref2[2]), ch


def test_pin_host_makes_numpy_tiles_readable_in_place(eng):
    """rs_host_register: a plain numpy tile buffer is page-locked in place, then read by the kernel without a copy; the
    default (mapped=None) picks the transport from the buffer"""
    g = synth.Grid(6, 4)
    rr = synth.ribbon_roads(g, 30, seed=83)
    tiles = synth.host_tiles(g, 3)
    tb = TileBatch.from_arrays(tiles, g.transforms())
    ref = eng.zonal_stats_host(rr.roads, tb, rr.pairs, want_hist=True, mapped=False)
    auto = eng.zonal_stats_host(rr.roads, tb, rr.pairs, want_hist=True)                 # pageable: copied
    assert np.array_equal(auto[1], ref[1]) and np.array_equal(auto[0], ref[0], equal_nan=True)
    eng.pin_host(tb.pixels)
    try:
        for kw in (dict(mapped=True), dict()):
            got = eng.zonal_stats_host(rr.roads, tb, rr.pairs, want_hist=True, **kw)
            assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
            assert np.array_equal(got[0], ref[0], equal_nan=True)
        eng.pin_host(tb.pixels)                                                          # already page-locked: no-op
    finally:
        eng.unpin_host(tb.pixels)
    with pytest.raises(Exception, match="RS_ERR_NOT_PINNED"):
        eng.zonal_stats_host(rr.roads, tb, rr.pairs, mapped=True)


@pytest.mark.parametrize("mode", ["bands", "class_score"])
def test_two_kernel_form_equals_fused_kernel_and_oracle(eng, mode, monkeypatch):
    """rasterize -> entry pool -> accumulate (RS_ZONAL_SPLIT=1, the default for resident tiles) against the fused kernel
    (RS_ZONAL_SPLIT=0) and the C oracle: bit-identical histograms, also for roads split over several work items."""
    import torch
    g = synth.Grid(16, 16)
    rr = synth.ribbon_roads(g, 400, seed=17, max_vertices=300)
    ch, kind = (3, 0) if mode == "bands" else (2, 2)
    gt = g.transforms()
    dt = eng.synth_tiles_dev(g.keys(), 256, 256, ch, kind=kind, gt=gt)
    dr, dp = eng.upload_roads(rr.roads), eng.upload_pairs(rr.pairs)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("RS_ZONAL_SPLIT", flag)
        l0 = eng.launch_count
        h, z = eng.zonal_hist_dev(dr, dt, dp, hist_mode=mode)
        torch.cuda.synchronize()
        out[flag] = (h.cpu().numpy().view(np.uint32), z.cpu().numpy().view(np.uint32), eng.launch_count - l0)
    assert out["1"][2] == out["0"][2] + 2                 # emit + accumulate + (empty) fused overflow pass
    assert np.array_equal(out["0"][0], out["1"][0]) and np.array_equal(out["0"][1], out["1"][1])
    oh, onz = cport.zonal_accumulate(rr.roads.xy, rr.roads.ring_off, rr.roads.road_ring_off, rr.pairs.road_pair_off,
                                     rr.pairs.pair_tile, dt.pixels.cpu().numpy(), gt, joint=(mode == "class_score"))
    assert np.array_equal(out["1"][0].astype(np.uint64), oh) and np.array_equal(out["1"][1].astype(np.uint64), onz)
    assert (np.diff(rr.pairs.road_pair_off) > 8).any()    # some roads span several work items (atomic merge path)


def test_two_kernel_form_pool_overflow_falls_back_to_the_fused_kernel(eng, monkeypatch):
    """a pool that holds a fraction of the entries: the items that do not fit are redone by the fused kernel, same result"""
    import torch
    g = synth.Grid(12, 12)
    rr = synth.ribbon_roads(g, 300, seed=29)
    gt = g.transforms()
    dt = eng.synth_tiles_dev(g.keys(), 256, 256, 3, kind=0, gt=gt)
    dr, dp = eng.upload_roads(rr.roads), eng.upload_pairs(rr.pairs)
    monkeypatch.setenv("RS_ZONAL_SPLIT", "0")
    h0, z0 = eng.zonal_hist_dev(dr, dt, dp)
    monkeypatch.setenv("RS_ZONAL_SPLIT", "1")
    for units in ("0", "4096", "40960"):                  # nothing fits / one team's chunk / ten chunks
        monkeypatch.setenv("RS_ZONAL_POOL_UNITS", units)
        h1, z1 = eng.zonal_hist_dev(dr, dt, dp)
        torch.cuda.synchronize()
        assert torch.equal(h0, h1) and torch.equal(z0, z1), units


# ------------------------------------------------------------------------------------------
# wide-window kernel (rs_wide.cu): tiles of 512 .. 2048 px, lane-private histograms, TMA chunk pipeline
# ------------------------------------------------------------------------------------------
def _wide_case(size, seed, n_poly=7):
    g = synth.Grid(2, 3, size=size)
    rng = np.random.default_rng(seed)
    X0, Y1 = g.origin
    res = g.res
    geoms = []
    for i in range(n_poly):
        n = [40, 700, 3100, 260, 5200, 129, 1500][i % 7]
        ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
        rad = g.span * (0.25 + 0.3 * rng.random()) * (1 + 0.2 * np.sin(ang * (5 + i)) + 0.03 * rng.random(n))
        c = np.array([X0 + g.span * (0.3 + 1.4 * rng.random()), Y1 - g.span * (0.3 + 2.4 * rng.random())])
        ext = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        rings = [np.concatenate([ext, ext[:1]])]
        for hq in range(i % 5):
            hc = c + g.span * 0.1 * np.array([np.cos(hq * 1.7), np.sin(hq * 1.7)])
            hr = g.span * 0.03
            rings.append(ring((hc[0] - hr, hc[1] - hr), (hc[0] - hr, hc[1] + hr), (hc[0] + hr, hc[1] + hr), (hc[0] + hr, hc[1] - hr)))
        geoms.append(rings)
    # rectangles on the half-pixel lattice: horizontal edges exactly on scanlines (GDAL's separate burn), vertical on centres
    for k in range(3):
        x0, y0 = X0 + res * (40.5 + 90 * k), Y1 - res * (30.5 + 200 * k)
        geoms.append([ring((x0, y0), (x0 + res * (size * 0.9), y0), (x0 + res * (size * 0.9), y0 - res * 77.0), (x0, y0 - res * 77.0))])
        geoms.append([ring((x0, y0 - res * 100), (x0, y0 - res * 160), (x0 + res * 333.0, y0 - res * 160), (x0 + res * 333.0, y0 - res * 100))])
    # a comb of 40 small holes: more rings than the kernel keeps in shared memory
    cx, cy = X0 + g.span * 1.0, Y1 - g.span * 1.5
    comb = [ring((cx - g.span * 0.45, cy - g.span * 0.2), (cx + g.span * 0.45, cy - g.span * 0.2), (cx + g.span * 0.45, cy + g.span * 0.2),
                 (cx - g.span * 0.45, cy + g.span * 0.2))]
    for k in range(40):
        hx = cx - g.span * 0.42 + g.span * 0.021 * k
        comb.append(ring((hx, cy - g.span * 0.1), (hx, cy + g.span * 0.1), (hx + g.span * 0.01, cy + g.span * 0.1), (hx + g.span * 0.01, cy - g.span * 0.1)))
    geoms.append(comb)
    geoms.append([ring((X0 - 10 * g.span, Y1 + 9 * g.span), (X0 - 9 * g.span, Y1 + 9 * g.span), (X0 - 9 * g.span, Y1 + 10 * g.span))])   # far away
    return g, RoadSet.from_geometries(geoms)


@pytest.mark.parametrize("size,channels", [(512, 3), (1024, 3), (1024, 1), (1024, 2), (2048, 3)])
def test_wide_kernel_matches_oracle_and_fused_kernel(eng, size, channels, monkeypatch):
    g, roads = _wide_case(size, seed=size + channels)
    tiles = synth.host_tiles(g, channels)
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    pairs = pairs_by_bbox(roads, tb)
    l0 = eng.launch_count
    h, z = eng.zonal_hist_host(roads, tb, pairs)
    n_wide = eng.launch_count - l0
    oh, onz = oracle_hist(roads, pairs, tiles, gt)
    assert np.array_equal(h.astype(np.uint64), oh) and np.array_equal(z.astype(np.uint64), onz)
    assert oh[:, 0].sum() > 100000 and (oh[-1] == 0).all()
    monkeypatch.setenv("RS_ZONAL_WIDE", "0")
    l0 = eng.launch_count
    hf, zf = eng.zonal_hist_host(roads, tb, pairs)
    assert eng.launch_count - l0 != n_wide                 # the two calls really took different kernels
    assert np.array_equal(h, hf) and np.array_equal(z, zf)


def test_wide_kernel_window_modes_and_slots(eng, monkeypatch):
    """full / boundless windows, a border, and a road -> row map through both kernels"""
    g, roads = _wide_case(1024, seed=3, n_poly=4)
    tiles = synth.host_tiles(g, 3)
    tb = TileBatch.from_arrays(tiles, g.transforms())
    pairs = pairs_by_bbox(roads, tb)
    slot = np.random.default_rng(0).permutation(roads.n_roads + 5)[:roads.n_roads].astype(np.int32)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("RS_ZONAL_WIDE", flag)
        res[flag] = [eng.zonal_hist_host(roads, tb, pairs, window=w, border_px=b, road_slot=s_, n_slots=None if s_ is None else roads.n_roads + 5)
                     for w, b, s_ in (("full", 0, None), ("boundless", 0, None), ("crop", 13, None), ("crop", 0, slot))]
    for (h1, z1), (h0, z0) in zip(res["1"], res["0"]):
        assert np.array_equal(h1, h0) and np.array_equal(z1, z0) and h1.sum() > 0


@pytest.mark.parametrize("f32", [False, True])
def test_u16_rescale_integer_form_equals_floating_form(eng, f32, monkeypatch):
    """PxU16x4Lut (thresholds + fixed-point guess, verified on all 65 536 inputs by the launcher) against the floating-point
    policy (RS_ZONAL_LUT=0) and the oracle, for ordinary ranges, for ranges whose steps fall on exact .5 ties, and for a
    range narrower than 255 (scale >= 1: the integer form is refused and the floating one runs)."""
    from proj_roadsurf_b200.engine import scale_params
    g = synth.Grid(3, 3)
    rr = synth.ribbon_roads(g, 12, seed=5)
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 65536, (g.n_tiles, 256, 256, 4), dtype=np.uint16)        # every source value, also the clamped tails
    tiles[..., 3] = rng.integers(0, 700, tiles.shape[:3])
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    for smin, smax in (([150.0, 300.0, 300.0, 300.0], [9000.0, 6000.0, 6000.0, 6000.0]),
                       ([0.0, 0.0, 0.0, 0.0], [510.0, 65535.0, 1020.0, 256.0]),          # k = 1/2, 1/257, 1/4: ties at x.5
                       ([0.0, 0.0, 0.0, 0.0], [65535.0, 65535.0, 65535.0, 65535.0]),       # both precisions agree on every input
                       ([1000.0, 10.0, 0.0, 100.0], [1100.0, 60000.0, 65535.0, 600.0])):   # 100-wide range: scale 2.55
        k, off = scale_params(smin, smax, f32)
        monkeypatch.setenv("RS_ZONAL_LUT", "1")                # integer thresholds
        monkeypatch.setenv("RS_ZONAL_F32EQ", "0")
        h1, z1 = eng.zonal_hist_host(rr.roads, tb, rr.pairs, rescale=(k, off, f32))
        monkeypatch.delenv("RS_ZONAL_F32EQ")                   # binary64 through the float32 evaluation where the launcher verified it
        h3, z3 = eng.zonal_hist_host(rr.roads, tb, rr.pairs, rescale=(k, off, f32))
        assert np.array_equal(h3, h1) and np.array_equal(z3, z1), (smin, smax)
        monkeypatch.setenv("RS_ZONAL_LUT", "0")                # plain floating point
        h0, z0 = eng.zonal_hist_host(rr.roads, tb, rr.pairs, rescale=(k, off, f32))
        assert np.array_equal(h1, h0) and np.array_equal(z1, z0), (smin, smax)
        monkeypatch.delenv("RS_ZONAL_LUT")                     # the default choice of the launcher
        h2, z2 = eng.zonal_hist_host(rr.roads, tb, rr.pairs, rescale=(k, off, f32))
        assert np.array_equal(h2, h0) and np.array_equal(z2, z0), (smin, smax)
        oh, onz = oracle_hist(rr.roads, rr.pairs, tiles, gt, scale_k=k, scale_off=off, rescale_f32=f32)
        assert np.array_equal(h1.astype(np.uint64), oh) and np.array_equal(z1.astype(np.uint64), onz), (smin, smax)
