"""Hand-derived known-answer tests for the oracle rasterizer (the reference ships none).

Each expected mask below was worked out by hand from the GDAL 3.0.x fill rule
(SURVEY.md A.1): scanline through pixel centres, an edge is active for
y_low <= cy < y_high, crossings are rounded floor(x+0.5) BEFORE sorting, pixels
xs..xe-1 between paired crossings are burnt, horizontal edges lying exactly on a
scanline are burnt separately only when they run towards -x.
"""
import numpy as np
import pytest

from oracle import cport, gdal_fill

IMPLS = [("py", gdal_fill.rasterize), ("c", cport.rasterize)]


def ring(*pts):
    pts = list(pts)
    if pts[0] != pts[-1]:
        pts.append(pts[0])
    return np.array(pts, np.float64)


def pix(mask):
    ys, xs = np.nonzero(mask)
    return sorted(zip(ys.tolist(), xs.tolist()))


@pytest.mark.parametrize("name,rast", IMPLS)
def test_integer_rectangle(name, rast):
    m = rast([ring((2, 1), (6, 1), (6, 4), (2, 4))], (8, 8))
    assert pix(m) == [(y, x) for y in (1, 2, 3) for x in (2, 3, 4, 5)]


@pytest.mark.parametrize("name,rast", IMPLS)
def test_rectangle_on_pixel_centres_orientation_matters(name, rast):
    # edges exactly through pixel centres: xa < cx <= xb keeps columns 2..4;
    # y_low <= cy < y_high keeps rows 1,2; the far horizontal edge (y = 3.5) lies on
    # scanline 3 and is burnt only if it runs towards -x.
    a = rast([ring((1.5, 1.5), (4.5, 1.5), (4.5, 3.5), (1.5, 3.5))], (6, 8))
    assert pix(a) == [(y, x) for y in (1, 2, 3) for x in (2, 3, 4)]
    b = rast([ring((1.5, 1.5), (1.5, 3.5), (4.5, 3.5), (4.5, 1.5))], (6, 8))
    assert pix(b) == [(y, x) for y in (1, 2) for x in (2, 3, 4)]


@pytest.mark.parametrize("name,rast", IMPLS)
def test_triangle_vertex_on_scanline_counted_once(name, rast):
    m = rast([ring((1, 0), (7, 0), (4, 2.5))], (4, 8))
    assert pix(m) == [(0, 2), (0, 3), (0, 4), (0, 5), (1, 3), (1, 4)]


@pytest.mark.parametrize("name,rast", IMPLS)
def test_hole(name, rast):
    m = rast([ring((0, 0), (10, 0), (10, 10), (0, 10)), ring((3, 3), (3, 7), (7, 7), (7, 3))], (10, 10))
    exp = np.ones((10, 10), np.uint8)
    exp[3:7, 3:7] = 0
    assert np.array_equal(m, exp)
    assert m.sum() == 84


@pytest.mark.parametrize("name,rast", IMPLS)
def test_overlapping_parts_cancel_even_odd(name, rast):
    m = rast([ring((0, 0), (6, 0), (6, 4), (0, 4)), ring((4, 0), (10, 0), (10, 4), (4, 4))], (4, 10))
    exp = np.ones((4, 10), np.uint8)
    exp[:, 4:6] = 0
    assert np.array_equal(m, exp)


@pytest.mark.parametrize("name,rast", IMPLS)
def test_partly_and_fully_outside(name, rast):
    m = rast([ring((-5, -5), (3, -5), (3, 2), (-5, 2))], (8, 8))
    assert pix(m) == [(y, x) for y in (0, 1) for x in (0, 1, 2)]
    assert rast([ring((20, 20), (30, 20), (30, 30), (20, 30))], (8, 8)).sum() == 0
    assert rast([ring((-20, 1), (-10, 1), (-10, 5), (-20, 5))], (8, 8)).sum() == 0
    m = rast([ring((-100, -100), (100, -100), (100, 100), (-100, 100))], (8, 8))
    assert m.sum() == 64


@pytest.mark.parametrize("name,rast", IMPLS)
def test_slivers(name, rast):
    # both crossings round to 3 -> empty span
    assert rast([ring((2.6, 0), (2.9, 0), (2.9, 5), (2.6, 5))], (5, 8)).sum() == 0
    # crossings round to 2 and 3 -> exactly column 2 (centre 2.5 in (2.4, 2.6])
    m = rast([ring((2.4, 0), (2.6, 0), (2.6, 5), (2.4, 5))], (5, 8))
    assert pix(m) == [(y, 2) for y in range(5)]


@pytest.mark.parametrize("name,rast", IMPLS)
def test_transform_north_up(name, rast):
    # 2 m pixels, origin (100, 200), north-up: world rect x 104..112, y 190..196
    # -> pixel cols 2..6, rows 2..5
    t = (2.0, 0.0, 100.0, 0.0, -2.0, 200.0)
    m = rast([ring((104, 190), (112, 190), (112, 196), (104, 196))], (8, 8), t)
    assert pix(m) == [(y, x) for y in (2, 3, 4) for x in (2, 3, 4, 5)]


def test_unclosed_ring_is_closed_by_wraparound_edge():
    closed = ring((1, 1), (6, 1), (6, 5), (1, 5))
    for _, rast in IMPLS:
        assert np.array_equal(rast([closed], (8, 8)), rast([closed[:-1]], (8, 8)))


def test_geometry_window_and_crop_equivalence():
    t = (0.5, 0.0, 1000.0, 0.0, -0.5, 5000.0)      # 16x16 px raster covers x 1000..1008, y 4992..5000
    rings = [ring((1001.2, 4995.1), (1005.7, 4995.1), (1005.7, 4998.9), (1001.2, 4998.9))]
    # pixel space: cols 2.4..11.4 -> 2..12 ; rows 2.2..9.8 -> 2..10
    assert gdal_fill.geometry_window(t, rings, 16, 16) == (2, 2, 10, 8)
    assert cport.geometry_window(t, rings, 16, 16) == (2, 2, 10, 8)
    # partly outside: clipped to the raster
    rings2 = [ring((990.0, 4990.0), (1002.0, 4990.0), (1002.0, 4997.0), (990.0, 4997.0))]
    assert gdal_fill.geometry_window(t, rings2, 16, 16) == (0, 6, 4, 10)
    assert cport.geometry_window(t, rings2, 16, 16) == (0, 6, 4, 10)
    # touching only: WindowError in rasterio
    rings3 = [ring((1008.0, 4990.0), (1010.0, 4990.0), (1010.0, 4997.0), (1008.0, 4997.0))]
    assert gdal_fill.geometry_window(t, rings3, 16, 16) is None
    assert cport.geometry_window(t, rings3, 16, 16) is None
    inside, win = gdal_fill.raster_geometry_mask(t, rings, 16, 16)
    full = gdal_fill.rasterize(rings, (16, 16), t)
    c0, r0, w, h = win
    assert np.array_equal(full[r0:r0 + h, c0:c0 + w], inside)
    assert full.sum() == inside.sum()
    assert np.array_equal(cport.pair_mask_full(t, rings, 16, 16), full)


def _random_polygon(rng, W, H, n):
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    rad = rng.uniform(0.2, 1.0, n) * min(W, H) * 0.6
    cx, cy = rng.uniform(-0.2 * W, 1.2 * W), rng.uniform(-0.2 * H, 1.2 * H)
    pts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)
    if rng.random() < 0.3:   # snap to the half-pixel lattice: exercises ties and horizontal edges
        pts = np.round(pts * 2) / 2
    return np.concatenate([pts, pts[:1]])


def test_python_and_c_oracles_agree_on_random_polygons():
    rng = np.random.default_rng(20261018)
    for i in range(300):
        W, H = int(rng.integers(4, 48)), int(rng.integers(4, 48))
        rings = [_random_polygon(rng, W, H, int(rng.integers(3, 12))) for _ in range(int(rng.integers(1, 4)))]
        a = gdal_fill.rasterize(rings, (H, W))
        b = cport.rasterize(rings, (H, W))
        assert np.array_equal(a, b), i


def test_band_ratios_known_answers():
    """statistical_analysis.py:279-293 restated (oracle/stats.band_ratios): hand-checked rows"""
    import pandas as pd
    from oracle import stats as ostats
    df = pd.DataFrame({"band1": np.array([1, 0, 5, 200, 2], np.uint8), "band2": np.array([3, 0, 0, 100, 3], np.uint8),
                       "band3": np.array([2, 7, 1, 255, 0], np.uint8), "band4": np.array([4, 0, 9, 100, 6], np.uint8)})
    out = ostats.band_ratios(df)
    assert out["R/G"].tolist() == [0.333, 0.0, 1.0, 2.0, 0.667]          # 1/3, 0/0 -> 0, 5/0 -> 1, 200/100, 2/3
    assert out["R/B"].tolist() == [0.5, 0.0, 5.0, 0.784, 1.0]            # 2/0 -> inf -> 1
    assert out["B/NIR"].tolist() == [0.5, 1.0, 0.111, 2.55, 0.0]
    v = out["VgNIR-BI"].tolist()
    assert v[0] == -0.14286 and np.isnan(v[1]) and v[2] == -1.0 and v[3] == 0.0 and v[4] == -0.33333
    assert list(out.columns) == ["band1", "band2", "band3", "band4", "R/G", "R/B", "R/NIR", "G/B", "G/NIR", "B/NIR", "VgNIR-BI"]


# ------------------------------------------------------------------------------------------
# Known answers published in rasterio's OWN test suite (rasterio 1.3.x tests/conftest.py: basic_geometry, basic_image,
# basic_image_2x2; tests/test_features.py: test_rasterize, test_geometry_mask, test_geometry_window_no_pad;
# tests/test_mask.py: test_mask_crop).  rasterio is the third-party module behind fct_misc.py:77 and
# add_tile_mask.py:112; it is not installable here, so the vectors are restated from that suite (they are the only
# published input/output pairs for this call chain the reference's dependencies carry).
# ------------------------------------------------------------------------------------------
RIO_BASIC_GEOMETRY = ring((2, 2), (2, 4.25), (4.25, 4.25), (4.25, 2), (2, 2))
RIO_SHAPE = (10, 10)


def rio_basic_image():              # all_touched=True answer; its [2:4, 2:4] part is the pixel-centre answer
    im = np.zeros(RIO_SHAPE, np.uint8)
    im[2:5, 2:5] = 1
    return im


def rio_basic_image_2x2():
    im = np.zeros(RIO_SHAPE, np.uint8)
    im[2:4, 2:4] = 1
    return im


@pytest.mark.parametrize("name,rast", IMPLS)
def test_rasterio_suite_rasterize_and_geometry_mask(name, rast):
    m = rast([RIO_BASIC_GEOMETRY], RIO_SHAPE)                       # test_rasterize: == basic_image_2x2
    assert np.array_equal(m, rio_basic_image_2x2())
    # test_geometry_mask: geometry_mask(...) == (basic_image_2x2 == 0), i.e. True outside
    assert np.array_equal(m == 0, rio_basic_image_2x2() == 0)


def test_rasterio_suite_geometry_window_and_mask_crop():
    ident = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
    # test_geometry_window_no_pad: window.flatten() == (2, 2, 3, 3)  (col_off, row_off, width, height)
    assert gdal_fill.geometry_window(ident, [RIO_BASIC_GEOMETRY], 10, 10) == (2, 2, 3, 3)
    assert cport.geometry_window(ident, [RIO_BASIC_GEOMETRY], 10, 10) == (2, 2, 3, 3)
    # test_mask_crop: masked.shape == (1, 3, 3) and masked[0] == image[2:5, 2:5] after image[4, :] = 0; image[:, 4] = 0
    from oracle import raster as oraster
    tile = {"data": rio_basic_image()[..., None], "transform": ident, "nodata": None}
    masked = oraster.mask_crop(tile, [RIO_BASIC_GEOMETRY])
    image = rio_basic_image()
    image[4, :] = 0
    image[:, 4] = 0
    assert masked.shape == (1, 3, 3)
    assert np.array_equal(masked[0], image[2:5, 2:5])


def test_polygon_within_known_answers():
    """'within' of gpd.sjoin(roads, buffered_quarries, predicate='within') (determine_class.py:57), oracle/vote.polygon_within"""
    from oracle import vote as ovote
    big = [ring((0, 0), (10, 0), (10, 10), (0, 10))]
    holed = big + [ring((4, 4), (4, 6), (6, 6), (6, 4))]
    notch = [ring((0, 0), (10, 0), (10, 10), (6, 10), (6, 4), (4, 4), (4, 10), (0, 10))]       # U shape
    W = ovote.polygon_within
    assert W([ring((1, 1), (3, 1), (3, 3), (1, 3))], big)
    assert W([ring((0, 0), (3, 0), (3, 3), (0, 3))], big)                      # touching from inside
    assert W(big, big)                                                          # equal polygons
    assert not W([ring((8, 8), (12, 8), (12, 12), (8, 12))], big)              # partly outside
    assert not W([ring((20, 20), (21, 20), (21, 21), (20, 21))], big)          # disjoint
    assert not W([ring((4.5, 4.5), (5.5, 4.5), (5.5, 5.5), (4.5, 5.5))], holed)   # inside the hole
    assert not W([ring((3, 3), (7, 3), (7, 7), (3, 7))], holed)                # covers the hole
    assert W([ring((1, 1), (3, 1), (3, 9), (1, 9))], holed)
    assert not W([ring((1, 6), (9, 6), (9, 8), (1, 8))], notch)                # all vertices inside, bridges the notch
    assert W([ring((1, 1), (9, 1), (9, 3), (1, 3))], notch)
    assert not W(big, [ring((1, 1), (3, 1), (3, 3), (1, 3))])                  # contains, not within


def test_bin_accuracy_known_answers():
    """calibration bins of final_metrics.py:541-571 (oracle/vote.bin_accuracy): hand-counted.  The lower bound of a bin is
    threshold - 0.5 (the reference's constant), so a score of 0.30 falls in every bin from 0.30 to 0.75."""
    import pandas as pd
    from oracle import vote as ovote
    df = pd.DataFrame({
        "art_score": [0.30, 0.30, 0.90, 0.10], "nat_score": [0.10, 0.20, 0.00, 0.80], "diff_score": [0.20, 0.10, 0.90, 0.70],
        "CATEGORY": ["artificial", "artificial", "artificial", "natural"],
        "cover_type": ["artificial", "natural", "artificial", "natural"], "gt_type": ["val"] * 4})
    t = ovote.bin_accuracy(df)
    assert [x.name for x in t] == ["artifical score for val", "natural score for val", "score diff in artificial roads for val",
                                   "score diff in natural roads for val"]
    art = dict(zip(np.round(t[0]["threshold"], 2), t[0]["accuracy"]))
    # artificial roads: scores 0.30 (hit), 0.30 (miss), 0.90 (hit).  Bin 0.30 holds the first two; bins 0.90-1.00 hold the
    # third only (0.30 is no longer above threshold - 0.5); bins 0.80 and 0.85 are empty.
    assert art[0.3] == 0.5 and art[0.75] == 0.5 and 0.85 not in art and 0.25 not in art
    assert art[0.9] == 1.0 and art[1.0] == 1.0 and 0.8 not in art
    nat = dict(zip(np.round(t[1]["threshold"], 2), t[1]["accuracy"]))
    assert nat[0.8] == 1.0 and 0.75 not in nat                      # the only natural road has nat_score 0.80


def test_overlay_area_known_answers():
    """area of polygon intersections (gpd.overlay(how='intersection').area, determine_class.py:110-114), oracle/overlay.py"""
    from oracle import overlay as ov

    def sq(x0, y0, x1, y1, cw=False):
        r = ring((x0, y0), (x1, y0), (x1, y1), (x0, y1))
        return [r[::-1].copy() if cw else r]
    assert ov.polygon_area(sq(0, 0, 4, 3)) == 12.0 and ov.polygon_area(sq(0, 0, 4, 3, cw=True)) == 12.0
    assert ov.intersection_area(sq(0, 0, 4, 3), sq(2, 1, 6, 5)) == 4.0
    assert ov.intersection_area(sq(0, 0, 4, 3), sq(4, 0, 6, 3)) == 0.0             # sharing an edge only
    assert ov.intersection_area(sq(0, 0, 4, 4), sq(1, 1, 2, 2)) == 1.0             # containment
    holed = sq(0, 0, 10, 10) + sq(4, 4, 6, 6)
    assert ov.polygon_area(holed) == 96.0
    assert ov.intersection_area(holed, sq(3, 3, 7, 7)) == 12.0                      # 16 minus the hole
    assert ov.intersection_area(holed, sq(4.5, 4.5, 5.5, 5.5)) == 0.0              # inside the hole
    tri = [ring((0, 0), (4, 0), (0, 4))]
    assert ov.intersection_area(tri, sq(0, 0, 2, 2)) == 4.0 and ov.intersection_area(tri, sq(1, 1, 3, 3)) == 2.0
    rows = ov.get_weighted_scores([sq(0, 0, 10, 2)], [sq(0, 0, 5, 2), sq(9.6, 0, 12, 2), sq(20, 0, 21, 1)], [0.9, 0.8, 0.7])
    assert rows == [(0, 0, 10.0, 0.5, 0.45)]                                        # the second covers 4 % <= 5 %, the third nothing
