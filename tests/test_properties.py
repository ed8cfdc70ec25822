"""Property tests (hypothesis) of the rasterization oracles: the Python restatement (oracle/gdal_fill.py) and the plain-C one
(oracle/c/roadsurf_oracle.c) are two independent codings of GDAL's GDALdllImageFilledPolygon; they must agree on polygons
biased towards the cases where a scanline fill goes wrong -- vertices on the half-pixel lattice (pixel centres), vertices
exactly on scanlines, crossings within 1e-9 of x.5, horizontal edges on scanlines, slivers, repeated points.  The GPU
counterpart (kernel masks against the C oracle on > 1e5 polygons) is tests/test_gpu_properties.py."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import cport, gdal_fill

W, H = 24, 20

# coordinates: integers, halves (pixel centres), near-halves, and arbitrary reals, inside and slightly outside the raster
_coord = st.one_of(
    st.integers(-3, 27).map(float),
    st.integers(-6, 54).map(lambda v: v / 2.0),
    st.tuples(st.integers(-3, 27), st.sampled_from([0.5 - 1e-9, 0.5 + 1e-9, 0.5 - 1e-12, 0.5 + 1e-12, 1e-9, -1e-9])).map(lambda t: t[0] + t[1]),
    st.floats(-3.0, 27.0, allow_nan=False, allow_infinity=False),
)
_point = st.tuples(_coord, _coord)
_ring = st.lists(_point, min_size=3, max_size=9).map(lambda pts: np.array(pts + [pts[0]], np.float64))
_polygon = st.lists(_ring, min_size=1, max_size=3)


@settings(max_examples=400, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(_polygon)
def test_python_and_c_oracles_agree_on_adversarial_polygons(rings):
    a = gdal_fill.rasterize(rings, (H, W))
    b = cport.rasterize(rings, (H, W))
    assert np.array_equal(a.astype(np.uint8), b), (rings, np.argwhere(a != b)[:4])


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(_polygon, st.sampled_from([(0.5, 0.0, 1000.0, 0.0, -0.5, 5000.0), (0.25, 0.0, -7.125, 0.0, -0.25, 3.0625),
                                  (0.5971642834779395, 0.0, 829045.2, 0.0, -0.5971642834779395, 5933729.9)]))
def test_crop_window_masks_agree_between_oracles(rings, t):
    """rasterio.mask.mask(crop=True): window + window transform + fill, world coordinates through three tile transforms"""
    world = [np.stack([t[2] + r[:, 0] * t[0], t[5] + r[:, 1] * t[4]], 1) for r in rings]
    inside, win = gdal_fill.raster_geometry_mask(t, world, W, H)
    full = cport.pair_mask_full(t, world, W, H)
    if inside is None:
        assert full.sum() == 0
        return
    c0, r0, w, h = win
    exp = np.zeros((H, W), np.uint8)
    exp[r0:r0 + h, c0:c0 + w] = inside
    assert np.array_equal(exp, full)
    assert cport.geometry_window(t, world, W, H) == tuple(win)


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(_polygon)
def test_fill_properties(rings):
    """size-independent properties of the even-odd fill: ring order and ring direction do not matter; a polygon and its copy
    cancel (even-odd over all rings, the way GDAL collects the parts of a MultiPolygon)"""
    m = cport.rasterize(rings, (H, W))
    assert np.array_equal(m, cport.rasterize(rings[::-1], (H, W)))
    # GDAL burns a horizontal edge lying exactly on a scanline only when it runs towards -x (llrasterize.cpp), so the ring
    # direction matters for such edges and for nothing else
    on_scanline = any(r[k, 1] == r[k + 1, 1] and r[k, 1] - np.floor(r[k, 1]) == 0.5 for r in rings for k in range(len(r) - 1))
    if not on_scanline:
        assert np.array_equal(m, cport.rasterize([r[::-1].copy() for r in rings], (H, W)))
    twice = cport.rasterize(list(rings) + [r.copy() for r in rings], (H, W))
    # the doubled crossings pair up into empty spans: what is left is GDAL's separate burn of horizontal edges lying on
    # scanlines, which the single polygon has too
    assert (twice & ~m).sum() == 0
    shifted = [r + np.array([0.0, 1.0]) for r in rings]                   # one pixel down == the mask rolled by one row
    ms = cport.rasterize(shifted, (H + 1, W))
    assert np.array_equal(ms[1:], cport.rasterize(rings, (H, W)))
