"""Holds the rasterization oracles (and, with -m gpu, the CUDA kernels) to masks produced by rasterio / GDAL itself, when
tests/golden/rasterio_masks.npz exists (tests/golden/make_rasterio_golden.py writes it on a box that has rasterio 1.3.2 / GDAL
3.0.4; this repository's build container has neither, so until then these tests skip and the rasterization oracle stays
"parity unpinned")."""
import os

import numpy as np
import pytest

from oracle import cport, gdal_fill

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rasterio_masks.npz")
needs_fixture = pytest.mark.skipif(not os.path.exists(PATH), reason="no rasterio-made fixture (tests/golden/make_rasterio_golden.py)")


def _cases():
    d = np.load(PATH)
    H, W = (int(v) for v in d["shape"])
    for i in range(len(d["names"])):
        g0, g1 = d["poly_off"][i], d["poly_off"][i + 1]
        rings = [d["xy"][d["ring_off"][g]:d["ring_off"][g + 1]] for g in range(g0, g1)]
        t = tuple(d["transforms"][d["transform_index"][i]])
        yield (str(d["names"][i]), rings, t, H, W, np.unpackbits(d["rasterize"][i])[:H * W].reshape(H, W),
               np.unpackbits(d["mask_crop"][i])[:H * W].reshape(H, W), tuple(int(v) for v in d["crop_window"][i]))


@needs_fixture
def test_oracles_match_rasterio():
    n = 0
    for name, rings, t, H, W, full, crop, win in _cases():
        assert np.array_equal(cport.rasterize(rings, (H, W), t), full), name
        assert np.array_equal(cport.pair_mask_full(t, rings, W, H), crop), name
        if n % 10 == 0:                                                    # the Python restatement is slow: a tenth of the cases
            assert np.array_equal(gdal_fill.rasterize(rings, (H, W), t).astype(np.uint8), full), name
            assert (cport.geometry_window(t, rings, W, H) or (-1, -1, 0, 0)) == win, name
        n += 1
    assert n > 1000


@needs_fixture
@pytest.mark.gpu
def test_kernels_match_rasterio():
    from proj_roadsurf_b200.engine import Engine
    from proj_roadsurf_b200.geometry import PairList, RoadSet
    eng = Engine(0)
    cases = list(_cases())
    by_t = {}
    for c in cases:
        by_t.setdefault(c[2], []).append(c)
    for t, cs in by_t.items():
        roads = RoadSet.from_geometries([c[1] for c in cs])
        pairs = PairList(np.arange(len(cs) + 1, dtype=np.int32), np.zeros(len(cs), np.int32))
        H, W = cs[0][3], cs[0][4]
        full = eng.rasterize_pairs_host(roads, np.array([t]), H, W, pairs, window="full")
        crop = eng.rasterize_pairs_host(roads, np.array([t]), H, W, pairs, window="crop")
        for i, c in enumerate(cs):
            assert np.array_equal(full[i], c[5]), c[0]
            assert np.array_equal(crop[i], c[6]), c[0]
    eng.close()
