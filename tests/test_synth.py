"""Synthetic inputs: shapes, determinism, and that the generated pair list loses no pixel."""
import numpy as np

from oracle import cport, gdal_fill
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.geometry import PairList, TileBatch, pairs_by_bbox, xyz_tile_transform


def test_grid_transforms_match_xyz_formula():
    g = synth.Grid(3, 2)
    gt = g.transforms()
    for i in range(6):
        exp = xyz_tile_transform(g.x0 + i % 3, g.y0 + i // 3, g.z)
        np.testing.assert_allclose(gt[i], exp, rtol=0, atol=1e-6)
    assert g.tile_ids()[4] == "(136679, 92198, 18)"
    assert len(set(g.keys().tolist())) == 6


def test_ribbons_are_deterministic_and_shaped_like_the_shx_distribution():
    g = synth.Grid(64, 64)
    a = synth.ribbon_roads(g, 4096)
    b = synth.ribbon_roads(g, 4096)
    assert np.array_equal(a.roads.xy, b.roads.xy) and np.array_equal(a.pairs.pair_tile, b.pairs.pair_tile)
    n = a.n_centre
    assert 9 <= np.median(n) <= 13 and n.min() >= 2 and n.max() <= 1042
    assert 30 <= np.percentile(n, 90) <= 60
    assert a.roads.n_rings > a.roads.n_roads            # holes / second parts exist
    # rings are closed
    ro = a.roads.ring_off
    assert np.array_equal(a.roads.xy[ro[:-1]], a.roads.xy[ro[1:] - 1])


def test_pair_list_is_a_superset_of_the_pixel_carrying_pairs():
    g = synth.Grid(6, 5)
    rr = synth.ribbon_roads(g, 40, seed=4)
    gt = g.transforms()
    have = set(zip(rr.pairs.road_of_pair().tolist(), rr.pairs.pair_tile.tolist()))
    for r in range(40):
        rings = rr.roads.rings(r)
        for t in range(g.n_tiles):
            if cport.pair_mask_full(gt[t], rings, 256, 256).any():
                assert (r, t) in have, (r, t)
    # and the generic bbox broad phase contains it too
    tb = TileBatch(None, gt, 256, 256, 3)
    bb = pairs_by_bbox(rr.roads, tb)
    assert have <= set(zip(bb.road_of_pair().tolist(), bb.pair_tile.tolist()))


def test_c_and_python_oracle_agree_on_world_space_ribbons():
    g = synth.Grid(3, 3)
    rr = synth.ribbon_roads(g, 10, seed=2)
    gt = g.transforms()
    road_of = rr.pairs.road_of_pair()
    n = 0
    for p in range(rr.pairs.n_pairs):
        rings = rr.roads.rings(int(road_of[p]))
        t = gt[rr.pairs.pair_tile[p]]
        inside, win = gdal_fill.raster_geometry_mask(tuple(t), rings, 256, 256)
        full = np.zeros((256, 256), np.uint8)
        if inside is not None:
            c0, r0, w, h = win
            full[r0:r0 + h, c0:c0 + w] = inside
        assert np.array_equal(full, cport.pair_mask_full(t, rings, 256, 256)), p
        n += int(full.sum())
    assert n > 5000


def test_host_tiles_kinds():
    g = synth.Grid(2, 1)
    u = synth.host_tiles(g, 3)
    assert u.shape == (2, 256, 256, 3) and u.dtype == np.uint8
    assert 0.005 < (u.max(axis=3) == 0).mean() < 0.02
    a = synth.host_tiles(g, 3, "asphalt")
    assert 100 < a[a > 0].mean() < 120
    cs = synth.host_tiles(g, 2, "class_score")
    assert set(np.unique(cs[..., 0]).tolist()) <= {0, 1, 2}
    w = synth.host_tiles(g, 4, dtype=np.uint16)
    assert w.dtype == np.uint16 and w.shape[-1] == 4
