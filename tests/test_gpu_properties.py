"""Property tests on the GPU (-m gpu): the kernel's masks against the plain-C oracle on > 1e5 polygons biased to the hard cases
(half-pixel lattice, vertices on scanlines, crossings within 1e-9 / 1e-12 of x.5 -- the guard of the reciprocal fast path of
the crossing evaluation), and invariants of the accumulators (sum of a histogram == pixel count of the mask, road order and
pair order do not matter)."""
import numpy as np
import pytest

from oracle import cport
from proj_roadsurf_b200.geometry import PairList, RoadSet, TileBatch

pytestmark = pytest.mark.gpu

W, H = 32, 24


def _adversarial_polygons(n, seed):
    rng = np.random.default_rng(seed)
    geoms = []
    eps = np.array([0.0, 0.0, 0.0, 1e-9, -1e-9, 1e-12, -1e-12, 3e-5, -3e-5, 9.9e-5, 1.01e-4])
    for i in range(n):
        k = int(rng.integers(3, 9))
        kind = i % 5
        if kind == 0:
            pts = rng.integers(-2, W + 3, (k, 2)).astype(np.float64)                       # integer lattice
        elif kind == 1:
            pts = rng.integers(-4, 2 * W + 5, (k, 2)) / 2.0                                # half-pixel lattice (centres)
        elif kind == 2:
            pts = rng.integers(-2, W + 3, (k, 2)) + 0.5 + rng.choice(eps, (k, 2))          # a hair off the centres
        elif kind == 3:
            pts = rng.uniform(-2, W + 2, (k, 2))
            pts[:, 1] = np.floor(pts[:, 1]) + 0.5                                          # every vertex on a scanline
        else:
            pts = rng.uniform(-2, W + 2, (k, 2))
        rings = [np.concatenate([pts, pts[:1]])]
        if i % 7 == 0:                                                                     # a second ring (hole / part)
            q = rng.integers(0, 2 * W, (4, 2)) / 2.0
            rings.append(np.concatenate([q, q[:1]]))
        geoms.append(rings)
    return geoms


@pytest.fixture(scope="module")
def eng():
    from proj_roadsurf_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_masks_match_the_c_oracle_on_1e5_adversarial_polygons(eng):
    n_total, batch, bad = 0, 20000, 0
    ident = np.array([[1.0, 0.0, 0.0, 0.0, 1.0, 0.0]])
    for b in range(6):                                                     # 120 000 polygons
        geoms = _adversarial_polygons(batch, seed=1000 + b)
        roads = RoadSet.from_geometries(geoms)
        pairs = PairList(np.arange(batch + 1, dtype=np.int32), np.zeros(batch, np.int32))
        masks = eng.rasterize_pairs_host(roads, ident, H, W, pairs, window="full")
        for i, rings in enumerate(geoms):
            exp = cport.rasterize(rings, (H, W))
            if not np.array_equal(masks[i], exp):
                bad += 1
                assert bad < 1, (b, i, rings, np.argwhere(masks[i] != exp)[:5])
        n_total += batch
    assert n_total >= 100000


def test_histogram_invariants(eng):
    from proj_roadsurf_b200 import synth
    g = synth.Grid(6, 6)
    rr = synth.ribbon_roads(g, 80, seed=4)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    hist, nz = eng.zonal_hist_host(rr.roads, tb, rr.pairs)
    masks = eng.rasterize_pairs_host(rr.roads, gt, 256, 256, rr.pairs, window="crop")
    per_road = np.zeros(80, np.int64)
    np.add.at(per_road, rr.pairs.road_of_pair(), masks.reshape(rr.pairs.n_pairs, -1).sum(1))
    for c in range(3):                                                     # sum of every band's histogram == pixels of the masks
        assert np.array_equal(hist[:, c].sum(1).astype(np.int64), per_road)
    # road order: a permuted road set gives the permuted table
    perm = np.random.default_rng(1).permutation(80)
    h2, z2 = eng.zonal_hist_host(rr.roads.subset(perm), tb, rr.pairs.take_roads(perm))
    assert np.array_equal(h2, hist[perm]) and np.array_equal(z2, nz[perm])
    # tile order: renumbered tiles (and so a different pair order inside every road) change nothing
    tperm = np.random.default_rng(2).permutation(g.n_tiles)
    inv = np.argsort(tperm)
    p3 = PairList.from_pairs(80, rr.pairs.road_of_pair(), inv[rr.pairs.pair_tile])
    h3, z3 = eng.zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles[tperm], gt[tperm]), p3)
    assert np.array_equal(h3, hist) and np.array_equal(z3, nz)
