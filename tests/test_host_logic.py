"""Host-side logic of the product package against the oracle (no GPU)."""
import numpy as np

from oracle import gdal_fill
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.geometry import (PairList, RoadSet, TileBatch, lattice_of, pairs_by_bbox, rasterio_window, rings_of)


def test_rasterio_window_matches_the_oracle_geometry_window():
    rng = np.random.default_rng(3)
    t = (0.5971642834779395, 0.0, 815000.3, 0.0, -0.5971642834779395, 5935000.7)
    n_none = 0
    for _ in range(500):
        c = np.array([815000.3 + rng.uniform(-60, 215), 5935000.7 - rng.uniform(-60, 215)])
        pts = c + rng.normal(0, rng.uniform(1, 60), (6, 2))
        if rng.random() < 0.2:
            pts = np.round(pts / 0.5971642834779395) * 0.5971642834779395      # vertices on pixel corners
        ring = np.concatenate([pts, pts[:1]])
        rs = RoadSet.from_geometries([[ring]])
        got = rasterio_window(t, rs.bbox[0], 256, 256)
        exp = gdal_fill.geometry_window(t, [ring], 256, 256)
        assert got == exp
        n_none += exp is None
    assert 0 < n_none < 450


def test_rings_of_accepts_every_geometry_form():
    sq = [[0, 0], [4, 0], [4, 4], [0, 4], [0, 0]]
    hole = [[1, 1], [1, 2], [2, 2], [2, 1], [1, 1]]
    assert len(rings_of({"type": "Polygon", "coordinates": [sq, hole]})) == 2
    assert len(rings_of({"type": "MultiPolygon", "coordinates": [[sq], [sq, hole]]})) == 3
    assert len(rings_of({"type": "Feature", "geometry": {"type": "Polygon", "coordinates": [sq]}})) == 1
    assert len(rings_of({"type": "GeometryCollection", "geometries": [{"type": "Polygon", "coordinates": [sq]}] * 2})) == 2

    class Shp:
        __geo_interface__ = {"type": "Polygon", "coordinates": [sq]}
    assert len(rings_of(Shp())) == 1
    assert rings_of(np.array(sq, float))[0].shape == (5, 2)
    rs = RoadSet.from_geometries([{"type": "Polygon", "coordinates": [[[0, 0, 9], [4, 0, 9], [4, 4, 9], [0, 0, 9]]]}])   # Z dropped
    assert rs.xy.shape == (4, 2) and list(rs.bbox[0]) == [0, 0, 4, 4]


def test_pair_list_dedup_order_and_filters():
    p = PairList.from_pairs(4, [2, 0, 2, 2, 0, 3], [5, 1, 5, 3, 0, 9])          # duplicates dropped like drop_duplicates
    assert p.road_pair_off.tolist() == [0, 2, 2, 4, 5]
    assert p.pair_tile.tolist() == [0, 1, 3, 5, 9]
    q = p.restrict_tiles(1, 6)
    assert q.road_pair_off.tolist() == [0, 1, 1, 3, 3] and q.pair_tile.tolist() == [0, 2, 4]
    r = p.take_roads([3, 0])
    assert r.road_pair_off.tolist() == [0, 1, 3] and r.pair_tile.tolist() == [9, 0, 1]


def test_lattice_detection_and_host_broad_phase():
    g = synth.Grid(9, 6)
    rr = synth.ribbon_roads(g, 60, seed=9)
    tb = TileBatch(None, g.transforms(), 256, 256, 3)
    lat = lattice_of(tb)
    assert lat is not None and (lat.nx, lat.ny) == (9, 6) and (lat.lut >= 0).all()
    # irregular tile set: no lattice, brute-force broad phase still works and agrees with the lattice one
    gt = g.transforms().copy()
    shuffled = TileBatch(None, gt[::-1].copy(), 256, 256, 3)
    a = pairs_by_bbox(rr.roads, tb)
    b = pairs_by_bbox(rr.roads, shuffled)
    assert a.n_pairs == b.n_pairs
    bad = gt.copy()
    bad[3, 0] *= 1.5                                       # one tile with another resolution
    assert lattice_of(TileBatch(None, bad, 256, 256, 3)) is None
    c = pairs_by_bbox(rr.roads, TileBatch(None, bad, 256, 256, 3))
    assert c.n_pairs >= a.n_pairs - 20


def test_clip_border_px():
    from proj_roadsurf_b200.road_segmentation.determine_class import clip_border_px
    assert [clip_border_px(w) for w in (64, 256, 512, 1024)] == [0, 1, 3, 5]
    assert clip_border_px(256, 1.0) == 0


def test_geotiff_tiles_open_without_rasterio(tmp_path):
    from PIL import Image, TiffImagePlugin
    from proj_roadsurf_b200.functions import fct_misc
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (16, 24, 3), dtype=np.uint8)
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550] = (0.5, 0.5, 0.0); ifd.tagtype[33550] = 12
    ifd[33922] = (0.0, 0.0, 0.0, 1000.0, 5000.0, 0.0); ifd.tagtype[33922] = 12
    ifd[42113] = "0"; ifd.tagtype[42113] = 2
    path = str(tmp_path / "18_136678_92197.tif")
    Image.fromarray(a).save(path, tiffinfo=ifd)
    t = fct_misc.open_tile(path)
    assert np.array_equal(t["data"], a) and t["transform"] == (0.5, 0.0, 1000.0, 0.0, -0.5, 5000.0) and t["nodata"] == 0
    assert fct_misc.open_tile(str(tmp_path / "missing.tif")) is None
    plain = str(tmp_path / "plain.tif")
    Image.fromarray(a[..., 0]).save(plain)
    t2 = fct_misc.open_tile(plain)
    assert t2["data"].shape == (16, 24, 1) and t2["nodata"] is None and t2["transform"][0] == 1.0


def test_bbox_pairs_matches_brute_force():
    from proj_roadsurf_b200.geometry import bbox_pairs
    rng = np.random.default_rng(9)
    a = rng.uniform(0, 100, (700, 2)); a = np.concatenate([a, a + rng.uniform(0, 8, (700, 2))], 1)
    b = rng.uniform(0, 100, (300, 2)); b = np.concatenate([b, b + rng.uniform(0, 12, (300, 2))], 1)
    b[0] = [a[5, 2], a[5, 1], a[5, 2] + 1, a[5, 3]]                 # touching boxes overlap (closed comparison)
    ia, ib = bbox_pairs(a, b, chunk=128)
    exp = [(i, j) for i in range(len(a)) for j in range(len(b))
           if a[i, 0] <= b[j, 2] and a[i, 2] >= b[j, 0] and a[i, 1] <= b[j, 3] and a[i, 3] >= b[j, 1]]
    assert list(zip(ia.tolist(), ib.tolist())) == exp and (5, 0) in exp
    ia, ib = bbox_pairs(np.zeros((0, 4)), b)
    assert len(ia) == 0 and len(ib) == 0


def test_clip_labels_areas_match_the_overlay_oracle():
    """determine_class.clip_labels_host (the numpy restatement of determine_class.py:62-95 the GPU form is held to): the clipped label has the area of
    label AND scaled tile (oracle/overlay.py), rows follow the (label, tile) join"""
    import pandas as pd
    from oracle import overlay as ov
    from proj_roadsurf_b200.road_segmentation import determine_class as dc
    rng = np.random.default_rng(13)

    def star(c, rmin, rmax, n):
        ang = (np.arange(n) + rng.uniform(0.0, 0.8, n)) * (2 * np.pi / n)
        rad = rng.uniform(rmin, rmax, n)
        pts = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        return np.concatenate([pts, pts[:1]])
    labels = []
    for i in range(12):
        c = rng.uniform(5, 35, 2)
        rings = [star(c, 4, 9, int(rng.integers(8, 20)))]
        if i % 2 == 0:
            rings.append(star(c, 0.5, 2, 6))
        labels.append({"type": "Polygon", "coordinates": [r.tolist() for r in rings]})
    tiles, tid = [], []
    for ty in range(4):
        for tx in range(4):
            x0, y0 = 10.0 * tx, 10.0 * ty
            tiles.append({"type": "Polygon", "coordinates": [[[x0, y0], [x0 + 10, y0], [x0 + 10, y0 + 10], [x0, y0 + 10], [x0, y0]]]})
            tid.append(f"({tx}, {ty}, 18)")
    lab_df = pd.DataFrame({"OBJECTID": np.arange(12) + 1, "BELAGSART": 100, "geometry": labels})
    til_df = pd.DataFrame({"id": tid, "title": "t", "geometry": tiles})
    out = dc.clip_labels_host(lab_df, til_df, fact=0.99)
    assert list(out.columns) == ["OBJECTID", "BELAGSART", "tile_id", "title", "geometry"] and len(out) > 20
    n_checked = 0
    for row in out.itertuples():
        lab = [np.array(r) for r in labels[row.OBJECTID - 1]["coordinates"]]
        tx, ty = [int(v) for v in row.tile_id.strip("()").split(",")[:2]]
        cx, cy, h = 10.0 * tx + 5, 10.0 * ty + 5, 5 * 0.99
        rect = [np.array([[cx - h, cy - h], [cx + h, cy - h], [cx + h, cy + h], [cx - h, cy + h]])]
        exp = ov.intersection_area(lab, rect)
        got = ov.polygon_area([np.array(r) for r in row.geometry["coordinates"]]) if row.geometry["coordinates"] else 0.0
        assert abs(got - exp) <= 1e-9 * max(1.0, exp), (row.OBJECTID, row.tile_id, got, exp)
        n_checked += exp > 0
    assert n_checked > 15
    # every (label, tile) pair whose closed shapes intersect is a row, in label-major order
    assert out["OBJECTID"].tolist() == sorted(out["OBJECTID"].tolist())


def test_table_from_rows_equals_the_per_pair_frames():
    """the column-wise pixel table of get_pixel_values_batch == the concatenation of the per-pair frames (fct_misc.py:87-121)"""
    import pandas as pd
    from proj_roadsurf_b200.functions import fct_misc
    rng = np.random.default_rng(3)
    for no_data in (None, 0, 7):
        for trial in range(6):
            P = int(rng.integers(1, 9))
            n = rng.integers(0, 40, P)
            n[rng.integers(0, P)] = 0                                  # a pair without pixels
            off = np.concatenate([[0], np.cumsum(n)])
            vals = rng.integers(0, 4, (int(off[-1]), 3)).astype(np.uint8) * np.uint8(7 if no_data == 7 else 1)
            vals[rng.random(len(vals)) < 0.2] = 0
            ids = rng.integers(100, 105, P)
            bands = (1, 2, 3) if trial % 2 == 0 else (1, 2)        # (the reference indexes with band - 1: lists start at 1)
            frames = []
            for p in range(P):
                rows = vals[off[p]:off[p + 1]]
                if len(rows) == 0 and no_data is not None:
                    continue
                frames.append(fct_misc._frames_from_rows(rows, no_data, bands, str(p), {"road_id": ids[p]}))
            exp = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
            got = fct_misc.table_from_rows(vals, off, no_data, bands, ids)
            if len(exp) == 0:
                assert len(got) == 0
                continue
            assert list(got.columns) == list(exp.columns)
            for c in exp.columns:
                assert np.array_equal(got[c].to_numpy(), exp[c].to_numpy()), (no_data, trial, c)
                assert got[c].dtype == exp[c].dtype, (no_data, trial, c, got[c].dtype, exp[c].dtype)
