"""The reference-shaped helpers (proj_roadsurf_b200.functions / .road_segmentation) on the GPU against the golden
fixtures produced by the reference's own code (tests/golden/) and against the CPU oracle.  -m gpu."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import cport, gdal_fill, raster as oraster, stats as ostats, vote as ovote
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.geometry import PairList, RoadSet, TileBatch
from test_golden import assert_frame_matches, frame, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from proj_roadsurf_b200.functions import fct_misc, fct_rasters, fct_statistics
    from proj_roadsurf_b200.road_segmentation import determine_class, final_metrics
    return fct_misc, fct_statistics, fct_rasters, determine_class, final_metrics


def test_get_pixel_values_golden(mods):
    fct_misc = mods[0]
    g = load("pixel_values")
    data = np.array(g["data"], np.uint8)
    acc = {None: pd.DataFrame(), 0: pd.DataFrame()}
    for case in g["cases"]:
        nodata = case["nodata"]
        fct_misc.register_tile("18_1_1.tif", data, g["transform"], nodata)
        if case["geom"] == "__accumulated__":
            assert_frame_matches(case["result"], acc[nodata])
            continue
        geom = g["geoms"][case["geom"]]
        one = fct_misc.get_pixel_values(geom, "18_1_1.tif", range(1, 4), pd.DataFrame(), road_id=case["geom"])
        assert_frame_matches(case["result"], one)
        acc[nodata] = fct_misc.get_pixel_values(geom, "18_1_1.tif", range(1, 4), acc[nodata], road_id=case["geom"])
    # missing tile: error logged, empty frame (fct_misc.py:83-85)
    assert len(fct_misc.get_pixel_values(g["geoms"]["rect"], "nope.tif", range(1, 4), pd.DataFrame(), road_id=1)) == 0
    # shapes that miss the raster: rasterio.mask.mask raises ValueError
    far = {"type": "Polygon", "coordinates": [[[0, 0], [1, 0], [1, 1], [0, 0]]]}
    with pytest.raises(ValueError):
        fct_misc.get_pixel_values(far, "18_1_1.tif", range(1, 4), pd.DataFrame())
    fct_misc.clear_tiles()


def test_get_pixel_values_batch_equals_the_reference_double_loop(mods):
    fct_misc = mods[0]
    g = synth.Grid(3, 3)
    rr = synth.ribbon_roads(g, 8, seed=6)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    for nodata in (None, 0):
        tb = TileBatch.from_arrays(tiles, gt, nodata)
        got = fct_misc.get_pixel_values_batch(rr.roads, tb, rr.pairs, range(1, 4), road_ids=np.arange(100, 108))
        exp = pd.DataFrame()
        road_of = rr.pairs.road_of_pair()
        for p in range(rr.pairs.n_pairs):                      # statistical_analysis.py:180-193 with the oracle
            t = int(rr.pairs.pair_tile[p])
            tile = {"data": tiles[t], "transform": tuple(gt[t]), "nodata": nodata}
            try:
                exp = oraster.get_pixel_values(rr.roads.rings(int(road_of[p])), tile, range(1, 4), exp, road_id=100 + int(road_of[p]))
            except ValueError:
                pass
        assert len(exp) > 1000
        assert list(got.columns) == list(exp.columns)
        for c in exp.columns:
            assert np.array_equal(got[c].to_numpy().astype(np.int64), exp[c].to_numpy().astype(np.int64)), (nodata, c)


def test_stats_helpers_golden(mods):
    fs = mods[1]
    g = load("stats_groupby")
    px = frame(g["pixels"], {"band1": np.uint8, "band2": np.uint8})
    for case in g["cases"]:
        assert_frame_matches(case["result"], fs.get_df_stats_groupby(px, case["col"], ["road_id"], case["suffix"]))
    g = load("stats_no_group")
    px = frame(g["pixels"], {"band1": np.uint8, "band2": np.uint8})
    d = None
    for grp in g["groups"]:
        d = fs.get_df_stats_no_group(px.loc[grp["rows"]], "band1", d, "_1")
    from test_golden import same
    for k, v in g["result"].items():
        assert all(same(x, y) for x, y in zip(v, d[k])) and len(v) == len(d[k]), k
    assert_frame_matches(g["as_df"], fs.get_df_stats_no_group(px, "band2", None, "", True))


def test_vote_and_metrics_golden(mods):
    determine_class, final_metrics = mods[3], mods[4]
    g = load("vote")
    roads = frame(g["roads"])
    preds = frame(g["predictions"])
    thresholds = [s["threshold"] for s in g["sweep"]]
    sweep = determine_class.determine_detected_class_sweep(preds, roads, thresholds)
    for step, comp in zip(g["sweep"], sweep):
        comp["tag"] = comp.apply(lambda row: final_metrics.get_tag(row), axis=1)
        gold = frame(step["comparison"])
        assert comp["road_id"].tolist() == gold["road_id"].tolist()
        assert comp["cover_type"].tolist() == gold["cover_type"].tolist()          # bit-exact classes
        assert comp["tag"].tolist() == gold["tag"].tolist()
        for col in ("nat_score", "art_score", "diff_score"):
            np.testing.assert_allclose(comp[col].astype(float), gold[col].astype(float), rtol=1e-6, atol=1e-12)
        by_class, glob = final_metrics.get_metrics(comp, ["artificial", "natural"])
        assert_frame_matches(step["by_class"], by_class)
        assert_frame_matches(step["global"], glob)
    one = determine_class.determine_detected_class(preds, roads, thresholds[3])
    assert one["cover_type"].tolist() == frame(g["sweep"][3]["comparison"])["cover_type"].tolist()
    # from_preds_to_metrics (final_metrics.py:126-159): the same step through the reference's accumulating call
    comp, bc, gl = final_metrics.from_preds_to_metrics(preds, roads, pd.DataFrame(), pd.DataFrame(), "val", thresholds[3], show=True)
    assert comp["tag"].tolist() == frame(g["sweep"][3]["comparison"])["tag"].tolist()
    assert bc["dataset"].tolist() == ["val", "val"] and gl["threshold"].tolist() == [thresholds[3]]
    assert_frame_matches(g["sweep"][3]["by_class"], bc.drop(columns=["dataset", "threshold"]))
    assert_frame_matches(g["sweep"][3]["global"], gl.drop(columns=["dataset", "threshold"]))


def test_zonal_stats_like_rasterstats(mods):
    fr = mods[2]
    rng = np.random.default_rng(4)
    raster = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    raster[rng.random((40, 50)) < 0.1] = 0
    affine = (2.0, 0.0, 600000.0, 0.0, -2.0, 200000.0)

    def poly(cx, cy, r, n):
        a = np.linspace(0, 2 * np.pi, n, endpoint=False)
        pts = np.stack([cx + r * np.cos(a), cy + r * np.sin(a)], 1)
        return np.concatenate([pts, pts[:1]])
    vectors = [[poly(600040, 199960, 25, 9)], [poly(600095, 199930, 30, 7)],          # the second one hangs off the raster
               [poly(599990, 200010, 18, 5)], [poly(600300, 199960, 5, 4)],            # the last one misses it entirely
               [poly(600050, 199950, 30, 12), poly(600050, 199950, 10, 6)]]           # with a hole
    stats = ["min", "max", "mean", "median", "std", "count", "sum", "percentile_10", "percentile_90"]
    for nodata in (None, 0):
        got = fr.zonal_stats(vectors, raster, affine=affine, stats=stats, nodata=nodata)
        exp = ostats.zonal_stats(vectors, raster, affine, stats=[s for s in stats if not s.startswith("percentile")], nodata=nodata,
                                 percentiles=[10, 90])
        assert len(got) == len(exp) == 5
        for a, b in zip(got, exp):
            assert a["count"] == b["count"]
            for k in stats:
                if b[k] is None:
                    assert a[k] is None
                elif k in ("min", "max", "median", "count", "sum"):
                    assert a[k] == b[k], k
                else:
                    assert abs(a[k] - b[k]) <= 1e-6 * abs(b[k]), k
        assert got[3]["count"] == 0 and got[0]["count"] > 100


def test_rasterize_like_rasterio(mods):
    fr = mods[2]
    rng = np.random.default_rng(8)
    shapes = []
    for i in range(5):
        a = np.sort(rng.uniform(0, 2 * np.pi, 7))
        pts = np.stack([30 + 20 * rng.random() + 18 * np.cos(a), 25 + 10 * rng.random() + 15 * np.sin(a)], 1)
        shapes.append({"type": "Polygon", "coordinates": [np.concatenate([pts, pts[:1]]).tolist()]})
    got = fr.rasterize(shapes, (64, 80))
    exp = np.zeros((64, 80), np.uint8)
    for sh in shapes:
        exp |= gdal_fill.rasterize(gdal_fill.rings_from_geojson(sh), (64, 80))
    assert np.array_equal(got, exp) and exp.sum() > 500
    valued = fr.rasterize([(shapes[0], 7), (shapes[1], 9)], (64, 80), fill=1)
    m0 = gdal_fill.rasterize(gdal_fill.rings_from_geojson(shapes[0]), (64, 80)).astype(bool)
    m1 = gdal_fill.rasterize(gdal_fill.rings_from_geojson(shapes[1]), (64, 80)).astype(bool)
    exp2 = np.ones((64, 80), np.uint8); exp2[m0] = 7; exp2[m1] = 9
    assert np.array_equal(valued, exp2)


def test_road_stats_table_and_filter(mods):
    fs = mods[1]
    from proj_roadsurf_b200.engine import default_engine
    g = synth.Grid(6, 6)
    rr = synth.ribbon_roads(g, 30, seed=15)
    tiles = synth.host_tiles(g, 3, "asphalt")
    gt = g.transforms()
    tb = TileBatch.from_arrays(tiles, gt)
    stats, hist, nz = default_engine().zonal_stats_host(rr.roads, tb, rr.pairs, nodata_mode="none", ddof=1, want_hist=True)
    oh, onz = cport.zonal_accumulate(rr.roads.xy, rr.roads.ring_off, rr.roads.road_ring_off, rr.pairs.road_pair_off,
                                     rr.pairs.pair_tile, tiles, gt)
    assert np.array_equal(hist.astype(np.uint64), oh)
    ids = np.arange(30) + 1000
    got = fs.road_stats_from_accumulators(stats, ids, (1, 2, 3))
    exp = ostats.road_stats_table(oh, onz, ids, "N", (1, 2, 3))
    assert list(got["road_id"]) == list(exp["road_id"]) and len(got) > 10
    for c in exp.columns:
        a, b = got[c].to_numpy().astype(float), exp[c].to_numpy().astype(float)
        ok = (a == b) | (np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= 0.0100001)     # 2-decimal rounding ties
        assert ok.all(), c
        if c.startswith(("min", "max", "median")) or c == "count":
            assert np.array_equal(a, b), c
    f1 = fs.filter_roads(got, (1, 2, 3))
    f2 = ostats.filter_roads(exp, (1, 2, 3))
    assert list(f1["road_id"]) == list(f2["road_id"])


def test_threshold_sweep_raster(mods):
    determine_class, final_metrics = mods[3], mods[4]
    g = synth.Grid(8, 8)
    rr = synth.ribbon_roads(g, 80, seed=23)
    tiles = synth.host_tiles(g, 2, "class_score")
    tb = TileBatch.from_arrays(tiles, g.transforms())
    jh = determine_class.accumulate_class_planes(rr.roads, tb, rr.pairs)
    by_class, glob, bi, bt, cover = final_metrics.threshold_sweep(jh, rr.gt_class)
    rows, obest = ovote.sweep(jh, rr.gt_class)
    assert bi == obest and len(glob) == 20
    for i, r in enumerate(rows):
        assert abs(glob["f1b"][i] - r["f1b"]) <= 1e-6 and abs(glob["Pw"][i] - r["Pw"]) <= 1e-6
        assert by_class["TP"][2 * i] == r["TP_0"] and by_class["FN"][2 * i + 1] == r["FN_1"]


def test_workflows_match_the_reference_shaped_pipeline(mods):
    """workflows.road_band_statistics == the reference pipeline (per-pair get_pixel_values -> groupby stats -> filter) run
    with the oracle; workflows.road_surface_vote == the oracle sweep."""
    from proj_roadsurf_b200 import workflows
    g = synth.Grid(4, 4)
    rr = synth.ribbon_roads(g, 14, seed=33)
    tiles = synth.host_tiles(g, 3, "asphalt")
    # zero values that differ per band AND per tile: with nodata == 0 get_pixel_values pads every (road, tile) call on its
    # own (fct_misc.py:95-111), so a multi-tile road is not the same as one call over its merged pixels
    rng = np.random.default_rng(5)
    for t in range(g.n_tiles):
        tiles[t, :, :, t % 3][rng.random((256, 256)) < 0.05 * (1 + t % 4)] = 0
    tiles[rng.random(tiles.shape[:3]) < 0.02] = 0
    gt = g.transforms()
    for nodata in (None, 0):
        tb = TileBatch.from_arrays(tiles, gt, nodata)
        table, filtered = workflows.road_band_statistics(rr.roads, tb, BANDS=(1, 2, 3))        # GPU broad phase inside
        # reference shape with the oracle: statistical_analysis.py:180-246
        bb = pairs_by_bbox_local(rr.roads, tb)
        pix = pd.DataFrame()
        road_of = bb.road_of_pair()
        for p in range(bb.n_pairs):
            t = int(bb.pair_tile[p])
            try:
                pix = oraster.get_pixel_values(rr.roads.rings(int(road_of[p])), {"data": tiles[t], "transform": tuple(gt[t]), "nodata": nodata},
                                               range(1, 4), pix, road_id=int(road_of[p]))
            except ValueError:
                pass
        exp = None
        for b in (1, 2, 3):
            sub = ostats.get_df_stats_groupby(pix, f"band{b}", ["road_id"], f"_{b}")
            sub["road_id"] = sub.index
            sub = sub.reset_index(drop=True)
            exp = sub if exp is None else pd.merge(exp, sub, on="road_id")
        exp["count"] = exp["count_1"]
        assert table["road_id"].tolist() == exp["road_id"].tolist()
        for c in ("count",) + tuple(f"{k}_{b}" for b in (1, 2, 3) for k in ("min", "max", "median")):
            assert np.array_equal(table[c].to_numpy().astype(np.int64) if c != f"median_{c[-1]}" else table[c].to_numpy(),
                                  exp[c].to_numpy().astype(np.int64) if c != f"median_{c[-1]}" else exp[c].to_numpy()), (nodata, c)
        for b in (1, 2, 3):
            for k in ("mean", "std", "margin"):
                a, e = table[f"{k}_{b}"].to_numpy(float), exp[f"{k}_{b}"].to_numpy(float)
                ok = (np.abs(a - e) <= 0.0100001) | (np.isnan(a) & np.isnan(e))
                assert ok.all(), (nodata, k, b)
        assert len(filtered) <= len(table)
    cs = synth.host_tiles(g, 2, "class_score")
    res = workflows.road_surface_vote(rr.roads, TileBatch.from_arrays(cs, gt), rr.gt_class)
    rows, best = ovote.sweep(res["joint_hist"], rr.gt_class)
    assert res["best_index"] == best
    assert abs(res["global_metrics"]["f1b"][best] - rows[best]["f1b"]) <= 1e-6
    assert set(res["comparison"]["tag"]) <= {"TP", "FN", "wrong class"}


def pairs_by_bbox_local(roads, tb):
    from proj_roadsurf_b200.geometry import pairs_by_bbox
    return pairs_by_bbox(roads, tb)


def test_border_px_is_clip_labels_in_raster_form(mods):
    determine_class = mods[3]
    from proj_roadsurf_b200.engine import default_engine
    assert determine_class.clip_border_px(256) == 1 and determine_class.clip_border_px(1024) == 5
    g = synth.Grid(3, 3)
    rr = synth.ribbon_roads(g, 10, seed=44)
    tiles = synth.host_tiles(g, 3)
    gt = g.transforms()
    for border in (1, 7):
        h, z = default_engine().zonal_hist_host(rr.roads, TileBatch.from_arrays(tiles, gt), rr.pairs, border_px=border)
        inner = tiles.copy()
        exp = np.zeros_like(h, dtype=np.uint64)
        road_of = rr.pairs.road_of_pair()
        for p in range(rr.pairs.n_pairs):
            t = int(rr.pairs.pair_tile[p])
            m = cport.pair_mask_full(gt[t], rr.roads.rings(int(road_of[p])), 256, 256).astype(bool)
            m[:border] = False; m[-border:] = False; m[:, :border] = False; m[:, -border:] = False
            px = inner[t][m]
            for c in range(3):
                exp[road_of[p], c] += np.bincount(px[:, c], minlength=256).astype(np.uint64)
        assert np.array_equal(h.astype(np.uint64), exp), border
    assert exp.sum() > 10000


def test_diff_score_sweep_matches_the_reference_loop(mods):
    determine_class, final_metrics = mods[3], mods[4]
    g = load("vote")
    roads = frame(g["roads"])
    preds = frame(g["predictions"])
    comp = determine_class.determine_detected_class(preds, roads, 0.1)
    comp["tag"] = comp.apply(lambda row: final_metrics.get_tag(row), axis=1)
    thresholds = np.arange(0, 1., 0.05)
    by_class, glob, best_thr, best_results = final_metrics.diff_score_sweep(comp, thresholds)
    # final_metrics.py:429-478 with the oracle
    best, max_f1 = 0, None
    for i, thr in enumerate(thresholds):
        f = comp.copy().drop(columns=["tag"])
        f.loc[f["diff_score"] < thr, "cover_type"] = "undetermined"
        f["tag"] = [ovote.get_tag(c, k) for c, k in zip(f["cover_type"], f["CATEGORY"])]
        obc, og = ovote.get_metrics(f)
        for col in ("Pw", "Rw", "f1w", "Pb", "Rb", "f1b"):
            assert abs(glob[col][i] - og[col][0]) <= 1e-6
        assert by_class["TP"][2 * i] == obc["TP"][0] and by_class["FN"][2 * i + 1] == obc["FN"][1]
        if i == 0 or og["f1b"][0] > max_f1:
            best, max_f1 = i, og["f1b"][0]
    assert best_thr == (0 if best == 0 else round(float(thresholds[best]), 2))
    assert set(best_results["tag"]) <= {"TP", "FN", "wrong class"}


def test_instance_faithful_weighted_scores(mods):
    """get_weighted_scores with overlay areas counted in pixels per (road, detection) -> determine_detected_class -> metrics"""
    determine_class, final_metrics = mods[3], mods[4]
    g = synth.Grid(4, 4)
    rr = synth.ribbon_roads(g, 16, seed=52)
    gt = g.transforms()
    rng = np.random.default_rng(6)
    # blobby detections: 32 px blocks, ~60 % of them carry a detection id
    blocks = rng.integers(1, 200, (g.n_tiles, 8, 8)).astype(np.uint16)
    blocks[rng.random(blocks.shape) < 0.4] = 0
    inst = np.repeat(np.repeat(blocks, 32, 1), 32, 2)
    score = np.round(rng.uniform(0.05, 1.0, 200), 3)
    cls = np.where(rng.random(200) < 0.6, "artificial", "natural").astype(object)
    tb = TileBatch.from_arrays(inst[..., None], gt)
    ids = np.arange(16) + 500
    got = determine_class.get_weighted_scores_raster(rr.roads, tb, rr.pairs, score, cls, road_ids=ids)
    road_of = rr.pairs.road_of_pair()
    masks = [cport.pair_mask_full(gt[rr.pairs.pair_tile[p]], rr.roads.rings(int(road_of[p])), 256, 256).astype(bool)
             for p in range(rr.pairs.n_pairs)]
    exp = ovote.weighted_scores_raster(masks, inst, road_of, rr.pairs.pair_tile, ids, score, cls)
    assert len(exp) > 10
    a = got.sort_values(["OBJECTID", "instance"]).reset_index(drop=True)
    b = exp.sort_values(["OBJECTID", "instance"]).reset_index(drop=True)
    assert a["OBJECTID"].tolist() == b["OBJECTID"].tolist() and a["instance"].tolist() == b["instance"].tolist()
    assert a["det_class_name"].tolist() == b["det_class_name"].tolist()
    assert np.array_equal(a["area_pred_in_label"].to_numpy(), b["area_pred_in_label"].to_numpy())      # 2-dp rounded, exact
    np.testing.assert_allclose(a["weighted_score"], b["weighted_score"], rtol=1e-12)
    roads_df = pd.DataFrame({"OBJECTID": ids, "CATEGORY": np.where(rr.gt_class == 0, "artificial", "natural"), "gt_type": "gt"})
    comp = determine_class.determine_detected_class(got, roads_df, 0.2)
    ocomp = ovote.determine_detected_class(exp, roads_df, 0.2)
    assert comp["cover_type"].tolist() == ocomp["cover_type"].tolist()
    # with the clip_labels border the label area shrinks
    clipped = determine_class.get_weighted_scores_raster(rr.roads, tb, rr.pairs, score, cls, road_ids=ids, clip_fact=0.99)
    assert len(clipped) > 0 and set(clipped["OBJECTID"]) <= set(ids.tolist())


def test_get_pixel_values_from_a_geotiff_file(mods, tmp_path):
    """the reference's call shape: get_pixel_values(geometry, '<z>_<x>_<y>.tif', BANDS, df, road_id=...)"""
    from PIL import Image, TiffImagePlugin
    fct_misc = mods[0]
    g = load("pixel_values")
    data = np.array(g["data"], np.uint8)
    for nodata in (None, 0):
        ifd = TiffImagePlugin.ImageFileDirectory_v2()
        t = g["transform"]
        ifd[33550] = (t[0], -t[4], 0.0); ifd.tagtype[33550] = 12
        ifd[33922] = (0.0, 0.0, 0.0, t[2], t[5], 0.0); ifd.tagtype[33922] = 12
        if nodata is not None:
            ifd[42113] = str(nodata); ifd.tagtype[42113] = 2
        path = str(tmp_path / f"18_1_{0 if nodata is None else 1}.tif")
        Image.fromarray(data).save(path, tiffinfo=ifd)
        for case in g["cases"]:
            if case["nodata"] != nodata or case["geom"] == "__accumulated__":
                continue
            one = fct_misc.get_pixel_values(g["geoms"][case["geom"]], path, range(1, 4), pd.DataFrame(), road_id=case["geom"])
            assert_frame_matches(case["result"], one)


def test_ks_from_histograms_matches_scipy(mods):
    from scipy import stats as sstats
    fs = mods[1]
    rng = np.random.default_rng(17)
    R = 12
    hist = np.zeros((R, 256), np.uint32)
    samples = []
    for r in range(R):
        n = int(rng.integers(30, 4000))
        v = np.clip(np.rint(rng.normal(100 + 6 * (r % 3), 5 + r, n)), 0, 255).astype(np.int64)
        samples.append(v)
        hist[r] = np.bincount(v, minlength=256)
    road_type = np.array([100, 200] * 6)
    got = fs.ks_test_from_hists(hist, road_type)
    for r in range(R):
        pooled = np.concatenate([samples[i] for i in range(R) if road_type[i] == road_type[r]])
        ref = sstats.ks_2samp(samples[r], pooled, method="asymp")
        assert abs(got["ks_D"][r] - round(float(ref.statistic), 3)) < 1e-12
        assert abs(got["ks_p"][r] - float("{:0.3e}".format(ref.pvalue))) <= 1e-6 * max(ref.pvalue, 1e-300) + 1e-300


def test_cover_stats_from_accumulators(mods):
    fct_misc, fs = mods[0], mods[1]
    from proj_roadsurf_b200.engine import default_engine
    g = synth.Grid(3, 3)
    rr = synth.ribbon_roads(g, 9, seed=5)
    tiles = synth.host_tiles(g, 3, "asphalt")
    tb = TileBatch.from_arrays(tiles, g.transforms())
    hist, nz = default_engine().zonal_hist_host(rr.roads, tb, rr.pairs)
    road_type = np.array([100, 200, 100, 100, 200, 100, 200, 100, 100])
    got = fs.cover_stats_from_accumulators(hist, nz, road_type, (1, 2, 3))
    # reference shape: pixel table -> get_df_stats_no_group per (cover, band), statistical_analysis.py:296-316
    pix = fct_misc.get_pixel_values_batch(rr.roads, tb, rr.pairs, range(1, 4))
    pix["road_type"] = road_type[pix["road_id"].to_numpy()]
    cover_stats = {'cover': [], 'band': [], 'min': [], 'max': [], 'mean': [], 'median': [], 'std': [], 'margin': [], 'count': []}
    for cover in sorted(pix["road_type"].unique().tolist()):
        for b in (1, 2, 3):
            sub = pix[pix["road_type"] == cover]
            cover_stats['cover'].append(cover)
            cover_stats['band'].append(b)
            cover_stats = ostats.get_df_stats_no_group(sub, f"band{b}", cover_stats)
    exp = pd.DataFrame(cover_stats)
    exp['mean'] = exp['mean'].round(1); exp['std'] = exp['std'].round(1); exp['margin'] = exp['margin'].round(1)
    assert got["cover"].tolist() == exp["cover"].tolist() and got["band"].tolist() == exp["band"].tolist()
    for c in ("min", "max", "median", "count"):
        assert got[c].tolist() == exp[c].tolist(), c
    for c in ("mean", "std", "margin"):
        assert np.all(np.abs(got[c].to_numpy(float) - exp[c].to_numpy(float)) <= 0.1000001), c


def test_band_ratios_exhaustive(mods):
    """add_band_ratios (statistical_analysis.py:279-293): every (a, b) uint8 combination on every band pair, bit-exact
    against the reference's pandas statements (oracle/stats.band_ratios)"""
    fs = mods[1]
    hi, lo = np.repeat(np.arange(256, dtype=np.uint8), 256), np.tile(np.arange(256, dtype=np.uint8), 256)
    df = pd.DataFrame({"band1": np.concatenate([hi, hi]), "band2": np.concatenate([lo, hi]),
                       "band3": np.concatenate([hi, lo]), "band4": np.concatenate([lo, lo]),
                       "road_id": np.arange(131072) % 7})
    exp = ostats.band_ratios(df)
    got = fs.add_band_ratios(df.copy())
    assert list(got.columns) == list(exp.columns)
    for c in exp.columns:
        assert got[c].dtype == exp[c].dtype, c
        assert np.array_equal(got[c].to_numpy(), exp[c].to_numpy(), equal_nan=True), c
    assert np.isnan(got["VgNIR-BI"].to_numpy()).sum() == 257        # 0/0 rows keep their NaN, as in the reference
    with pytest.raises(KeyError):
        fs.add_band_ratios(df[["band1", "band2", "band3"]].copy(), range(1, 4))


def test_bin_accuracy_matches_the_script_body(mods):
    """final_metrics.bin_accuracy (calibration tables, final_metrics.py:541-571) against the oracle restatement"""
    fm = mods[4]
    rng = np.random.default_rng(12)
    n = 5000
    art = np.round(rng.random(n), 3)
    nat = np.round(rng.random(n), 3)
    art[:200] = np.arange(0, 1.05, 0.05)[rng.integers(0, 21, 200)]            # scores exactly on bin edges
    nat[100:300] = (np.arange(0, 1.05, 0.05) - 0.5)[rng.integers(0, 21, 200)]
    df = pd.DataFrame({"art_score": art, "nat_score": nat, "diff_score": np.abs(art - nat),
                       "CATEGORY": rng.choice(["artificial", "natural"], n, p=[0.8, 0.2]),
                       "cover_type": rng.choice(["artificial", "natural", "undetermined", "undetected"], n, p=[0.6, 0.25, 0.05, 0.1]),
                       "gt_type": rng.choice(["val", "trn", "tst", "oth"], n)})
    exp = ovote.bin_accuracy(df)
    got = fm.bin_accuracy(df)
    assert len(got) == len(exp) == 16
    for a, b in zip(got, exp):
        assert a.name == b.name
        assert a["threshold"].tolist() == b["threshold"].tolist()
        assert a["accuracy"].tolist() == b["accuracy"].tolist()
    one = fm.bin_accuracy(df[df["gt_type"] == "tst"].iloc[:3])
    ref = ovote.bin_accuracy(df[df["gt_type"] == "tst"].iloc[:3])
    assert [t["accuracy"].tolist() for t in one] == [t["accuracy"].tolist() for t in ref]


def test_rasterio_suite_known_answers_on_the_gpu(mods):
    """The vectors rasterio's own tests publish for this call chain (tests/test_oracle_kat.py, RIO_*): rasterize,
    geometry_mask and mask(crop=True) through the CUDA-backed drop-ins."""
    from test_oracle_kat import RIO_BASIC_GEOMETRY, RIO_SHAPE, rio_basic_image, rio_basic_image_2x2
    fct_misc, _, fct_rasters = mods[0], mods[1], mods[2]
    geom = {"type": "Polygon", "coordinates": [RIO_BASIC_GEOMETRY.tolist()]}
    assert np.array_equal(fct_rasters.rasterize([geom], out_shape=RIO_SHAPE), rio_basic_image_2x2())
    outside = fct_rasters.rasterize([geom], out_shape=RIO_SHAPE, fill=1, default_value=0).astype(bool)    # geometry_mask
    assert np.array_equal(outside, rio_basic_image_2x2() == 0)
    # mask(crop=True) on basic_image: the 3 x 3 window keeps the 2 x 2 block of ones -> get_pixel_values returns 4 rows of 1
    fct_misc.register_tile("rio_basic.tif", rio_basic_image()[..., None], (1.0, 0.0, 0.0, 0.0, 1.0, 0.0), None)
    df = fct_misc.get_pixel_values(geom, "rio_basic.tif", range(1, 2), pd.DataFrame(), road_id=7)
    assert df["band1"].tolist() == [1, 1, 1, 1] and df["road_id"].tolist() == [7] * 4
    fct_misc.clear_tiles()


def test_roads_in_quarries_within_join(mods):
    """determine_class.get_roads_in_quarries (determine_class.py:41-62): the 'within' join on the GPU against the oracle's
    exact-predicate restatement, on random roads over buffered-quarry-like polygons (with holes and concave outlines)"""
    from test_oracle_kat import ring
    dc = mods[3]
    rng = np.random.default_rng(5)
    quarries = []
    for q in range(6):
        c = rng.uniform(20, 80, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 24))
        rad = rng.uniform(8, 22, 24)
        ext = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        rings = [np.concatenate([ext, ext[:1]])]
        if q % 2 == 0:
            rings.append(ring((c[0] - 2, c[1] - 2), (c[0] - 2, c[1] + 2), (c[0] + 2, c[1] + 2), (c[0] + 2, c[1] - 2)))
        quarries.append({"type": "Polygon", "coordinates": [r.tolist() for r in rings]})
    roads, ids = [], []
    for i in range(400):
        c = rng.uniform(5, 95, 2)
        w, h = rng.uniform(0.5, 6, 2)
        th = rng.uniform(0, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        pts = (np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) @ R.T) + c
        roads.append({"type": "Polygon", "coordinates": [np.concatenate([pts, pts[:1]]).tolist()]})
        ids.append(1000 + i)
    roads_df = pd.DataFrame({"OBJECTID": ids, "BELAGSART": 100, "geometry": roads})
    quarries_df = pd.DataFrame({"id": np.arange(6) + 1, "geometry": quarries})
    inq, notq = dc.get_roads_in_quarries(quarries_df, roads_df)
    exp = [(ids[i], j + 1) for i, r in enumerate(roads) for j, q in enumerate(quarries)
           if ovote.polygon_within([np.array(x) for x in r["coordinates"]], [np.array(x) for x in q["coordinates"]])]
    assert len(exp) > 10
    assert list(zip(inq["OBJECTID"].tolist(), inq["id"].tolist())) == exp
    assert inq["index_right"].tolist() == [j - 1 for _, j in exp]
    assert notq["OBJECTID"].tolist() == [i for i in ids if i not in {e[0] for e in exp}]
    assert list(notq.index) == list(range(len(notq)))
    # the hand-derived cases of tests/test_oracle_kat.py::test_polygon_within_known_answers through the kernel
    big = {"type": "Polygon", "coordinates": [ring((0, 0), (10, 0), (10, 10), (0, 10)).tolist()]}
    holed = {"type": "Polygon", "coordinates": [ring((0, 0), (10, 0), (10, 10), (0, 10)).tolist(), ring((4, 4), (4, 6), (6, 6), (6, 4)).tolist()]}
    notch = {"type": "Polygon", "coordinates": [ring((0, 0), (10, 0), (10, 10), (6, 10), (6, 4), (4, 4), (4, 10), (0, 10)).tolist()]}
    cases = [ring((1, 1), (3, 1), (3, 3), (1, 3)), ring((0, 0), (3, 0), (3, 3), (0, 3)), ring((8, 8), (12, 8), (12, 12), (8, 12)),
             ring((4.5, 4.5), (5.5, 4.5), (5.5, 5.5), (4.5, 5.5)), ring((3, 3), (7, 3), (7, 7), (3, 7)), ring((1, 6), (9, 6), (9, 8), (1, 8)),
             ring((1, 1), (9, 1), (9, 3), (1, 3)), ring((0, 0), (10, 0), (10, 10), (0, 10))]
    rd = pd.DataFrame({"OBJECTID": np.arange(len(cases)), "geometry": [{"type": "Polygon", "coordinates": [c.tolist()]} for c in cases]})
    inq, _ = dc.get_roads_in_quarries([big, holed, notch], rd)
    got = sorted(zip(inq["OBJECTID"].tolist(), inq["index_right"].tolist()))
    assert got == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2), (3, 0), (4, 0), (5, 0), (5, 1), (6, 0), (6, 1), (6, 2), (7, 0)]


def test_statistical_analysis_example_end_to_end(tmp_path):
    """examples/statistical_analysis_b200.py: GeoTIFF tiles on disk -> ingest -> pairs -> statistics -> pixel table ->
    ratios -> per-type statistics -> KS, checked for consistency between the accumulator path and the pixel-table path"""
    pytest.importorskip("PIL.Image")
    import importlib.util
    spec = importlib.util.spec_from_file_location("sa_example", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                             "statistical_analysis_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(["--out", str(tmp_path), "--tiles-x", "4", "--tiles-y", "3", "--roads", "14"])
    grid = synth.Grid(4, 3)
    assert np.array_equal(out["tiles"].pixels, synth.host_tiles(grid, 4, "asphalt"))          # the ingest read back what was written
    assert np.allclose(out["tiles"].gt, grid.transforms(), rtol=0, atol=1e-6)
    px, st = out["pixels_per_band"], out["roads_stats"]
    assert len(st) > 5 and len(px) == int(st["count"].sum())
    for b in range(1, 5):                       # per-road statistics from the accumulators == pandas on the pixel table
        grp = ostats.get_df_stats_groupby(px, f"band{b}", ["road_id"], f"_{b}")
        m = st.merge(grp, on="road_id", suffixes=("", "_tbl"))
        assert len(m) == len(st)
        for c in ("min", "max", "median", "mean", "std", "margin"):
            a, e = m[f"{c}_{b}"].to_numpy(float), m[f"{c}_{b}_tbl"].to_numpy(float)
            assert np.allclose(a, e, rtol=0, atol=0.0051, equal_nan=True), (c, b)
    assert {"R/G", "B/NIR", "VgNIR-BI"} <= set(px.columns)
    assert sorted(out["cover_stats"]["cover"].unique().tolist()) == sorted(np.unique(out["road_type"]).tolist())
    assert os.path.exists(os.path.join(str(tmp_path), "tables", "ks_test.csv"))
    assert {"ks_p_band1", "ks_D_band4"} <= set(out["ks"].columns)


def test_final_metrics_example_end_to_end(tmp_path):
    """examples/final_metrics_b200.py: quarries rule -> raster vote sweep -> tags / metrics -> diff-score sweep -> calibration
    bins; the vote of the example is re-derived with the oracle from the same accumulators"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fm_example", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                             "final_metrics_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(["--out", str(tmp_path), "--tiles-x", "5", "--tiles-y", "4", "--roads", "40"])
    keep, comp = out["keep"], out["comparison"]
    assert len(out["in_quarries"]) > 0 and len(keep) + out["in_quarries"]["OBJECTID"].nunique() == 40
    ids = np.arange(40) if out["roads"].ids is None else np.asarray(out["roads"].ids)
    assert set(comp["road_id"]) <= set(ids[keep].tolist())
    assert set(comp["tag"]) <= {"TP", "FN", "wrong class"}
    assert len(out["accuracy_tables"]) == 4 and all(list(t.columns) == ["threshold", "accuracy"] for t in out["accuracy_tables"])
    # oracle re-derivation of the best-threshold vote from the joint histogram the workflow returns
    jh = out["vote"]["joint_hist"]
    thr = np.arange(0, 1., 0.05)
    bi = out["vote"]["best_index"]
    from proj_roadsurf_b200.road_segmentation.determine_class import score_cutoffs
    cover_o = ovote.raster_vote(jh, int(score_cutoffs(thr)[bi]), rule="score", min_area_frac=0.05)[0]
    known = np.isin(out["gt_class"][keep], (0, 1))
    names = np.array(["artificial", "natural", "undetermined", "undetected"], dtype=object)
    assert comp["cover_type"].tolist() == names[np.asarray(cover_o)[known]].tolist()
    for f in ("by_class_metrics.csv", "global metrics.csv", "comparison_best_threshold.csv"):
        assert os.path.exists(os.path.join(str(tmp_path), "tables", f))


def test_detections_to_planes(mods):
    """determine_class.detections_to_planes: detection polygons (score, det_class) burnt into the class / score / instance
    planes of the raster vote, against the oracle fill composed in ascending score order"""
    dc = mods[3]
    g = synth.Grid(3, 2)
    gt = g.transforms()
    rng = np.random.default_rng(21)
    X0, Y1 = g.origin
    span = g.span
    dets, rows = [], []
    for i in range(25):
        c = np.array([X0 + rng.uniform(0, 3 * span), Y1 - rng.uniform(0, 2 * span)])
        ang = np.sort(rng.uniform(0, 2 * np.pi, 7))
        rad = rng.uniform(5, 45, 7)
        pts = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        dets.append({"type": "Polygon", "coordinates": [np.concatenate([pts, pts[:1]]).tolist()]})
        rows.append((float(np.round(rng.uniform(0.05, 1.0), 3)), int(rng.integers(0, 2))))
    df = pd.DataFrame({"geometry": dets, "score": [r[0] for r in rows], "det_class": [r[1] for r in rows]})
    tiles = TileBatch(None, gt, 256, 256, 1)
    cs, inst = dc.detections_to_planes(df, tiles, batch_pairs=7)
    exp_cs = np.zeros((6, 256, 256, 2), np.uint8)
    exp_in = np.zeros((6, 256, 256, 1), np.uint16)
    for d in np.argsort(df["score"].to_numpy(), kind="stable"):
        ring = [np.array(dets[d]["coordinates"][0])]
        for t in range(6):
            m = oraster.pair_inside_mask(gt[t], ring, 256, 256) != 0
            exp_cs[t, :, :, 0][m] = 1 + rows[d][1]
            exp_cs[t, :, :, 1][m] = int(np.rint(rows[d][0] * 255.0))
            exp_in[t, :, :, 0][m] = d + 1
    assert exp_in.any() and (exp_cs[..., 0] == 2).any()
    assert np.array_equal(cs.pixels, exp_cs) and np.array_equal(inst.pixels, exp_in)
    with pytest.raises(SystemExit):
        dc.detections_to_planes(pd.DataFrame({"geometry": dets[:1], "score": [0.5], "det_class": [3]}), tiles)


def test_overlay_areas_and_vector_get_weighted_scores(mods):
    """rs_overlay_area_host / determine_class.get_weighted_scores (determine_class.py:97-120, vector form) against the
    strip-decomposition oracle (oracle/overlay.py): hand cases with shared edges, holes, containment and both ring
    orientations, then random star polygons with holes"""
    from oracle import overlay as ov
    from proj_roadsurf_b200.engine import default_engine
    from test_oracle_kat import ring
    dc = mods[3]
    eng = default_engine()

    def sq(x0, y0, x1, y1, cw=False):
        r = ring((x0, y0), (x1, y0), (x1, y1), (x0, y1))
        return [r[::-1].copy() if cw else r]
    A = [sq(0, 0, 4, 3), sq(0, 0, 4, 3, cw=True), sq(0, 0, 10, 10) + sq(4, 4, 6, 6), sq(0, 0, 10, 10, cw=True) + sq(4, 4, 6, 6, cw=True),
         [ring((0, 0), (4, 0), (0, 4))]]
    B = [sq(2, 1, 6, 5), sq(4, 0, 6, 3), sq(1, 1, 2, 2, cw=True), sq(3, 3, 7, 7), sq(4.5, 4.5, 5.5, 5.5), sq(0, 0, 2, 2), sq(1, 1, 3, 3),
         sq(0, 0, 4, 3), sq(-5, -5, 20, 20) + sq(1, 1, 3, 2)]
    ia, ib = np.repeat(np.arange(len(A)), len(B)), np.tile(np.arange(len(B)), len(A))
    got, area_a = eng.overlay_area_host(RoadSet.from_geometries(A), RoadSet.from_geometries(B), ia, ib)
    exp = np.array([ov.intersection_area(A[i], B[j]) for i, j in zip(ia, ib)])
    assert np.array_equal(got, exp), np.stack([ia, ib, got, exp], 1)[got != exp]     # small integers and halves: exact
    assert area_a.tolist() == [12.0, 12.0, 96.0, 96.0, 8.0]

    rng = np.random.default_rng(31)

    def star(c, rmin, rmax, n):
        ang = (np.arange(n) + rng.uniform(0.0, 0.8, n)) * (2 * np.pi / n)     # bounded gaps: a valid ring around c
        rad = rng.uniform(rmin, rmax, n)
        pts = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        return pts[::-1].copy() if rng.random() < 0.5 else pts
    labels, preds = [], []
    for i in range(40):
        c = rng.uniform(0, 100, 2)
        rings_ = [star(c, 6, 14, int(rng.integers(8, 24)))]
        if i % 3 == 0:
            rings_.append(star(c, 1, 3, 6))                        # a hole around the centre
        labels.append(rings_)
    for j in range(60):
        preds.append([star(rng.uniform(0, 100, 2), 3, 12, int(rng.integers(4, 30)))])
    a, b = RoadSet.from_geometries(labels), RoadSet.from_geometries(preds)
    from proj_roadsurf_b200.geometry import bbox_pairs
    ia, ib = bbox_pairs(a.bbox, b.bbox)
    assert len(ia) > 100
    got, area_a = eng.overlay_area_host(a, b, ia, ib)
    exp = np.array([ov.intersection_area(labels[i], preds[j]) for i, j in zip(ia, ib)])
    assert np.allclose(got, exp, rtol=1e-9, atol=1e-9) and (exp > 1.0).sum() > 30
    assert np.allclose(area_a, [ov.polygon_area(l) for l in labels], rtol=1e-12)
    # the table form
    score = np.round(rng.uniform(0.05, 1.0, len(preds)), 3)
    gt_df = pd.DataFrame({"OBJECTID": np.arange(len(labels)) + 500, "BELAGSART": 100, "geometry": labels})
    pr_df = pd.DataFrame({"score": score, "det_class_name": "artificial", "geometry": preds})
    out = dc.get_weighted_scores(gt_df, pr_df)
    rows = ov.get_weighted_scores(labels, preds, score)
    assert len(rows) > 20 and len(out) == len(rows)
    assert out["OBJECTID"].tolist() == [500 + r[0] for r in rows]
    assert np.allclose(out["score"].to_numpy(), [score[r[1]] for r in rows])
    assert np.allclose(out["joined_area"].to_numpy(), [r[2] for r in rows], rtol=1e-9)
    assert out["area_pred_in_label"].tolist() == [r[3] for r in rows]
    assert np.allclose(out["weighted_score"].to_numpy(), [r[4] for r in rows], rtol=1e-12)
    assert "area_label" in gt_df.columns


def test_vector_flow_clip_overlay_vote_metrics(mods):
    """The vector chain of final_metrics.py:254-316 on synthetic polygons: clip_labels -> get_weighted_scores (overlay areas on
    the GPU) -> determine_detected_class -> tags -> metrics, against the same chain through the oracles"""
    from oracle import overlay as ov
    dc, fm = mods[3], mods[4]
    rng = np.random.default_rng(41)

    def star(c, rmin, rmax, n):
        ang = (np.arange(n) + rng.uniform(0.0, 0.8, n)) * (2 * np.pi / n)
        rad = rng.uniform(rmin, rmax, n)
        pts = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        return np.concatenate([pts, pts[:1]])
    labels = [{"type": "Polygon", "coordinates": [star(rng.uniform(3, 37, 2), 2, 5, int(rng.integers(8, 16))).tolist()]} for _ in range(30)]
    cat = rng.choice([100, 200], 30, p=[0.7, 0.3])
    tiles = [{"type": "Polygon", "coordinates": [[[10.0 * tx, 10.0 * ty], [10.0 * tx + 10, 10.0 * ty], [10.0 * tx + 10, 10.0 * ty + 10],
                                                   [10.0 * tx, 10.0 * ty + 10], [10.0 * tx, 10.0 * ty]]]} for ty in range(4) for tx in range(4)]
    preds = [{"type": "Polygon", "coordinates": [star(rng.uniform(3, 37, 2), 1.5, 5, int(rng.integers(6, 14))).tolist()]} for _ in range(80)]
    lab_df = pd.DataFrame({"OBJECTID": np.arange(30) + 1, "BELAGSART": cat, "geometry": labels})
    lab_df["CATEGORY"] = lab_df.apply(dc.determine_category, axis=1)
    lab_df["gt_type"] = "val"
    til_df = pd.DataFrame({"id": [f"({i % 4}, {i // 4}, 18)" for i in range(16)], "geometry": tiles})
    pr_df = pd.DataFrame({"score": np.round(rng.uniform(0.05, 1.0, 80), 3), "det_class_name": rng.choice(["artificial", "natural"], 80),
                          "dataset": "val", "geometry": preds})
    visible = dc.clip_labels(lab_df, til_df)
    visible = visible[[len(g["coordinates"]) > 0 for g in visible["geometry"]]].reset_index(drop=True)
    got = dc.get_weighted_scores(visible, pr_df)
    lab_rings = [[np.array(r) for r in g["coordinates"]] for g in visible["geometry"]]
    pr_rings = [[np.array(r) for r in g["coordinates"]] for g in preds]
    rows = ov.get_weighted_scores(lab_rings, pr_rings, pr_df["score"].to_numpy())
    assert len(rows) > 30 and len(got) == len(rows)
    assert got["OBJECTID"].tolist() == [int(visible["OBJECTID"][r[0]]) for r in rows]
    assert got["area_pred_in_label"].tolist() == [r[3] for r in rows]
    assert np.allclose(got["weighted_score"].to_numpy(), [r[4] for r in rows], rtol=1e-12)
    roads = lab_df[["OBJECTID", "CATEGORY", "gt_type", "geometry"]]
    got = got.drop(columns=["BELAGSART", "CATEGORY", "gt_type", "tile_id", "area_label", "joined_area"], errors="ignore")   # :263-265
    for thr in (0.0, 0.3, 0.6):
        comp, bc, gl = fm.from_preds_to_metrics(got, roads, pd.DataFrame(), pd.DataFrame(), "val", thr)
        exp = ovote.determine_detected_class(got, roads, thr)
        assert comp["road_id"].tolist() == exp["road_id"].tolist() and comp["cover_type"].tolist() == exp["cover_type"].tolist()
        exp["tag"] = [ovote.get_tag(c, k) for c, k in zip(exp["cover_type"], exp["CATEGORY"])]
        assert comp["tag"].tolist() == exp["tag"].tolist()
        obc, ogl = ovote.get_metrics(exp)
        assert np.allclose(gl[["Pb", "Rb", "f1b"]].to_numpy(float), ogl[["Pb", "Rb", "f1b"]].to_numpy(float), rtol=1e-9, atol=1e-12)


def test_multi_tile_nodata_padding_golden():
    """workflows.road_band_statistics on a road over several tiles against the statistics the REFERENCE's get_pixel_values +
    get_df_stats_groupby produce (tests/golden/multi_tile.json): with tile nodata 0 every (road, tile) call is zero-padded on
    its own (fct_misc.py:95-111) before the calls are concatenated (statistical_analysis.py:187-193)."""
    from proj_roadsurf_b200 import workflows
    g = load("multi_tile")
    tiles = np.ascontiguousarray(np.array(g["tiles"], np.uint8))
    tr = np.array(g["transforms"], np.float64)
    names = list(g["geoms"])
    roads = RoadSet.from_geometries([g["geoms"][n] for n in names], ids=np.arange(len(names)))
    for case in g["cases"]:
        tb = TileBatch.from_arrays(tiles, tr, case["nodata"])
        table, _ = workflows.road_band_statistics(roads, tb, BANDS=(1, 2, 3))
        assert table["road_id"].tolist() == list(range(len(names)))
        for b in (1, 2, 3):
            gold = case["stats"][str(b)]
            for name, row in zip(gold["index"], gold["data"]):
                exp = dict(zip(gold["columns"], row))
                got = table.iloc[names.index(name)]
                for k in ("min", "max", "median", "mean", "std", "margin"):
                    assert abs(float(got[f"{k}_{b}"]) - float(exp[f"{k}_{b}"])) <= 1e-9, (case["nodata"], name, k, b)
                if b == 1:
                    assert int(got["count"]) == int(exp["count_1"])


def test_vote_many_detections_near_ties_golden(mods):
    """20-60 detections per (road, class), both classes holding the same multiset in different orders: the reference's
    groupby(...).sum() is Kahan-compensated (pandas group_sum), which decides the exact tie test of determine_class.py:164;
    vote_table_kernel sums the same way."""
    determine_class = mods[3]
    g = load("vote_many")
    assert g["naive_differs"] > 0
    comp = determine_class.determine_detected_class(frame(g["predictions"]), frame(g["roads"]), 0.0)
    gold = frame(g["comparison"])
    assert comp["road_id"].tolist() == gold["road_id"].tolist()
    assert comp["cover_type"].tolist() == gold["cover_type"].tolist()
    for col in ("nat_score", "art_score", "diff_score"):
        np.testing.assert_allclose(comp[col].astype(float), gold[col].astype(float), rtol=1e-12, atol=0)


def test_overlay_area_hole_touching_its_shell():
    """A valid OGC hole may touch its shell at single points, and the parts of a MultiPolygon may touch each other: the ring
    orientation pass (ring_sign_kernel) must not take such a vertex as the ring's representative point."""
    from oracle import overlay as ov
    from proj_roadsurf_b200.engine import default_engine
    from test_oracle_kat import ring
    eng = default_engine()
    shell = ring((0, 0), (10, 0), (10, 10), (0, 10))
    hole_first_on_shell = ring((0, 5), (4, 3), (4, 7))                     # the hole's FIRST vertex lies on the shell's edge x = 0
    hole_corner = ring((10, 10), (6, 9), (9, 6))                           # ... and one whose first vertex is the shell's corner
    part_touching = [ring((20, 0), (24, 0), (24, 4), (20, 4)), ring((24, 4), (28, 4), (28, 8), (24, 8))]   # two parts sharing a vertex
    A = [[shell, hole_first_on_shell], [shell[::-1].copy(), hole_first_on_shell], [shell, hole_corner, hole_first_on_shell],
         part_touching]
    B = [[ring((-5, -5), (40, -5), (40, 40), (-5, 40))], [ring((2, 2), (8, 2), (8, 8), (2, 8))], [ring((23, 3), (25, 3), (25, 5), (23, 5))]]
    ia, ib = np.repeat(np.arange(len(A)), len(B)), np.tile(np.arange(len(B)), len(A))
    got, area_a = eng.overlay_area_host(RoadSet.from_geometries(A), RoadSet.from_geometries(B), ia, ib)
    exp = np.array([ov.intersection_area(A[i], B[j]) for i, j in zip(ia, ib)])
    assert np.array_equal(got, exp), np.stack([ia, ib, got, exp], 1)[got != exp]
    assert area_a.tolist() == [ov.polygon_area(a) for a in A] == [92.0, 92.0, 84.5, 32.0]


def test_zonal_stats_float_dem_like_rasterstats(mods):
    """the reference's live zonal_stats call (fct_rasters.py:147-163): polygons over a float32 DEM with nodata=-9999, on a raster
    wider than 2048 px, with NaN pixels, features partly and wholly off the raster.  count / min / max / median exact, mean / std
    within the north star's 1e-6 (the oracle reduces in binary64 like the kernel; rasterstats itself in float32)."""
    fr = mods[2]
    rng = np.random.default_rng(12)
    Hh, Ww = 300, 3000
    yy, xx = np.mgrid[0:Hh, 0:Ww]
    dem = (450.0 + 0.05 * xx + 0.3 * yy + 3.0 * np.sin(xx / 37.0) + rng.normal(0, 0.2, (Hh, Ww))).astype(np.float32)
    dem[rng.random((Hh, Ww)) < 0.03] = -9999.0
    dem[rng.random((Hh, Ww)) < 0.01] = np.nan
    dem[100:140, 2500:2560] = -9999.0                                   # a feature with no valid pixel
    affine = (2.0, 0.0, 2600000.0, 0.0, -2.0, 1200000.0)               # swissALTI3D 2 m grid
    from test_oracle_kat import ring

    def quad(c0, r0, c1, r1, jitter=0.3):
        pts = [(c0, r0), (c1, r0 + 3), (c1 + 2, r1), (c0 - 1, r1 - 2)]
        return [ring(*[(affine[2] + (c + rng.uniform(-jitter, jitter)) * 2.0, affine[5] - (r + rng.uniform(-jitter, jitter)) * 2.0) for c, r in pts])]
    vectors = [quad(10, 10, 80, 40), quad(2100, 50, 2950, 120), quad(-30, -20, 25, 30), quad(2980, 280, 3100, 330),
               quad(2505, 105, 2555, 135), quad(5000, 5000, 5100, 5100), quad(1000.5, 150.5, 1001.4, 151.4, 0.0),
               quad(300, 5, 1900, 295)]
    vectors[0].append(ring(*[(affine[2] + c * 2.0, affine[5] - r * 2.0) for c, r in [(30, 20), (30, 30), (50, 30), (50, 20)]]))   # hole
    stats = ["count", "min", "max", "mean", "median", "std", "sum", "percentile_10", "percentile_90"]
    got = fr.zonal_stats(vectors, dem, affine=affine, stats=stats, nodata=-9999)
    exp = ostats.zonal_stats(vectors, dem, affine, stats=[s for s in stats if not s.startswith("percentile")], nodata=-9999,
                             percentiles=(10.0, 90.0))
    assert [g["count"] for g in got] == [e["count"] for e in exp]
    assert got[4]["count"] == 0 and got[4]["mean"] is None and got[5]["count"] == 0 and got[7]["count"] > 100000
    for g, e in zip(got, exp):
        for k in stats:
            if e[k] is None:
                assert g[k] is None
            elif k in ("count", "min", "max"):
                assert g[k] == e[k], k
            elif k == "median":
                assert abs(g[k] - e[k]) <= 1e-6 * abs(e[k]), k          # float32 mean of the two middle values vs binary64
            else:
                assert abs(g[k] - e[k]) <= 1e-6 * max(abs(e[k]), 1e-30), (k, g[k], e[k])
    # other dtypes: int16 goes through the same float path; wide integers and float64 are refused, never rounded silently
    dem16 = np.nan_to_num(dem, nan=-9999.0).astype(np.int16)
    g16 = fr.zonal_stats(vectors[:2], dem16, affine=affine, stats=["count", "min", "max", "median"], nodata=-9999)
    e16 = ostats.zonal_stats(vectors[:2], dem16.astype(np.float64), affine, stats=["count", "min", "max", "median"], nodata=-9999)
    assert g16 == e16
    with pytest.raises(TypeError):
        fr.zonal_stats(vectors[:1], dem.astype(np.float64), affine=affine, nodata=-9999)


def test_exact_intersects_reject_and_grid_broad_phase():
    """f1: the exact 'intersects' reject of the bounding-box pairs (gpd.sjoin's predicate, statistical_analysis.py:170-171) against
    an exact rational oracle, and the GPU broad phase for tile sets that are not on a lattice against the brute-force host one."""
    from oracle import overlay as ov
    from proj_roadsurf_b200.engine import default_engine
    from proj_roadsurf_b200.geometry import pairs_by_bbox
    from test_oracle_kat import ring
    eng = default_engine()
    g = synth.Grid(7, 6)
    rr = synth.ribbon_roads(g, 70, seed=41)
    tb = TileBatch.from_arrays(np.zeros((g.n_tiles, 4, 4, 1), np.uint8), g.transforms() * np.array([64.0, 1, 1, 1, 64.0, 1]))   # 4 px tiles, same extents
    ext = tb.extents()
    # hand cases on tile 0: a polygon that contains the tile, one whose hole contains the tile, one touching it at a corner,
    # one touching along an edge, one near-miss, one inside the tile
    x0, y0, x1, y1 = ext[0]
    d = x1 - x0
    hand = [[ring((x0 - d, y0 - d), (x1 + d, y0 - d), (x1 + d, y1 + d), (x0 - d, y1 + d))],
            [ring((x0 - 2 * d, y0 - 2 * d), (x1 + 2 * d, y0 - 2 * d), (x1 + 2 * d, y1 + 2 * d), (x0 - 2 * d, y1 + 2 * d)),
             ring((x0 - d, y0 - d), (x0 - d, y1 + d), (x1 + d, y1 + d), (x1 + d, y0 - d))],
            [ring((x0 - d, y0 - d), (x0, y0), (x0 - d, y0))],
            [ring((x0 - d, y0), (x0 - d, y1), (x0, y1), (x0, y0))],
            [ring((x0 - d, y0 - d), (x0 - 1e-6 * d, y0 - 1e-6 * d), (x0 - d, y0 - 1e-6 * d))],
            [ring((x0 + 0.3 * d, y0 + 0.3 * d), (x0 + 0.6 * d, y0 + 0.3 * d), (x0 + 0.5 * d, y0 + 0.7 * d))]]
    geoms = [rr.roads.rings(r) for r in range(rr.roads.n_roads)] + hand
    roads = RoadSet.from_geometries(geoms)
    cand = pairs_by_bbox(roads, tb)
    kept = eng.pairs_intersect_host(roads, tb, cand)
    road_of = cand.road_of_pair()
    exp = np.array([ov.polygon_intersects_rect(geoms[r], ext[t]) for r, t in zip(road_of, cand.pair_tile)])
    got = np.zeros(cand.n_pairs, bool)
    kept_set = set(zip(kept.road_of_pair().tolist(), kept.pair_tile.tolist()))
    for i, (r, t) in enumerate(zip(road_of.tolist(), cand.pair_tile.tolist())):
        got[i] = (r, t) in kept_set
    assert np.array_equal(got, exp), np.nonzero(got != exp)[0][:10]
    assert 0 < exp.sum() < len(exp)                                       # the reject removes pairs, and keeps some
    n_r = rr.roads.n_roads
    hand_rows = {(r - n_r, t) for r, t in kept_set if r >= n_r and t == 0}
    assert hand_rows == {(0, 0), (2, 0), (3, 0), (5, 0)}                   # contains / corner touch / edge touch / inside: yes
    # rejected pairs hold no pixel: statistics through either list agree
    tiles = synth.host_tiles(g, 3)
    tb2 = TileBatch.from_arrays(tiles, g.transforms())
    c2 = pairs_by_bbox(rr.roads, tb2)
    k2 = eng.pairs_intersect_host(rr.roads, tb2, c2)
    assert k2.n_pairs < c2.n_pairs
    h_all, z_all = eng.zonal_hist_host(rr.roads, tb2, c2)
    h_kept, z_kept = eng.zonal_hist_host(rr.roads, tb2, k2)
    assert np.array_equal(h_all, h_kept) and np.array_equal(z_all, z_kept)

    # ---- tiles that are not on a lattice: random rectangles of very different sizes, shuffled, some degenerate ----
    rng = np.random.default_rng(8)
    T = 600
    cx, cy = rng.uniform(0, 1000, T), rng.uniform(0, 700, T)
    w, h = rng.uniform(2, 90, T), rng.uniform(2, 60, T)
    gt = np.stack([w / 4, np.zeros(T), cx, np.zeros(T), -h / 4, cy + h], 1)         # 4 x 4 px rasters of those extents
    tb3 = TileBatch.from_arrays(np.zeros((T, 4, 4, 1), np.uint8), gt)
    boxes = []
    for _ in range(400):
        bx, by = rng.uniform(-50, 1050), rng.uniform(-50, 750)
        bw, bh = rng.uniform(0.0, 120), rng.uniform(0.0, 120)
        boxes.append([ring((bx, by), (bx + bw, by), (bx + bw, by + bh), (bx, by + bh))])
    r3 = RoadSet.from_geometries(boxes)
    got3 = eng.pairs_bbox_grid_host(r3, tb3)
    exp3 = pairs_by_bbox(r3, tb3)
    assert np.array_equal(got3.road_pair_off, exp3.road_pair_off) and np.array_equal(got3.pair_tile, exp3.pair_tile)
    assert exp3.n_pairs > 1000


def test_clip_labels_gpu_equals_host_restatement(mods):
    """determine_class.clip_labels on the GPU (bounding-box join, exact intersects reject, re-entrant Sutherland-Hodgman) against
    the numpy restatement clip_labels_host: same rows in the same order, same rings vertex for vertex (bit-exact)."""
    dc = mods[3]
    rng = np.random.default_rng(21)

    def star(c, rmin, rmax, n):
        ang = (np.arange(n) + rng.uniform(0.0, 0.8, n)) * (2 * np.pi / n)
        rad = rng.uniform(rmin, rmax, n)
        pts = np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], 1)
        return np.concatenate([pts, pts[:1]])
    labels = []
    for i in range(60):
        c = rng.uniform(-3, 63, 2)
        rings = [star(c, 2, 14, int(rng.integers(5, 40)))]
        if i % 3 == 0:
            rings.append(star(c, 0.3, 1.8, 7))
        labels.append({"type": "Polygon", "coordinates": [r.tolist() for r in rings]})
    labels.append({"type": "Polygon", "coordinates": [[[10.0, 10.0], [20.0, 10.0], [20.0, 20.0], [10.0, 20.0], [10.0, 10.0]]]})   # exactly a tile
    labels.append({"type": "Polygon", "coordinates": [[[500.0, 500.0], [501.0, 500.0], [501.0, 501.0], [500.0, 500.0]]]})        # joins no tile
    tiles, tid = [], []
    for ty in range(6):
        for tx in range(6):
            x0, y0 = 10.0 * tx, 10.0 * ty
            tiles.append({"type": "Polygon", "coordinates": [[[x0, y0], [x0 + 10, y0], [x0 + 10, y0 + 10], [x0, y0 + 10], [x0, y0]]]})
            tid.append(f"({tx}, {ty}, 18)")
    lab_df = pd.DataFrame({"OBJECTID": np.arange(len(labels)) + 1, "BELAGSART": 100, "geometry": labels})
    til_df = pd.DataFrame({"id": tid, "title": "t", "geometry": tiles})
    got = dc.clip_labels(lab_df, til_df, fact=0.99)
    exp = dc.clip_labels_host(lab_df, til_df, fact=0.99)
    assert list(got.columns) == list(exp.columns) and len(got) == len(exp) > 150
    assert got["OBJECTID"].tolist() == exp["OBJECTID"].tolist() and got["tile_id"].tolist() == exp["tile_id"].tolist()
    n_empty = 0
    for a, b in zip(got["geometry"], exp["geometry"]):
        assert len(a["coordinates"]) == len(b["coordinates"])
        n_empty += len(b["coordinates"]) == 0
        for ra, rb in zip(a["coordinates"], b["coordinates"]):
            assert np.array_equal(np.array(ra), np.array(rb))
    assert n_empty < len(exp) // 4
    table, soup = dc.clip_labels(lab_df, til_df, fact=0.99, as_soup=True)
    assert len(table) == len(exp) == soup.n_roads and "geometry" not in table.columns
