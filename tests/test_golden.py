"""The oracle's table logic against fixtures produced by the REFERENCE's own Python functions
(tests/golden/make_golden.py ran them in the build container; see its docstring for what is stubbed)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import raster as oraster, stats as ostats, vote as ovote

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def frame(j, dtypes=None):
    df = pd.DataFrame(j["data"], columns=j["columns"], index=j["index"])
    if dtypes:
        df = df.astype(dtypes)
    return df


def same(a, b):
    """golden cell vs computed cell: None == NaN, floats at 1e-12 relative, everything else exactly"""
    if a is None:
        return b is None or (isinstance(b, float) and b != b) or (hasattr(b, "dtype") and np.isnan(b))
    if isinstance(a, float) or isinstance(b, (float, np.floating)):
        return abs(float(a) - float(b)) <= 1e-12 * max(1.0, abs(float(a)))
    return a == b


def assert_frame_matches(golden, df):
    assert list(map(str, df.columns)) == golden["columns"]
    assert [int(i) if isinstance(i, (int, np.integer)) else i for i in df.index.tolist()] == golden["index"]
    got = df.to_numpy().tolist()
    for r, (ga, gb) in enumerate(zip(golden["data"], got)):
        for c, (x, y) in enumerate(zip(ga, gb)):
            assert same(x, y), (r, golden["columns"][c], x, y)


def test_stats_groupby_matches_reference():
    g = load("stats_groupby")
    px = frame(g["pixels"], {"band1": np.uint8, "band2": np.uint8})
    for case in g["cases"]:
        res = ostats.get_df_stats_groupby(px, case["col"], ["road_id"], case["suffix"])
        assert_frame_matches(case["result"], res)


def test_stats_no_group_matches_reference():
    g = load("stats_no_group")
    px = frame(g["pixels"], {"band1": np.uint8, "band2": np.uint8})
    d = None
    for grp in g["groups"]:
        d = ostats.get_df_stats_no_group(px.loc[grp["rows"]], "band1", d, "_1")
    for k, v in g["result"].items():
        assert len(v) == len(d[k])
        for x, y in zip(v, d[k]):
            assert same(x, y), (k, x, y)
    assert_frame_matches(g["as_df"], ostats.get_df_stats_no_group(px, "band2", None, "", True))


def test_vote_tags_and_metrics_match_reference():
    g = load("vote")
    roads = frame(g["roads"])
    preds = frame(g["predictions"])
    assert len(g["sweep"]) == 20
    covers = set()
    for step in g["sweep"]:
        comp = ovote.determine_detected_class(preds, roads, step["threshold"])
        comp["tag"] = [ovote.get_tag(c, k) for c, k in zip(comp["cover_type"], comp["CATEGORY"])]
        gold = frame(step["comparison"])
        assert comp["road_id"].tolist() == gold["road_id"].tolist()
        assert comp["cover_type"].tolist() == gold["cover_type"].tolist()
        assert comp["tag"].tolist() == gold["tag"].tolist()
        for col in ("nat_score", "art_score", "diff_score"):
            np.testing.assert_allclose(comp[col].astype(float), gold[col].astype(float), rtol=1e-12, atol=1e-15)
        by_class, glob = ovote.get_metrics(comp)
        assert_frame_matches(step["by_class"], by_class)
        assert_frame_matches(step["global"], glob)
        covers |= set(gold["cover_type"])
    assert covers == {"artificial", "natural", "undetermined", "undetected"}


def test_get_pixel_values_wrapper_matches_reference():
    g = load("pixel_values")
    data = np.array(g["data"], np.uint8)
    from oracle import gdal_fill
    acc = {None: pd.DataFrame(), 0: pd.DataFrame()}
    n_rows = 0
    for case in g["cases"]:
        nodata = case["nodata"]
        tile = {"data": data, "transform": tuple(g["transform"]), "nodata": nodata}
        if case["geom"] == "__accumulated__":
            assert_frame_matches(case["result"], acc[nodata])
            continue
        rings = gdal_fill.rings_from_geojson(g["geoms"][case["geom"]])
        one = oraster.get_pixel_values(rings, tile, range(1, 4), pd.DataFrame(), road_id=case["geom"])
        assert_frame_matches(case["result"], one)
        acc[nodata] = oraster.get_pixel_values(rings, tile, range(1, 4), acc[nodata], road_id=case["geom"])
        n_rows += len(one)
    assert n_rows > 100
    assert g["missing_tile_rows"] == 0 and len(oraster.get_pixel_values([], None)) == 0


def test_multi_tile_nodata_padding_matches_reference():
    """A road over several tiles, tile nodata 0 / None: the accumulator form of the oracle (histograms + n_allzero + min_zero,
    apply_nodata_convention) reproduces the statistics the reference's get_pixel_values + get_df_stats_groupby give when the
    calls of a road are concatenated (statistical_analysis.py:187-193, :238-246)."""
    from oracle import gdal_fill
    g = load("multi_tile")
    tiles = np.array(g["tiles"], np.uint8)
    tr = np.array(g["transforms"], np.float64)
    names = list(g["geoms"])
    rings = [gdal_fill.rings_from_geojson(g["geoms"][n]) for n in names]
    pairs = [(t, r) for r in range(len(names)) for t in range(len(tiles))]
    hist, nz, mz = oraster.zonal_accumulate(tiles, tr, rings, pairs, want_min_zero=True)
    assert (mz != nz).any()
    for case in g["cases"]:
        mode = "Z" if case["nodata"] == 0 else "N"
        adj = ostats.apply_nodata_convention(hist, nz, mode, min_zero=mz)
        for b in (1, 2, 3):
            gold = case["stats"][str(b)]
            cols = gold["columns"]
            for name, row in zip(gold["index"], gold["data"]):
                s = ostats.stats_from_hist(adj[names.index(name), b - 1], ddof=1)
                exp = dict(zip(cols, row))
                assert s["count"] == exp[f"count_{b}"] and s["min"] == exp[f"min_{b}"] and s["max"] == exp[f"max_{b}"]
                assert s["median"] == exp[f"median_{b}"]
                assert round(s["mean"], 2) == exp[f"mean_{b}"] and round(s["std"], 2) == exp[f"std_{b}"]
        if mode == "Z":       # the road-level simplification (one call per road) is NOT what the reference computes
            wrong = ostats.apply_nodata_convention(hist, nz, "Z")
            assert (wrong[:, :, 0] != adj[:, :, 0]).any()


def test_vote_many_detections_near_ties_match_reference():
    g = load("vote_many")
    comp = ovote.determine_detected_class(frame(g["predictions"]), frame(g["roads"]), 0.0)
    gold = frame(g["comparison"])
    assert comp["cover_type"].tolist() == gold["cover_type"].tolist()
    np.testing.assert_allclose(comp["diff_score"].astype(float), gold["diff_score"].astype(float), rtol=1e-12, atol=0)
