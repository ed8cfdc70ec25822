"""Generate the golden fixtures of tests/golden/ by executing the REFERENCE's own Python functions.

Runs only in the build container (it imports /root/reference, which does not exist on the GPU box);
the JSON files it writes are committed and are what the tests read.

The reference modules import geopandas / shapely / rasterio / plotly / matplotlib at top level, none of
which is installed here.  They are stubbed in sys.modules just enough for the imports to succeed:
  * table functions (fct_statistics.get_df_stats_groupby / get_df_stats_no_group,
    determine_class.determine_detected_class, final_metrics.get_tag / get_metrics) never touch the
    stubs -- what runs is 100 % reference code on pandas/numpy, so these fixtures PIN the oracle;
  * fct_misc.get_pixel_values calls rasterio.open + rasterio.mask.mask; the stub answers with the oracle's
    restatement (oracle/raster.py mask_crop), so that fixture pins the reference's wrapper logic
    (np.extract per band, padding, all-zero row drop, concat) on top of an UNPINNED rasterization.
pandas here is 3.0 (reference pin 1.5.1); numpy 2.3 (pin 1.23.4).
"""
import json
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import raster as oraster  # noqa: E402

TILES = {}          # path -> tile dict {'data','transform','nodata'}


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class GeoDataFrame(pd.DataFrame):
        @property
        def _constructor(self):
            return GeoDataFrame
    mod("geopandas", GeoDataFrame=GeoDataFrame)
    mod("shapely")
    mod("shapely.geometry", mapping=lambda g: g)              # geometries are passed as GeoJSON dicts already
    mod("shapely.affinity", scale=lambda *a, **k: None)

    class RasterioIOError(Exception):
        pass

    class _Src:
        def __init__(self, path):
            if path not in TILES:
                raise RasterioIOError(path)
            self.path = path
            self.nodata = TILES[path].get("nodata")

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def _mask(src, geoms, crop=True):
        from oracle import gdal_fill
        rings = []
        for g in geoms:
            rings.extend(gdal_fill.rings_from_geojson(g))
        out = oraster.mask_crop(TILES[src.path], rings)
        if out is None:
            raise ValueError("Input shapes do not overlap raster.")
        return out, None
    rio = mod("rasterio", open=_Src)
    rio.mask = mod("rasterio.mask", mask=_mask)
    rio.errors = mod("rasterio.errors", RasterioIOError=RasterioIOError)
    mod("plotly")
    mod("plotly.graph_objects")
    mod("plotly.express")
    mod("matplotlib")
    mod("matplotlib.pyplot")
    mod("matplotlib.colors")


def to_jsonable(v):
    if isinstance(v, pd.DataFrame):
        return {"columns": list(map(str, v.columns)), "index": [to_jsonable(i) for i in v.index.tolist()],
                "data": [[to_jsonable(x) for x in row] for row in v.to_numpy().tolist()]}
    if isinstance(v, dict):
        return {str(k): to_jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [to_jsonable(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating, float)):
        return None if (v != v) else float(v)
    if isinstance(v, np.ndarray):
        return to_jsonable(v.tolist())
    return v


def ring(*pts):
    pts = [list(map(float, p)) for p in pts]
    if pts[0] != pts[-1]:
        pts.append(pts[0])
    return pts


def main():
    install_stubs()
    sys.path.insert(1, os.path.join(REF, "scripts"))
    sys.path.insert(1, os.path.join(REF, "scripts", "road_segmentation"))
    os.chdir(REF)
    import functions.fct_misc as fct_misc
    import functions.fct_statistics as fct_statistics
    import determine_class
    import final_metrics

    rng = np.random.default_rng(20261018)
    out = {}

    # ---------------- fct_statistics ----------------
    n = 400
    px = pd.DataFrame({"road_id": rng.integers(1, 9, n), "band1": rng.integers(0, 256, n).astype(np.uint8),
                       "band2": np.clip(rng.normal(110, 6, n), 0, 255).astype(np.uint8)})
    px.loc[px["road_id"] == 7, "band2"] = 93                          # constant group: std 0
    px = pd.concat([px, pd.DataFrame({"road_id": [99], "band1": np.array([17], np.uint8), "band2": np.array([200], np.uint8)})],
                   ignore_index=True)                                 # singleton group: std NaN
    cases = []
    for col, suffix in (("band1", "_1"), ("band2", "_2"), ("band1", "")):
        res = fct_statistics.get_df_stats_groupby(px, col, ["road_id"], suffix)
        cases.append({"col": col, "suffix": suffix, "result": to_jsonable(res)})
    out["stats_groupby"] = {"pixels": to_jsonable(px), "cases": cases}

    ng = []
    d = None
    for rt in (100, 200):
        sub = px[px["road_id"] % 2 == (0 if rt == 100 else 1)]
        d = fct_statistics.get_df_stats_no_group(sub, "band1", d, "_1")
        ng.append({"road_type": rt, "rows": sub.index.tolist()})
    out["stats_no_group"] = {"pixels": to_jsonable(px), "groups": ng, "result": to_jsonable(d),
                             "as_df": to_jsonable(fct_statistics.get_df_stats_no_group(px, "band2", None, "", True))}

    # ---------------- determine_class / final_metrics ----------------
    roads = pd.DataFrame({"OBJECTID": np.arange(1, 41), "geometry": [None] * 40,
                          "CATEGORY": np.where(rng.random(40) < 0.75, "artificial", "natural"),
                          "gt_type": ["gt"] * 40})
    rows = []
    for rid in roads["OBJECTID"]:
        if rid % 9 == 0:
            continue                                                   # undetected road
        for _ in range(int(rng.integers(1, 5))):
            area = round(float(rng.uniform(0.06, 0.9)), 2)
            score = round(float(rng.uniform(0.05, 1.0)), 3)
            cls = "artificial" if rng.random() < 0.6 else "natural"
            rows.append({"OBJECTID": int(rid), "score": score, "det_class_name": cls, "area_pred_in_label": area,
                         "weighted_score": area * score})
    # an exact tie -> 'undetermined'
    rows.append({"OBJECTID": 40, "score": 0.5, "det_class_name": "artificial", "area_pred_in_label": 0.2, "weighted_score": 0.1})
    rows.append({"OBJECTID": 40, "score": 0.5, "det_class_name": "natural", "area_pred_in_label": 0.4, "weighted_score": 0.2})
    preds = pd.DataFrame(rows)
    preds = preds[~((preds["OBJECTID"] == 40) & (preds.index < len(rows) - 2))]
    sweep = []
    for thr in np.arange(0, 1.0, 0.05):
        comp = determine_class.determine_detected_class(preds, roads, thr)
        comp["tag"] = comp.apply(lambda row: final_metrics.get_tag(row), axis=1)
        by_class, glob = final_metrics.get_metrics(comp, ["artificial", "natural"])
        sweep.append({"threshold": float(thr), "comparison": to_jsonable(pd.DataFrame(comp).drop(columns=["geometry"])),
                      "by_class": to_jsonable(by_class), "global": to_jsonable(glob)})
    out["vote"] = {"roads": to_jsonable(roads.drop(columns=["geometry"])), "predictions": to_jsonable(preds), "sweep": sweep}

    # many detections per (road, class) with near-ties: pandas' groupby sum is Kahan-compensated, a plain running sum is not
    rng2 = np.random.default_rng(77)
    roads_k = pd.DataFrame({"OBJECTID": np.arange(1, 61), "geometry": [None] * 60,
                            "CATEGORY": np.where(rng2.random(60) < 0.5, "artificial", "natural"), "gt_type": ["gt"] * 60})
    rows_k = []
    for rid in roads_k["OBJECTID"]:
        m = int(rng2.integers(20, 60))
        area = np.round(rng2.uniform(0.06, 0.9, m), 2)
        score = np.round(rng2.uniform(0.3, 1.0, m), 3)
        perm = rng2.permutation(m)
        for cls, order in (("artificial", np.arange(m)), ("natural", perm)):      # the same multiset in two orders: an exact tie
            for i in order:                                                          # only for order-insensitive sums
                rows_k.append({"OBJECTID": int(rid), "score": float(score[i]), "det_class_name": cls,
                               "area_pred_in_label": float(area[i]), "weighted_score": float(area[i] * score[i])})
    preds_k = pd.DataFrame(rows_k).sample(frac=1.0, random_state=3).sort_values("OBJECTID", kind="stable").reset_index(drop=True)
    comp_k = determine_class.determine_detected_class(preds_k, roads_k, 0.0)
    # the fixture must separate compensated from plain summation
    naive_differs = 0
    for rid, grp in preds_k.groupby("OBJECTID"):
        idx = {}
        for cls, sub in grp.groupby("det_class_name"):
            sw = sa = 0.0
            for w, a in zip(sub["weighted_score"], sub["area_pred_in_label"]):
                sw += w; sa += a
            idx[cls] = sw / sa
        naive = "undetermined" if idx["artificial"] == idx["natural"] else ("artificial" if idx["artificial"] > idx["natural"] else "natural")
        naive_differs += naive != comp_k.loc[comp_k["road_id"] == rid, "cover_type"].iloc[0]
    assert naive_differs > 0, "fixture does not separate Kahan from plain sums"
    out["vote_many"] = {"roads": to_jsonable(roads_k.drop(columns=["geometry"])), "predictions": to_jsonable(preds_k),
                        "comparison": to_jsonable(pd.DataFrame(comp_k).drop(columns=["geometry"])), "naive_differs": int(naive_differs)}

    # ---------------- fct_misc.get_pixel_values (rasterio stubbed by the oracle) ----------------
    data = rng.integers(0, 256, (12, 16, 3), dtype=np.uint8)
    data[rng.random((12, 16)) < 0.15] = 0                              # all-band zeros
    data[..., 1][rng.random((12, 16)) < 0.1] = 0                       # zeros on one band only
    t = (0.5, 0.0, 1000.0, 0.0, -0.5, 5000.0)
    pv = []
    geoms = {
        "rect": {"type": "Polygon", "coordinates": [ring((1001.2, 4995.1), (1005.7, 4995.1), (1005.7, 4998.9), (1001.2, 4998.9))]},
        "holed": {"type": "Polygon", "coordinates": [ring((1000.5, 4994.5), (1007.5, 4994.5), (1007.5, 4999.5), (1000.5, 4999.5)),
                                                     ring((1003.0, 4996.0), (1003.0, 4998.0), (1005.0, 4998.0), (1005.0, 4996.0))]},
        "multi": {"type": "MultiPolygon", "coordinates": [[ring((1000.2, 4999.8), (1002.2, 4999.8), (1002.2, 4997.3), (1000.2, 4997.3))],
                                                          [ring((1004.4, 4996.2), (1007.9, 4995.1), (1006.3, 4994.2))]]},
        "partly_outside": {"type": "Polygon", "coordinates": [ring((995.0, 4990.0), (1002.0, 4990.0), (1002.0, 4997.0), (995.0, 4997.0))]},
    }
    for nodata in (None, 0):
        TILES["18_1_1.tif"] = {"data": data, "transform": t, "nodata": nodata}
        acc = pd.DataFrame()
        for name, g in geoms.items():
            one = fct_misc.get_pixel_values(g, "18_1_1.tif", range(1, 4), pd.DataFrame(), road_id=name)
            acc = fct_misc.get_pixel_values(g, "18_1_1.tif", range(1, 4), acc, road_id=name)
            pv.append({"nodata": nodata, "geom": name, "result": to_jsonable(one)})
        pv.append({"nodata": nodata, "geom": "__accumulated__", "result": to_jsonable(acc)})
    missing = fct_misc.get_pixel_values(geoms["rect"], "nope.tif", range(1, 4), pd.DataFrame(), road_id=1)
    out["pixel_values"] = {"data": data.tolist(), "transform": list(t), "geoms": geoms, "cases": pv,
                           "missing_tile_rows": int(len(missing))}

    # ---------------- a road over several tiles with tile nodata 0: the zero padding is per (road, tile) call ----------------
    # (fct_misc.py:95-111 pads inside one call; statistical_analysis.py:187-193 concatenates the calls; :238-246 groupby)
    mt_tiles, mt_t = [], []
    for k in range(3):
        d = rng.integers(1, 256, (10, 12, 3), dtype=np.uint8)
        d[..., k][rng.random((10, 12)) < 0.15 + 0.1 * k] = 0          # a different band loses pixels on every tile
        for c in range(3):
            d[..., c][rng.random((10, 12)) < 0.08] = 0                 # independent zeros on every band: min over bands > all-zero rows
        d[rng.random((10, 12)) < 0.05] = 0
        mt_tiles.append(d)
        mt_t.append((0.5, 0.0, 2000.0 + 6.0 * k, 0.0, -0.5, 7000.0))   # three tiles side by side (12 px * 0.5)
    mt_geoms = {
        "long": {"type": "Polygon", "coordinates": [ring((2000.7, 6996.2), (2017.4, 6996.9), (2017.4, 6998.8), (2000.7, 6998.1))]},
        "two_tiles": {"type": "Polygon", "coordinates": [ring((2004.2, 6995.4), (2009.9, 6995.4), (2009.9, 6999.6), (2004.2, 6999.6))]},
    }
    mt_cases = []
    for nodata in (0, None):
        acc = pd.DataFrame()
        for name, g in mt_geoms.items():
            pv_road = pd.DataFrame()
            for k in range(3):
                path = f"18_9_{k}.tif"
                TILES[path] = {"data": mt_tiles[k], "transform": mt_t[k], "nodata": nodata}
                try:
                    pv_road = fct_misc.get_pixel_values(g, path, range(1, 4), pv_road, road_id=name)
                except ValueError:                                     # "Input shapes do not overlap raster."
                    pass
            acc = pd.concat([acc, pv_road], ignore_index=True)
        st = {}
        for b in (1, 2, 3):
            st[str(b)] = to_jsonable(fct_statistics.get_df_stats_groupby(acc, f"band{b}", ["road_id"], f"_{b}"))
        mt_cases.append({"nodata": nodata, "pixels": to_jsonable(acc), "stats": st})
    out["multi_tile"] = {"tiles": [d.tolist() for d in mt_tiles], "transforms": [list(t) for t in mt_t], "geoms": mt_geoms,
                         "cases": mt_cases}

    for k, v in out.items():
        with open(os.path.join(HERE, f"{k}.json"), "w") as f:
            json.dump(v, f)
        print("wrote", k, os.path.getsize(os.path.join(HERE, f"{k}.json")), "bytes")


if __name__ == "__main__":
    main()
