"""Mirror of the vote helpers of scripts/road_segmentation/determine_class.py.

Two forms of the per-road vote:
  * ``determine_detected_class(predictions, roads, threshold)`` -- the reference's table form (:122-190):
    rows of (OBJECTID, score, det_class_name, weighted_score, area_pred_in_label) are summed per road and
    class on the GPU (rs_vote_table_host) and the class indices compared;
  * ``raster_vote(...)`` / ``accumulate_class_planes(...)`` -- the raster form named by the north star: class
    and score planes are counted inside each road polygon by the fused kernel (joint histogram of
    (class, score) per road), then argmax of the pixel counts ('count') or of the mean scores ('score').
The GEOS overlay of get_weighted_scores (:97-120) is replaced by the raster accumulation; the vector
preprocessing helpers keep their host (GEOS) buffering; the 'within' join of get_roads_in_quarries (:41-62) runs on the
GPU (rs_within_host) and clip_labels is the ``border_px`` window option of the raster accumulation.
"""
from __future__ import annotations

import sys
from typing import Optional, Sequence

import numpy as np
import pandas as pd

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("roadsurf_b200")

from ..engine import default_engine
from ..geometry import PairList, RoadSet, TileBatch

COVER_NAMES = np.array(["artificial", "natural", "undetermined", "undetected"], dtype=object)
CLASS_CODE = {"artificial": 0, "natural": 1}


def get_corresponding_class(row, labels_id):
    """determine_class.py:19-28: detector class id 0 / 1 -> name of labels id 1 / 2."""
    if row['det_class'] == 0:
        return labels_id.loc[labels_id['id'] == 1, 'name'].item()
    elif row['det_class'] == 1:
        return labels_id.loc[labels_id['id'] == 2, 'name'].item()
    else:
        logger.error(f"Unexpected class: {row['det_class']}")
        sys.exit(1)


def determine_category(row):
    """determine_class.py:30-39: BELAGSART 100 -> artificial, 200 -> natural."""
    if row['BELAGSART'] == 100:
        return 'artificial'
    if row['BELAGSART'] == 200:
        return 'natural'
    else:
        logger.error(f"Unexpected class: {row['BELAGSART']}")
        sys.exit(1)


def get_roads_in_quarries(quarries, roads, engine=None):
    """determine_class.py:41-62: the roads lying within a quarry buffered by 5 (gpd.sjoin(roads, buffered, predicate='within'))
    and the other roads.  ``roads``: table with 'OBJECTID' and 'geometry' columns.  ``quarries``: a GeoDataFrame (buffered
    by 5 in its own CRS and reprojected to EPSG:4326 with geopandas, as the reference does -- GEOS / PROJ preprocessing that
    stays on the host) or, where geopandas is not installed, a table / list of ALREADY buffered geometries in the CRS of the
    roads.  The 'within' predicate of every (road, quarry) combination runs on the GPU (rs_within_host).
    Returns (roads_in_quarries, roads_not_in_quarries): the join rows (road columns + 'index_right' + quarry columns) and
    the remaining roads with a fresh index."""
    if hasattr(quarries, 'buffer') and hasattr(quarries, 'to_crs'):
        buffered = quarries.copy()
        buffered['geometry'] = buffered.buffer(5)
        buffered = buffered.to_crs(epsg=4326)
        from ..functions import fct_misc
        fct_misc.test_crs(roads.crs, buffered.crs)
    elif isinstance(quarries, pd.DataFrame):
        buffered = quarries
    else:
        buffered = pd.DataFrame({'geometry': list(quarries)})
    rs_a = RoadSet.from_geometries(list(roads['geometry']))
    rs_b = RoadSet.from_geometries(list(buffered['geometry']))
    hit = (engine or default_engine()).within_host(rs_a, rs_b)
    ia, ib = np.nonzero(hit)                                     # row-major: road order, then quarry order (sjoin's order)
    left = roads.iloc[ia]
    right = buffered.drop(columns=['geometry']).iloc[ib]
    joined = left.copy()
    joined['index_right'] = buffered.index.to_numpy()[ib]
    for c in right.columns:
        name = c if c not in left.columns else f'{c}_right'
        joined[name] = right[c].to_numpy()
    in_ids = joined['OBJECTID'].unique().tolist()
    roads_not_in_quarries = roads[~roads['OBJECTID'].isin(in_ids)].reset_index(drop=True)
    return joined, roads_not_in_quarries


def determine_detected_class(predictions, roads, threshold=0):
    """Per road: detections with score >= threshold vote; index_k = sum(weighted_score) / sum(area_pred_in_label)
    per class (0 if absent or the weighted sum is 0); artificial > natural -> 'artificial', < -> 'natural',
    == -> 'undetermined', no valid detection -> 'undetected'.  Returns the reference's comparison table
    (road_id, cover_type, nat_score, art_score, diff_score, OBJECTID, geometry [, CATEGORY, gt_type])."""
    return determine_detected_class_sweep(predictions, roads, [threshold])[0]


def determine_detected_class_sweep(predictions, roads, thresholds: Sequence[float]):
    """All thresholds of the sweep (final_metrics.py:277-316) in one GPU call; one comparison table each."""
    road_ids = pd.unique(roads['OBJECTID'])
    R = len(road_ids)
    code_of = pd.Series(np.arange(R), index=road_ids)
    pred_code = predictions['OBJECTID'].map(code_of)
    known = pred_code.notna().to_numpy()
    p = predictions[known]
    codes = pred_code[known].to_numpy().astype(np.int64)
    order = np.argsort(codes, kind="stable")                      # group by road, keep the table order inside a road
    codes = codes[order]
    row_off = np.zeros(R + 1, np.int64)
    np.add.at(row_off, codes + 1, 1)
    row_off = np.cumsum(row_off)
    cls = p['det_class_name'].map(CLASS_CODE).fillna(-1).to_numpy().astype(np.int8)[order]
    score = p['score'].to_numpy(np.float64)[order]
    weighted = p['weighted_score'].to_numpy(np.float64)[order]
    area = p['area_pred_in_label'].to_numpy(np.float64)[order]
    cover, scores = default_engine().vote_table_host(row_off, cls, score, weighted, area, np.asarray(thresholds, np.float64))

    columns_to_keep = ['OBJECTID'] + (['geometry'] if 'geometry' in roads.columns else [])
    if 'gt_type' in roads.columns:
        columns_to_keep.extend(['CATEGORY', 'gt_type'])
    out = []
    for t in range(len(thresholds)):
        detected = cover[t] != 3
        final_type_df = pd.DataFrame({
            'road_id': road_ids,
            'cover_type': COVER_NAMES[cover[t]],
            'nat_score': [round(float(v), 3) if d else 0 for v, d in zip(scores[t, :, 1], detected)],
            'art_score': [round(float(v), 3) if d else 0 for v, d in zip(scores[t, :, 0], detected)],
            'diff_score': scores[t, :, 2],
        })
        out.append(final_type_df.merge(roads[columns_to_keep], how='inner', left_on='road_id', right_on='OBJECTID'))
    return out


# ------------------------------------------------------------------------------------------
# raster form
# ------------------------------------------------------------------------------------------
def clip_border_px(tile_size: int, fact: float = 0.99) -> int:
    """Raster form of clip_labels (determine_class.py:62-95), which intersects the labels with every tile scaled by
    ``fact`` about its centre: the number of outermost pixel rows / columns whose centres fall outside the scaled
    tile (1 for 256 px tiles, 5 for 1024 px tiles at fact = 0.99)."""
    import math
    margin = tile_size * (1.0 - fact) / 2.0          # 1.28 px for 256 px tiles
    return max(0, int(math.ceil(margin - 0.5)))


def accumulate_class_planes(roads: RoadSet, tiles: TileBatch, pairs: PairList, engine=None, clip_fact: Optional[float] = None) -> np.ndarray:
    """Joint (class, score) histogram per road, uint32 (R, 3, 256): tiles hold two uint8 channels, the class
    plane (0 none, 1 artificial, 2 natural) and the score plane (score * 255).  ``clip_fact`` (0.99 in the
    reference) ignores the tile border like clip_labels does."""
    eng = engine or default_engine()
    border = 0 if clip_fact is None else clip_border_px(tiles.width, clip_fact)
    hist, _ = eng.zonal_hist_host(roads, tiles, pairs, hist_mode="class_score", border_px=border)
    return hist


def _clip_ring_to_rect(ring: np.ndarray, x0: float, y0: float, x1: float, y1: float) -> np.ndarray:
    """Sutherland-Hodgman: a ring clipped to the closed axis-aligned rectangle, one half-plane at a time.  Pieces of a
    concave ring that leave and re-enter the rectangle stay connected by zero-width runs along the rectangle's edge, which
    carry no area (even-odd fill and the shoelace sum ignore them)."""
    pts = np.asarray(ring, np.float64)[:, :2]
    if len(pts) > 1 and pts[0, 0] == pts[-1, 0] and pts[0, 1] == pts[-1, 1]:
        pts = pts[:-1]
    for axis, bound, keep_ge in ((0, x0, True), (0, x1, False), (1, y0, True), (1, y1, False)):
        if len(pts) == 0:
            break
        v = pts[:, axis]
        inside = v >= bound if keep_ge else v <= bound
        nxt = np.roll(pts, -1, axis=0)
        nin = np.roll(inside, -1)
        out = []
        for k in range(len(pts)):
            p, q = pts[k], nxt[k]
            if inside[k]:
                out.append(p)
            if inside[k] != nin[k]:                              # the edge crosses the boundary line
                t = (bound - p[axis]) / (q[axis] - p[axis])
                c = p + t * (q - p)
                c[axis] = bound
                out.append(c)
        pts = np.array(out, np.float64).reshape(-1, 2)
    if len(pts) >= 3:
        pts = np.concatenate([pts, pts[:1]])                      # closed, like shapely's rings
    return pts


def clip_labels(labels_gdf, tiles_gdf, fact=0.99, engine=None, as_soup: bool = False):
    """determine_class.py:62-95 (copied there from the object detector's helpers): every label joined to every tile it
    intersects and cut to that tile scaled by ``fact`` about its centre (shapely.affinity.scale's default origin).
    ``labels_gdf`` / ``tiles_gdf``: tables with a 'geometry' column; tile geometries are axis-aligned rectangles (XYZ tiles).
    All three steps run on the GPU: the bounding-box join (rs_pairs_bbox_grid_host), the exact 'intersects' reject
    (rs_pairs_intersect_host) and the clip of every ring to its rectangle (rs_clip_rings_host, re-entrant Sutherland-Hodgman;
    same areas and raster as GEOS' intersection, see ``_clip_ring_to_rect`` for the host restatement the tests hold it to).
    Returns the joined table: label columns + tile columns ('id' renamed 'tile_id'), geometry = the clipped label as a
    GeoJSON-like dict; labels that only reach the outer 1 % frame of a tile keep an empty geometry, as in the reference.
    as_soup=True skips the per-row dicts (1 M labels) and returns (table without geometry, RoadSet of the clipped labels)."""
    from ..geometry import RoadSet, rings_of
    if hasattr(labels_gdf, 'crs') and hasattr(tiles_gdf, 'crs'):
        assert (labels_gdf.crs == tiles_gdf.crs)
    eng = engine or default_engine()
    labels = labels_gdf if isinstance(labels_gdf, RoadSet) else RoadSet.from_geometries(list(labels_gdf['geometry']))
    tile_rings = [rings_of(g) for g in tiles_gdf['geometry']]
    tb = np.array([[np.concatenate(r)[:, 0].min(), np.concatenate(r)[:, 1].min(), np.concatenate(r)[:, 0].max(), np.concatenate(r)[:, 1].max()]
                   if r else [np.inf, np.inf, -np.inf, -np.inf] for r in tile_rings], np.float64).reshape(-1, 4)
    pairs = eng.pairs_intersect_host(labels, tb, eng.pairs_bbox_grid_host(labels, tb))     # sjoin(labels, tiles, 'intersects')
    ia, ib = pairs.road_of_pair().astype(np.int64), pairs.pair_tile.astype(np.int64)
    cx, cy = (tb[ib, 0] + tb[ib, 2]) / 2, (tb[ib, 1] + tb[ib, 3]) / 2
    hw, hh = (tb[ib, 2] - tb[ib, 0]) / 2 * fact, (tb[ib, 3] - tb[ib, 1]) / 2 * fact
    clipped = eng.clip_rings_host(labels, ia, np.stack([cx - hw, cy - hh, cx + hw, cy + hh], 1))
    left_src = labels_gdf if not isinstance(labels_gdf, RoadSet) else pd.DataFrame({"label": np.arange(labels.n_roads)})
    left = left_src.drop(columns=[c for c in ['geometry'] if c in left_src.columns]).iloc[ia].reset_index(drop=True)
    right = tiles_gdf.drop(columns=['geometry']).iloc[ib].reset_index(drop=True).rename(columns={'id': 'tile_id'})
    dup = set(left.columns) & set(right.columns)
    left = left.rename(columns={c: f'{c}_left' for c in dup})
    right = right.rename(columns={c: f'{c}_right' for c in dup})
    out = pd.concat([left, right], axis=1)
    if as_soup:
        return out, clipped
    out['geometry'] = [{"type": "Polygon", "coordinates": [r.tolist() for r in clipped.rings(k)]} for k in range(clipped.n_roads)]
    return out


def clip_labels_host(labels_gdf, tiles_gdf, fact=0.99):
    """Host restatement of ``clip_labels`` (numpy, one ring at a time): what the GPU form is held to, vertex for vertex.
    determine_class.py:62-95 (copied there from the object detector's helpers): every label joined to every tile it
    intersects and cut to that tile scaled by ``fact`` about its centre (shapely.affinity.scale's default origin).
    ``labels_gdf`` / ``tiles_gdf``: tables with a 'geometry' column; tile geometries are axis-aligned rectangles (XYZ tiles).
    Vector preprocessing on the host, like the reference (GEOS there; a rectangle clip of every ring here, which gives the
    same areas -- the geometry comes back as a GeoJSON-like dict whose rings follow the even-odd rule, see
    ``_clip_ring_to_rect``).  Returns the joined table: label columns + tile columns ('id' renamed 'tile_id'), geometry =
    the clipped label; labels that only reach the outer 1 % frame of a tile keep an empty geometry, as in the reference."""
    from ..geometry import bbox_pairs, rings_of
    if hasattr(labels_gdf, 'crs') and hasattr(tiles_gdf, 'crs'):
        assert (labels_gdf.crs == tiles_gdf.crs)
    lab_rings = [rings_of(g) for g in labels_gdf['geometry']]
    tile_rings = [rings_of(g) for g in tiles_gdf['geometry']]

    def box(rings):
        v = np.concatenate(rings) if rings else np.zeros((0, 2))
        return [v[:, 0].min(), v[:, 1].min(), v[:, 0].max(), v[:, 1].max()] if len(v) else [np.inf, np.inf, -np.inf, -np.inf]
    lb = np.array([box(r) for r in lab_rings], np.float64).reshape(-1, 4)
    tb = np.array([box(r) for r in tile_rings], np.float64).reshape(-1, 4)
    ia, ib = bbox_pairs(lb, tb)
    keep, geoms = [], []
    for k, (i, j) in enumerate(zip(ia.tolist(), ib.tolist())):
        x0, y0, x1, y1 = tb[j]
        # 'intersects' of the spatial join: some part of the label lies in the closed tile
        if not any(len(_clip_ring_to_rect(r, x0, y0, x1, y1)) for r in lab_rings[i]):
            continue
        cx, cy, hw, hh = (x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0) / 2 * fact, (y1 - y0) / 2 * fact
        rings = [c for c in (_clip_ring_to_rect(r, cx - hw, cy - hh, cx + hw, cy + hh) for r in lab_rings[i]) if len(c) >= 4]
        keep.append(k)
        geoms.append({"type": "Polygon", "coordinates": [c.tolist() for c in rings]})
    ia, ib = ia[keep], ib[keep]
    left = labels_gdf.drop(columns=['geometry']).iloc[ia].reset_index(drop=True)
    right = tiles_gdf.drop(columns=['geometry']).iloc[ib].reset_index(drop=True).rename(columns={'id': 'tile_id'})
    dup = set(left.columns) & set(right.columns)
    left = left.rename(columns={c: f'{c}_left' for c in dup})
    right = right.rename(columns={c: f'{c}_right' for c in dup})
    out = pd.concat([left, right], axis=1)
    out['geometry'] = geoms
    return out


def get_weighted_scores(ground_truth, predictions, engine=None):
    """determine_class.py:97-120 in its vector form: the overlay of the labels with the predictions, the share of every label
    covered by every prediction (rounded to 2 decimals) and the confidence weighted by it; pairs covering 5 % of the label
    or less are dropped.  ``ground_truth`` / ``predictions``: tables with a 'geometry' column (anything with
    __geo_interface__, GeoJSON dicts or ring lists), 'BELAGSART' on the labels and 'score' on the predictions.  The areas of
    gpd.overlay(how='intersection') and GeoSeries.area come from one GPU call (rs_overlay_area_host) over the bounding-box
    candidate pairs; rows are in (label, prediction) order like the overlay's.  The result carries the columns of both
    tables (duplicated names get _1 / _2) plus area_label, joined_area, area_pred_in_label, weighted_score; the geometry of
    the intersections is not materialised (``geometry`` is None), nothing downstream reads it."""
    from ..geometry import bbox_pairs
    if hasattr(ground_truth, 'crs') and hasattr(predictions, 'crs'):
        from ..functions import fct_misc
        fct_misc.test_crs(ground_truth.crs, predictions.crs)
    eng = engine or default_engine()
    a = RoadSet.from_geometries(list(ground_truth['geometry']))
    b = RoadSet.from_geometries(list(predictions['geometry']))
    ia, ib = bbox_pairs(a.bbox, b.bbox)
    joined, area_label = eng.overlay_area_host(a, b, ia, ib)
    ground_truth['area_label'] = area_label                       # the reference adds the column to its argument too (:107)
    keep = joined > 0.0                                           # overlay keeps polygonal intersections only
    ia, ib, joined = ia[keep], ib[keep], joined[keep]
    left = ground_truth.drop(columns=['geometry']).iloc[ia].reset_index(drop=True)
    right = predictions.drop(columns=['geometry']).iloc[ib].reset_index(drop=True)
    dup = set(left.columns) & set(right.columns)
    left = left.rename(columns={c: f'{c}_1' for c in dup})
    right = right.rename(columns={c: f'{c}_2' for c in dup})
    out = pd.concat([left, right], axis=1)
    out['geometry'] = None
    out = out[(~out['BELAGSART'].isna()) & (~out['score'].isna())].copy()
    out['joined_area'] = joined[out.index.to_numpy()]
    out['area_pred_in_label'] = round(out['joined_area'] / out['area_label'], 2)
    out['weighted_score'] = out['area_pred_in_label'] * out['score']
    return out[out.area_pred_in_label > 0.05].copy()


def get_weighted_scores_raster(roads: RoadSet, instance_tiles: TileBatch, pairs: PairList, inst_score, inst_class_name,
                               road_ids=None, clip_fact: Optional[float] = None, min_area: float = 0.05, engine=None) -> pd.DataFrame:
    """Instance-faithful raster form of get_weighted_scores (determine_class.py:97-120).

    ``instance_tiles``: one uint16 channel holding the id of the detection covering each pixel (0 = none);
    ``inst_score[i]`` / ``inst_class_name[i]``: confidence and class name of detection i.  The GEOS overlay areas become
    pixel counts: area_label = pixels of the road polygon, joined_area = pixels of detection i inside it;
    area_pred_in_label = round(joined / label, 2), weighted_score = area_pred_in_label * score, rows with
    area_pred_in_label <= 0.05 dropped (:115-118).  The in-mask pixels come from the GPU extraction
    (rs_extract_pixels_host); the result feeds determine_detected_class unchanged."""
    eng = engine or default_engine()
    if instance_tiles.channels != 1:
        raise ValueError("instance tiles have one channel (the detection id of every pixel)")
    ids = np.arange(roads.n_roads) if road_ids is None else np.asarray(road_ids)
    px = np.asarray(instance_tiles.pixels)
    if clip_fact is not None:                    # clip_labels: ignore the tile border (ids there count as outside the label)
        b = clip_border_px(instance_tiles.width, clip_fact)
        inner = np.zeros(px.shape, bool)
        inner[:, b:px.shape[1] - b, b:px.shape[2] - b] = True
    pair_off, values = eng.extract_pixels_host(roads, instance_tiles, pairs, window="crop")
    road_of_pixel = np.repeat(pairs.road_of_pair().astype(np.int64), np.diff(pair_off))
    inst = values[:, 0].astype(np.int64)
    if clip_fact is not None:
        # the same extraction on a 0/1 plane tells which extracted pixels lie in the inner rectangle
        flag = TileBatch(inner.astype(np.uint8), instance_tiles.gt, instance_tiles.height, instance_tiles.width, 1)
        _, inside = eng.extract_pixels_host(roads, flag, pairs, window="crop")
        keep = inside[:, 0] != 0
        road_of_pixel, inst = road_of_pixel[keep], inst[keep]
    area_label = np.bincount(road_of_pixel, minlength=roads.n_roads)
    key = road_of_pixel * 65536 + inst
    uk, cnt = np.unique(key[inst > 0], return_counts=True)
    r, i = uk // 65536, uk % 65536
    frac = np.array([round(c / area_label[rr], 2) for c, rr in zip(cnt.tolist(), r.tolist())], float)
    score = np.asarray(inst_score, float)[i] if len(i) else np.zeros(0)
    keep = frac > min_area
    names = np.asarray(inst_class_name, object)
    return pd.DataFrame({"OBJECTID": ids[r[keep]], "instance": i[keep], "score": score[keep],
                         "det_class_name": names[i[keep]] if len(i) else np.zeros(0, object),
                         "area_pred_in_label": frac[keep], "weighted_score": frac[keep] * score[keep]})


def detections_to_planes(detections, tiles: TileBatch, engine=None, batch_pairs: int = 1024):
    """Burn detection polygons into the raster inputs of the vote: the detector's output is a table of polygons with
    ``score`` and ``det_class`` (0 / 1, determine_class.py:19-28) in the CRS of the tiles.  Every (detection, tile) pair is
    rasterized on the GPU with the same pixel-centre fill as the roads (rs_rasterize_pairs_host); detections are laid down in
    ascending score order, so where two overlap the more confident one is kept.
    Returns (class_score, instance): TileBatch (T, H, W, 2) uint8 -- class plane 0 none / 1 artificial / 2 natural and score
    plane rint(score * 255) -- for accumulate_class_planes / road_surface_vote, and TileBatch (T, H, W, 1) uint16 holding
    1 + the row number of the detection for get_weighted_scores_raster."""
    eng = engine or default_engine()
    n = len(detections)
    if n >= 65535:
        raise ValueError("at most 65534 detections per call (uint16 instance plane)")
    T, H, W = tiles.n_tiles, tiles.height, tiles.width
    cs = np.zeros((T, H, W, 2), np.uint8)
    inst = np.zeros((T, H, W, 1), np.uint16)
    if n:
        score = np.asarray(detections['score'], float)
        det_class = np.asarray(detections['det_class'])
        if not np.isin(det_class, (0, 1)).all():
            logger.error(f"Unexpected class: {det_class[~np.isin(det_class, (0, 1))][0]}")
            sys.exit(1)
        order = np.argsort(score, kind="stable")
        dets = RoadSet.from_geometries([list(detections['geometry'])[i] for i in order])
        from ..geometry import pairs_by_bbox
        pairs = pairs_by_bbox(dets, tiles)
        road_of = pairs.road_of_pair()
        q = np.rint(score * 255.0).clip(0, 255).astype(np.uint8)
        for lo in range(0, pairs.n_pairs, batch_pairs):            # masks are (pairs, H, W) bytes: bounded batches
            hi = min(pairs.n_pairs, lo + batch_pairs)
            sub = PairList.from_pairs(dets.n_roads, road_of[lo:hi], pairs.pair_tile[lo:hi])
            masks = eng.rasterize_pairs_host(dets, tiles.gt, H, W, sub, window="crop")
            sub_road = sub.road_of_pair()
            for k in range(sub.n_pairs):                           # pair order = ascending score
                d = int(order[sub_road[k]])
                m = masks[k] != 0
                t = int(sub.pair_tile[k])
                cs[t, :, :, 0][m] = 1 + int(det_class[d])
                cs[t, :, :, 1][m] = q[d]
                inst[t, :, :, 0][m] = d + 1
    return (TileBatch(cs, tiles.gt, H, W, 2, None, tiles.ids), TileBatch(inst, tiles.gt, H, W, 1, None, tiles.ids))


def score_cutoffs(thresholds: Sequence[float]) -> np.ndarray:
    """smallest uint8 score s with s / 255 >= threshold (256: none)"""
    s = np.arange(256) / 255.0
    return np.array([int(np.argmax(s >= t)) if (s >= t).any() else 256 for t in thresholds], np.int32)


def raster_vote(joint_hist: np.ndarray, gt_class: Optional[np.ndarray], thresholds: Sequence[float], rule: str = "count",
                min_area_frac: float = 0.0, engine=None):
    """Vote, tags, confusion counts and metrics for every threshold in one launch (rs_vote_metrics_host).
    Returns cover (T, R) int8 codes, scores (T, R, 3), confusion (T, 2, 4), metrics (T, 12)."""
    eng = engine or default_engine()
    return eng.vote_metrics_host(joint_hist, gt_class, score_cutoffs(thresholds), rule=rule, min_area_frac=min_area_frac)
