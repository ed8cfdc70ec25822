"""The two driver blocks of the reference that sit on the hot path, as single batched calls.

  road_band_statistics   scripts/statistical_analysis/statistical_analysis.py:170-270
                         (tile x road pairing, per-pair pixel extraction, per-road per-band statistics, filter)
  road_surface_vote      scripts/road_segmentation/final_metrics.py:262-316 on raster detections
                         (per-road class accumulation, vote at every threshold, tags, per-class and balanced metrics,
                         best threshold)
Inputs are the flat containers of ``geometry.py`` (shapely / GeoJSON geometries go through
``RoadSet.from_geometries``); outputs are the reference's pandas tables.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import pandas as pd

from .engine import default_engine
from .functions import fct_statistics
from .geometry import PairList, RoadSet, TileBatch, lattice_of
from .road_segmentation import determine_class, final_metrics


def pair_list(roads: RoadSet, tiles: TileBatch, engine=None, exact: bool = True) -> PairList:
    """gpd.sjoin(tiles, roads) + drop_duplicates (statistical_analysis.py:170-171) on the GPU: bounding-box broad phase (lattice
    look-up for XYZ tile sets, uniform-grid binning for any other tile set), then the exact 'intersects' reject."""
    eng = engine or default_engine()
    pairs = eng.pairs_bbox_host(roads, tiles) if lattice_of(tiles) is not None else eng.pairs_bbox_grid_host(roads, tiles)
    return eng.pairs_intersect_host(roads, tiles, pairs) if exact else pairs


def road_band_statistics(roads: RoadSet, tiles: TileBatch, pairs: Optional[PairList] = None, BANDS: Sequence[int] = (1, 2, 3),
                         COUNT_THRESHOLD=10, MAX_MOE=12.5, engine=None):
    """Returns (roads_stats, roads_stats_filtered): the table of statistical_analysis.py:235-246 (min_b, max_b, median_b,
    mean_b, std_b, margin_b per band, count) and its filtered form (:264-270).  ``tiles.nodata`` selects the nodata
    convention of get_pixel_values: None -> pixels with every band 0 are dropped (fct_misc.py:117-119), 0 -> per-band
    zeros dropped and short bands zero-padded (:95-111)."""
    eng = engine or default_engine()
    if pairs is None:
        pairs = pair_list(roads, tiles, eng)
    mode = "none" if tiles.nodata is None else ("zero" if tiles.nodata == 0 else None)
    if mode is None:
        raise ValueError("tile nodata must be None or 0 on this path (config_stats.yaml:39 requests nodata=0)")
    if max(BANDS) > tiles.channels or min(BANDS) < 1:
        raise IndexError(f"BANDS {list(BANDS)} do not fit {tiles.channels}-band tiles")
    stats = eng.zonal_stats_host(roads, tiles, pairs, nodata_mode=mode, ddof=1)
    ids = np.arange(roads.n_roads) if roads.ids is None else roads.ids
    table = fct_statistics.road_stats_from_accumulators(stats[:, [b - 1 for b in BANDS]], ids, BANDS)
    # count := count of the FIRST band (statistical_analysis.py:245), whatever BANDS[0] is
    return table, fct_statistics.filter_roads(table, BANDS, COUNT_THRESHOLD, MAX_MOE)


def road_surface_vote(roads: RoadSet, detection_tiles: TileBatch, gt_class: np.ndarray, pairs: Optional[PairList] = None,
                      thresholds=None, rule: str = "count", min_area_frac: float = 0.0, engine=None):
    """detection_tiles: 2-channel uint8 tiles, class plane (0 none, 1 artificial, 2 natural) and score plane (score * 255).
    gt_class: per road 0 artificial / 1 natural (anything else: not in the ground truth).
    Returns dict(by_class, global_metrics, best_index, best_threshold, comparison) where ``comparison`` is the
    per-road table at the best threshold (road_id, cover_type, CATEGORY, tag)."""
    eng = engine or default_engine()
    if pairs is None:
        pairs = pair_list(roads, detection_tiles, eng)
    jh = determine_class.accumulate_class_planes(roads, detection_tiles, pairs, eng)
    by_class, glob, bi, bt, cover, scores = final_metrics.threshold_sweep(jh, gt_class, thresholds, rule, min_area_frac,
                                                                          return_scores=True)
    ids = np.arange(roads.n_roads) if roads.ids is None else roads.ids
    gt = np.asarray(gt_class)
    known = (gt == 0) | (gt == 1)
    comparison = pd.DataFrame({
        "road_id": np.asarray(ids)[known],
        "cover_type": determine_class.COVER_NAMES[cover[bi][known]],
        "CATEGORY": np.where(gt[known] == 0, "artificial", "natural"),
    })
    comparison["tag"] = final_metrics.tags_from_codes(cover[bi][known], gt[known])
    # art_score / nat_score rounded to 3 decimals, diff_score as is (determine_class.py:150-167)
    comparison["art_score"] = np.round(scores[bi][known, 0], 3)
    comparison["nat_score"] = np.round(scores[bi][known, 1], 3)
    comparison["diff_score"] = scores[bi][known, 2]
    return {"by_class": by_class, "global_metrics": glob, "best_index": bi, "best_threshold": bt, "comparison": comparison,
            "joint_hist": jh}
