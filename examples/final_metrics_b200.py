#!/usr/bin/env python
"""The flow of scripts/road_segmentation/final_metrics.py on the B200 path, on synthetic raster detections.

  roads in quarries are set aside             ->  determine_class.get_roads_in_quarries   (final_metrics.py:246-248)
  labels limited to the visible tile area     ->  clip_fact=0.99 (border_px of the raster accumulation, :255)
  per-road class vote for 20 thresholds,
  tags, per-class and balanced metrics,
  best threshold                              ->  workflows.road_surface_vote            (:262-316)
  threshold sweep on the score difference     ->  final_metrics.diff_score_sweep         (:429-478)
  calibration bins                            ->  final_metrics.bin_accuracy             (:541-571)

The detector (Mask R-CNN of the object-detector repository) is out of scope: its output is replaced by synthetic class
and score planes (class 0 none / 1 artificial / 2 natural, score * 255).  Needs a B200.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from proj_roadsurf_b200 import synth, workflows                                       # noqa: E402
from proj_roadsurf_b200.geometry import TileBatch                                      # noqa: E402
from proj_roadsurf_b200.road_segmentation import determine_class, final_metrics       # noqa: E402


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/roadsurf_b200_demo")
    ap.add_argument("--tiles-x", type=int, default=8)
    ap.add_argument("--tiles-y", type=int, default=6)
    ap.add_argument("--roads", type=int, default=60)
    args = ap.parse_args(argv)
    tables = os.path.join(args.out, "tables")
    os.makedirs(tables, exist_ok=True)

    grid = synth.Grid(args.tiles_x, args.tiles_y)
    rr = synth.ribbon_roads(grid, args.roads, seed=11)
    roads, gt_class = rr.roads, rr.gt_class
    ids = np.arange(roads.n_roads) if roads.ids is None else np.asarray(roads.ids)
    detections = TileBatch.from_arrays(synth.host_tiles(grid, 2, "class_score"), grid.transforms())

    print("-- Roads in quarries are always naturals...")
    x0, y0, x1, y1 = roads.bbox[:, 0].min(), roads.bbox[:, 1].min(), roads.bbox[:, 2].max(), roads.bbox[:, 3].max()
    cx, cy, w, h = (x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0) / 4, (y1 - y0) / 4
    quarry = {"type": "Polygon", "coordinates": [[[cx - w, cy - h], [cx + w, cy - h], [cx + w, cy + h], [cx - w, cy + h], [cx - w, cy - h]]]}
    geoms = [[roads.xy[roads.ring_off[g]:roads.ring_off[g + 1]] for g in range(roads.road_ring_off[r], roads.road_ring_off[r + 1])]
             for r in range(roads.n_roads)]
    roads_df = pd.DataFrame({"OBJECTID": ids, "CATEGORY": np.where(gt_class == 0, "artificial", "natural"), "geometry": geoms})
    in_quarries, filtered = determine_class.get_roads_in_quarries([quarry], roads_df)       # already buffered geometry
    print(f"{len(in_quarries)} roads lie within the quarry and are set aside")
    keep = np.nonzero(np.isin(ids, filtered["OBJECTID"].to_numpy()))[0]          # positions of the remaining roads

    print("Determining the detected class of every road for every threshold...")
    sub = roads.subset(keep)
    res = workflows.road_surface_vote(sub, detections, gt_class[keep], rule="score", min_area_frac=0.05)
    comparison = res["comparison"].copy()
    if roads.ids is None:                                    # road_id counts the roads of `sub`: back to the ids of `roads`
        comparison["road_id"] = ids[keep][comparison["road_id"].to_numpy()]
    comparison["gt_type"] = "val"
    print(f"best threshold {res['best_threshold']}: f1b = {res['global_metrics']['f1b'][res['best_index']]:.3f}")

    print("Threshold on the difference between the class scores...")
    by_class_d, global_d, best_diff, best_results = final_metrics.diff_score_sweep(comparison)
    print(f"best difference threshold {best_diff}")

    print("Calculate the bin accuracy to estimate the calibration...")
    accuracy_tables = final_metrics.bin_accuracy(comparison)

    res["by_class"].to_csv(os.path.join(tables, "by_class_metrics.csv"), index=False)
    res["global_metrics"].to_csv(os.path.join(tables, "global metrics.csv"), index=False)
    comparison.to_csv(os.path.join(tables, "comparison_best_threshold.csv"), index=False)
    return {"vote": res, "comparison": comparison, "in_quarries": in_quarries, "diff_sweep": (by_class_d, global_d, best_diff, best_results),
            "accuracy_tables": accuracy_tables, "roads": roads, "detections": detections, "gt_class": gt_class, "keep": keep}


if __name__ == "__main__":
    main()
