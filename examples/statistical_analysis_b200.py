#!/usr/bin/env python
"""The flow of scripts/statistical_analysis/statistical_analysis.py on the B200 path, end to end, on synthetic inputs.

  tiles on disk (GeoTIFF, "<z>_<x>_<y>.tif")  ->  ingest.load_tiles                     (fct_misc.py:76, rasterio.open per pair)
  tiles x roads                               ->  workflows.pair_list                    (statistical_analysis.py:170-171)
  per-road, per-band statistics + filter      ->  workflows.road_band_statistics         (:179-270)
  pixels_per_band table                       ->  fct_misc.get_pixel_values_batch        (:180-193)
  band ratios, VgNIR-BI                       ->  fct_statistics.add_band_ratios         (:279-293)
  statistics per road type                    ->  fct_statistics.cover_stats_from_accumulators   (:296-316)
  Kolmogorov-Smirnov per road and band        ->  fct_statistics.ks_test_from_hists      (:436-461)
  elevation statistics per road over a DEM    ->  fct_rasters.zonal_stats                (scripts/functions/fct_rasters.py:147-167)

The reference ships no imagery (data/readme.md), so the script first writes synthetic 4-band tiles and ribbon roads.
Needs a B200 (there is no CPU path).   python examples/statistical_analysis_b200.py --out /tmp/roadsurf_demo
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

from proj_roadsurf_b200 import ingest, synth, workflows                     # noqa: E402
from proj_roadsurf_b200.engine import default_engine                         # noqa: E402
from proj_roadsurf_b200.functions import fct_misc, fct_rasters, fct_statistics as fs      # noqa: E402


def write_synthetic_tiles(grid: synth.Grid, folder: str) -> list:
    """4-band uint8 tiles as deflate GeoTIFFs named like the reference's tiles (statistical_analysis.py:138-139)."""
    from PIL import Image, TiffImagePlugin
    os.makedirs(folder, exist_ok=True)
    tiles = synth.host_tiles(grid, 4, "asphalt")
    gts = grid.transforms()
    paths = []
    for i in range(grid.n_tiles):
        x, y, z = grid.x0 + i % grid.nx, grid.y0 + i // grid.nx, grid.z
        ifd = TiffImagePlugin.ImageFileDirectory_v2()
        t = gts[i]
        ifd[33550] = (float(t[0]), float(-t[4]), 0.0); ifd.tagtype[33550] = 12
        ifd[33922] = (0.0, 0.0, 0.0, float(t[2]), float(t[5]), 0.0); ifd.tagtype[33922] = 12
        p = os.path.join(folder, f"{z}_{x}_{y}.tif")
        Image.fromarray(tiles[i], "RGBA").save(p, tiffinfo=ifd, compression="tiff_adobe_deflate")
        paths.append(p)
    return paths


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/roadsurf_b200_demo")
    ap.add_argument("--tiles-x", type=int, default=8)
    ap.add_argument("--tiles-y", type=int, default=6)
    ap.add_argument("--roads", type=int, default=40)
    ap.add_argument("--no-ks", action="store_true")
    args = ap.parse_args(argv)
    BANDS = range(1, 5)
    COUNT_THRESHOLD, MAX_MOE = 10, 12.5
    tables = fct_misc.ensure_dir_exists(os.path.join(args.out, "tables"))
    written_files = []

    grid = synth.Grid(args.tiles_x, args.tiles_y)
    paths = write_synthetic_tiles(grid, os.path.join(args.out, "tiles"))
    rr = synth.ribbon_roads(grid, args.roads, seed=7)
    roads = rr.roads
    road_type = np.where(rr.gt_class == 0, 100, 200)                # BELAGSART 100 artificial / 200 natural
    ids = np.arange(roads.n_roads) if roads.ids is None else roads.ids

    print("Reading the tiles...")
    tiles = ingest.load_tiles(paths)                                 # one kernel assembles the whole batch
    tiles.nodata = None                                              # get_pixel_values, mode "no nodata" (fct_misc.py:117-119)

    print("Getting the intersecting area between tiles and roads...")
    pairs = workflows.pair_list(roads, tiles)

    print("Calculating the statistics per band and road...")
    roads_stats, roads_stats_filtered = workflows.road_band_statistics(roads, tiles, pairs, BANDS, COUNT_THRESHOLD, MAX_MOE)
    roads_stats["road_type"] = road_type[np.searchsorted(ids, roads_stats["road_id"].to_numpy())]
    roads_stats_filtered["road_type"] = road_type[np.searchsorted(ids, roads_stats_filtered["road_id"].to_numpy())]
    print(f"{roads_stats.shape[0] - roads_stats_filtered.shape[0]} roads on {roads_stats.shape[0]} were dropped because they "
          f"contained less than {COUNT_THRESHOLD} pixels or their margin of error was higher than {MAX_MOE} on one or many bands.")
    for name, df in (("stats_roads.csv", roads_stats), ("stats_roads_filtered.csv", roads_stats_filtered)):
        df.to_csv(os.path.join(tables, name), index=False)
        written_files.append(os.path.join(tables, name))

    print("Extracting the pixels and calculating ratios between bands...")
    pixels_per_band = fct_misc.get_pixel_values_batch(roads, tiles, pairs, BANDS, road_ids=ids)
    pixels_per_band["road_type"] = road_type[np.searchsorted(ids, pixels_per_band["road_id"].to_numpy())]
    pixels_per_band = fs.add_band_ratios(pixels_per_band, BANDS)

    print("Calculating the statistics per band and cover...")
    hist, n_allzero = default_engine().zonal_hist_host(roads, tiles, pairs)
    cover_stats_df = fs.cover_stats_from_accumulators(hist, n_allzero, road_type, BANDS)
    cover_stats_df.to_csv(os.path.join(tables, "statistics_roads_by_type.csv"), index=False)
    written_files.append(os.path.join(tables, "statistics_roads_by_type.csv"))

    ks = None
    if not args.no_ks:
        print("Executing the Kolmogorov-Smirnov test...")
        keep = np.isin(ids, roads_stats_filtered["road_id"].to_numpy())
        ks = roads_stats_filtered.reset_index(drop=True).copy()
        h = hist.copy()
        h[:, :, 0] -= np.minimum(h[:, :, 0], n_allzero[:, None])     # the pixel table drops all-zero pixels
        for b in BANDS:
            # every road against ALL pixels of its type (pixels_per_band is not filtered, :447), reported for the kept roads
            res = fs.ks_test_from_hists(h[:, b - 1], road_type)
            ks[f"ks_p_band{b}"] = res["ks_p"].to_numpy()[keep]
            ks[f"ks_D_band{b}"] = res["ks_D"].to_numpy()[keep]
        ks.to_csv(os.path.join(tables, "ks_test.csv"), index=False)
        written_files.append(os.path.join(tables, "ks_test.csv"))

    print("Calculating zonal stats over the DEM...")
    # fct_rasters.py:147-167: a float32 elevation raster (2 m grid, nodata -9999) under the road polygons
    X0, Y1 = grid.origin
    dem_res = grid.span / 64.0
    dh, dw = args.tiles_y * 64, args.tiles_x * 64
    yy, xx = np.mgrid[0:dh, 0:dw]
    rng = np.random.default_rng(11)
    dem_array = (430.0 + 0.02 * xx + 0.05 * yy + rng.normal(0.0, 0.3, (dh, dw))).astype(np.float32)
    dem_array[rng.random((dh, dw)) < 0.01] = -9999.0
    affine = (dem_res, 0.0, X0, 0.0, -dem_res, Y1)
    labels = [roads.rings(r) for r in range(roads.n_roads)]
    zs_df = pd.DataFrame(fct_rasters.zonal_stats(labels, dem_array, affine=affine, stats=['min', 'max', 'mean', 'median', 'std'],
                                                 nodata=-9999))
    zs_roads = pd.concat([pd.DataFrame({"road_id": ids}), zs_df], axis=1)
    zs_roads.to_csv(os.path.join(tables, "roads_dem_zs.csv"), index=False)
    written_files.append(os.path.join(tables, "roads_dem_zs.csv"))

    print("The following files were written:")
    for f in written_files:
        print(f)
    return {"roads_stats": roads_stats, "roads_stats_filtered": roads_stats_filtered, "pixels_per_band": pixels_per_band,
            "cover_stats": cover_stats_df, "ks": ks, "tiles": tiles, "roads": roads, "pairs": pairs, "road_type": road_type,
            "dem_stats": zs_roads, "dem": (dem_array, affine)}


if __name__ == "__main__":
    main()
