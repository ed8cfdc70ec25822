"""Mirror of scripts/functions/fct_misc.py for the overlay path.

``get_pixel_values`` keeps the reference signature and return value (fct_misc.py:57-123); the mask and the
ordered pixel extraction run on the GPU (rs_extract_pixels_host), the reference's nodata handling
(:87-119) is applied to the extracted rows.  ``get_pixel_values_batch`` is the batched form the double loop
of statistical_analysis.py:180-193 collapses into.  A tile is either registered in memory
(``register_tile``), opened through rasterio when that is installed, or read by the package's own GeoTIFF ingest
(``proj_roadsurf_b200.ingest``; PIL for the TIFF layouts it does not cover).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import numpy as np
import pandas as pd

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("roadsurf_b200")

from ..engine import default_engine
from ..geometry import PairList, RoadSet, TileBatch, rasterio_window

_TILES: Dict[str, dict] = {}


def format_logger(logger):
    """fct_misc.py:16-26 -- four level-specific stderr sinks."""
    logger.remove()
    logger.add(sys.stderr, format="{time:YYYY-MM-DD HH:mm:ss} - {level} - {message}",
               level="INFO", filter=lambda record: record["level"].no < 25)
    logger.add(sys.stderr, format="{time:YYYY-MM-DD HH:mm:ss} - <green>{level}</green> - {message}",
               level="SUCCESS", filter=lambda record: record["level"].no < 30)
    logger.add(sys.stderr, format="{time:YYYY-MM-DD HH:mm:ss} - <yellow>{level}</yellow> - {message}",
               level="WARNING", filter=lambda record: record["level"].no < 40)
    logger.add(sys.stderr, format="{time:YYYY-MM-DD HH:mm:ss} - <red>{level}</red> - <level>{message}</level>",
               level="ERROR")
    return logger


def test_crs(crs1, crs2="EPSG:2056"):
    """fct_misc.py:28-41 -- print and sys.exit(1) on mismatch.  Accepts anything with a ``crs`` attribute."""
    crs1 = getattr(crs1, "crs", crs1)
    crs2 = getattr(crs2, "crs", crs2)
    try:
        assert (crs1 == crs2), f"CRS mismatch between the two files ({crs1} vs {crs2})."
    except Exception as e:  # noqa: BLE001
        print(e)
        sys.exit(1)


test_crs.__test__ = False       # not a pytest test


def ensure_dir_exists(dirpath):
    """fct_misc.py:43-54"""
    if not os.path.exists(dirpath):
        os.makedirs(dirpath)
        print(f"The directory {dirpath} was created.")
    return dirpath


# ------------------------------------------------------------------------------------------
# tiles
# ------------------------------------------------------------------------------------------
def register_tile(path: str, data: np.ndarray, transform, nodata=None, layout: str = "HWC") -> None:
    """Make an in-memory tile available under ``path`` (what rasterio.open(path) would read)."""
    a = np.asarray(data)
    if a.ndim == 2:
        a = a[..., None]
    elif layout == "CHW":
        a = np.moveaxis(a, 0, 2)
    _TILES[path] = {"data": np.ascontiguousarray(a), "transform": tuple(float(v) for v in transform)[:6], "nodata": nodata}


def clear_tiles() -> None:
    _TILES.clear()


def open_tile(tile) -> Optional[dict]:
    """dict {'data' (H, W, C), 'transform' (a, b, c, d, e, f), 'nodata'} of a tile given as such a dict, a
    registered path, or a file rasterio can open; None where the reference catches RasterioIOError."""
    if isinstance(tile, dict):
        return tile
    if tile in _TILES:
        return _TILES[tile]
    try:
        import rasterio                       # not part of this image; used when the deployment has it
    except Exception:  # noqa: BLE001
        return _open_geotiff(tile)
    try:
        with rasterio.open(tile) as src:
            t = src.transform
            return {"data": np.ascontiguousarray(np.moveaxis(src.read(), 0, 2)), "transform": (t.a, t.b, t.c, t.d, t.e, t.f),
                    "nodata": src.nodata}
    except Exception:  # noqa: BLE001  (RasterioIOError)
        return None


def _open_geotiff(path) -> Optional[dict]:
    """The package's own ingest (ingest.load_tiles: directory parse + inflate on the host, predictor / interleave on the
    GPU); layouts it does not cover go through PIL; None for a missing or unreadable file (the reference's RasterioIOError
    branch, fct_misc.py:83-85)."""
    from .. import ingest
    if not os.path.exists(path):
        return None
    try:
        tb = ingest.load_tiles([path], threads=1)
    except Exception:  # noqa: BLE001  (a layout outside the ingest's scope, or a host without the GPU: decode with PIL)
        return _open_geotiff_pil(path)
    return {"data": tb.pixels[0], "transform": tuple(float(v) for v in tb.gt[0]), "nodata": tb.nodata}


def _open_geotiff_pil(path) -> Optional[dict]:
    """Minimal GeoTIFF reader for north-up 8-bit tiles (what the tile server / generate_tilesets writes:
    "<z>_<x>_<y>.tif", statistical_analysis.py:138-139): pixels through PIL, the affine from ModelPixelScale (33550) +
    ModelTiepoint (33922) or ModelTransformation (34264), nodata from GDAL_NODATA (42113)."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            data = np.asarray(im)
            tags = im.tag_v2 if hasattr(im, "tag_v2") else {}
            scale, tie, mt, nd = tags.get(33550), tags.get(33922), tags.get(34264), tags.get(42113)
    except Exception:  # noqa: BLE001  (missing or unreadable file: the RasterioIOError branch of the reference)
        return None
    if data.ndim == 2:
        data = data[..., None]
    if data.dtype not in (np.uint8, np.uint16):
        return None
    if scale is not None and tie is not None:
        i, j, _, x, y, _ = [float(v) for v in tie[:6]]
        sx, sy = float(scale[0]), float(scale[1])
        transform = (sx, 0.0, x - i * sx, 0.0, -sy, y + j * sy)
    elif mt is not None:
        m = [float(v) for v in mt]
        transform = (m[0], m[1], m[3], m[4], m[5], m[7])
    else:
        transform = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
    nodata = None
    if nd is not None:
        try:
            nodata = float(str(nd).strip().strip("\x00"))
            nodata = int(nodata) if nodata == int(nodata) else nodata
        except ValueError:
            nodata = None
    return {"data": np.ascontiguousarray(data), "transform": transform, "nodata": nodata}


# ------------------------------------------------------------------------------------------
# get_pixel_values
# ------------------------------------------------------------------------------------------
def _frames_from_rows(values: np.ndarray, no_data, BANDS, tile_name: str, kwargs: dict) -> pd.DataFrame:
    """fct_misc.py:87-119 on the ordered in-mask rows `values` (n, C)."""
    bands = list(BANDS)
    dico = {}
    length_bands = []
    for band in bands:
        data = values[:, band - 1]
        # outside the mask rasterio fills with nodata (or 0): those pixels never reach this table when the
        # tile has a nodata value, and are dropped by the all-zero rule below when it has none
        val = data if no_data is None else data[data != no_data]
        dico[f"band{band}"] = val
        length_bands.append(len(val))
    max_length = max(length_bands) if length_bands else 0
    for band in bands:
        n = length_bands[band - 1]                 # the reference indexes with band-1 (BANDS starts at 1)
        if n < max_length:
            dico[f"band{band}"] = np.append(dico[f"band{band}"], [no_data] * (max_length - n))
            logger.warning(f"{max_length - n} pixels was/were missing on the band {band} on the tile {tile_name[-18:]} and"
                           + f" got replaced with the value used of no data ({no_data}).")
    dico.update(**kwargs)
    frame = pd.DataFrame(dico)
    if no_data is None:
        cols = [f"band{band}" for band in bands]
        frame = frame.drop(frame[frame[cols].max(axis=1) == 0].index)
    return frame


def get_pixel_values(geoms, tile, BANDS=range(1, 4), pixel_values=pd.DataFrame(), **kwargs):
    """Pixels of ``tile`` under the polygon ``geoms`` as DataFrame rows ``band{b}`` (+ kwargs columns), appended
    to ``pixel_values`` -- same signature, row order (row-major) and nodata quirks as fct_misc.py:57-123.
    A missing tile logs an error and returns an empty DataFrame (:83-85); shapes that miss the raster raise
    ValueError like rasterio.mask.mask."""
    t = open_tile(tile)
    if t is None:
        logger.error(f"The tile {tile} not found")
        return pd.DataFrame()
    data = np.asarray(t["data"])
    roads = RoadSet.from_geometries([geoms])
    tb = TileBatch.from_arrays(data[None], np.asarray(t["transform"], np.float64)[None], t.get("nodata"))
    pairs = PairList.from_pairs(1, [0], [0])
    eng = default_engine()
    if roads.n_verts == 0 or rasterio_window(tb.gt[0], roads.bbox[0], tb.width, tb.height) is None:
        raise ValueError("Input shapes do not overlap raster.")
    _, values = eng.extract_pixels_host(roads, tb, pairs, window="crop")
    frame = _frames_from_rows(values, t.get("nodata"), BANDS, tile if isinstance(tile, str) else "<memory>", kwargs)
    return pd.concat([pixel_values, frame], ignore_index=True)


def table_from_rows(values: np.ndarray, pair_off: np.ndarray, no_data, BANDS, pair_ids, tile_names=None) -> pd.DataFrame:
    """The concatenation of ``_frames_from_rows`` over every pair as ONE DataFrame, built column-wise (no per-pair frames):
    values (n, C) = the ordered in-mask rows of all pairs, pair_off (P + 1,) their slices, pair_ids (P,) the ``road_id`` of
    every pair.  Same rows, order, padding, dtypes and warnings as the per-pair loop (fct_misc.py:87-121 under
    statistical_analysis.py:180-193)."""
    bands = list(BANDS)
    pair_off = np.asarray(pair_off, np.int64)
    P = len(pair_off) - 1
    n_in = np.diff(pair_off)
    pair_of_row = np.repeat(np.arange(P), n_in)
    pair_ids = np.asarray(pair_ids)
    if no_data is None:
        keep = values[:, [b - 1 for b in bands]].max(axis=1) != 0 if len(values) else np.zeros(0, bool)
        cols = {f"band{b}": values[keep, b - 1] for b in bands}
        cols["road_id"] = pair_ids[pair_of_row[keep]]
        return pd.DataFrame(cols)
    # nodata value: per (pair, band) the pixels equal to it are dropped, then every band of the pair is padded with it up to the
    # pair's longest band
    kept, cnt = {}, {}
    for b in bands:
        k = values[:, b - 1] != no_data
        kept[b] = k
        cs = np.concatenate([[0], np.cumsum(k)])
        cnt[b] = cs[pair_off[1:]] - cs[pair_off[:-1]]
    # the reference indexes length_bands with band - 1 (BANDS starts at 1): keep that
    length_bands = [cnt[b] for b in bands]
    longest = np.max(np.stack(length_bands), axis=0) if bands else np.zeros(P, np.int64)
    out_off = np.concatenate([[0], np.cumsum(longest)])
    total = int(out_off[-1])
    cols = {}
    for b in bands:
        n_b = length_bands[b - 1]                  # IndexError for band lists that do not start at 1, like the reference
        padded = n_b < longest
        fill = np.asarray([no_data])
        dt = np.result_type(values.dtype, fill.dtype) if padded.any() else values.dtype
        col = np.full(total, no_data, dt)
        k = kept[b]
        rank = np.cumsum(k) - 1 - np.repeat(np.concatenate([[0], np.cumsum(cnt[b])])[:-1], n_in)       # rank inside the pair
        col[(out_off[:-1][pair_of_row] + rank)[k]] = values[k, b - 1]
        cols[f"band{b}"] = col
        for p in np.nonzero(padded)[0]:
            name = str(p) if tile_names is None else str(tile_names[p])
            logger.warning(f"{int(longest[p] - n_b[p])} pixels was/were missing on the band {b} on the tile {name[-18:]} and"
                           + f" got replaced with the value used of no data ({no_data}).")
    cols["road_id"] = np.repeat(pair_ids, longest)
    return pd.DataFrame(cols)


def get_pixel_values_batch(roads: RoadSet, tiles: TileBatch, pairs: PairList, BANDS=range(1, 4), road_ids=None,
                           engine=None) -> pd.DataFrame:
    """The whole double loop of statistical_analysis.py:180-193 in one call: the concatenated
    ``pixels_per_band`` table (columns band{b}, road_id), roads in order, tiles in pair order inside a road.
    One extraction call (count pass + ordered write pass on the device), one DataFrame built column-wise."""
    eng = engine or default_engine()
    pair_off, values = eng.extract_pixels_host(roads, tiles, pairs, window="crop")
    ids = np.arange(roads.n_roads) if road_ids is None else np.asarray(road_ids)
    if pairs.n_pairs == 0:
        return pd.DataFrame()
    return table_from_rows(values, pair_off, tiles.nodata, BANDS, ids[pairs.road_of_pair()])


logger = format_logger(logger) if hasattr(logger, "remove") else logger
