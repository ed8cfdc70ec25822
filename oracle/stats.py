"""Oracle: per-road statistics tables.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  scripts/functions/fct_statistics.py:44-70    get_df_stats_groupby
  scripts/functions/fct_statistics.py:72-105   get_df_stats_no_group
  scripts/statistical_analysis/statistical_analysis.py:235-246,264-270   table assembly + filter
  rasterstats 0.17.0 zonal_stats (call shape: statistical_analysis.py:221-222, fct_rasters.py:162-163)
Pinned: the two fct_statistics functions are checked against the reference's own
code executed with its plotting imports stubbed (tests/golden/stats_groupby.json,
tests/golden/stats_no_group.json).  The rasterstats restatement is **parity unpinned**
(rasterstats is not importable here).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from . import gdal_fill

Z_COEF = 2  # fct_statistics.py:58 / :98 -- "1.96 rounded up"


def get_df_stats_groupby(dataframe: pd.DataFrame, col: str, groups: List[str], suffix: str = "") -> pd.DataFrame:
    g = dataframe.groupby(groups)[col]
    out = pd.DataFrame({
        "min": g.min(), "max": g.max(), "median": g.median(),
        "mean": g.mean(), "count": g.count(), "std": g.std(),     # pandas std: ddof=1
    })
    margin = Z_COEF * out["std"] / np.sqrt(out["count"])
    out["mean"] = out["mean"].round(2)
    out["std"] = out["std"].round(2)
    out[f"margin{suffix}"] = margin.round(2)
    if suffix != "":
        out = out.rename(columns={k: f"{k}{suffix}" for k in ("min", "max", "median", "mean", "count", "std")})
    return out


def get_df_stats_no_group(dataframe: pd.DataFrame, col: str, results_dict: Optional[dict] = None,
                          suffix: str = "", to_df: bool = False):
    keys = ("min", "max", "mean", "median", "std", "count", "margin")
    if results_dict is None:
        results_dict = {f"{k}{suffix}": [] for k in keys}
    s = dataframe[col]
    std_r = np.round(s.std(), 2)
    n = s.count()
    results_dict[f"min{suffix}"].append(int(s.min()))
    results_dict[f"max{suffix}"].append(int(s.max()))
    results_dict[f"mean{suffix}"].append(np.round(s.mean(), 2))
    results_dict[f"median{suffix}"].append(s.median())
    results_dict[f"std{suffix}"].append(std_r)
    results_dict[f"count{suffix}"].append(n)
    results_dict[f"margin{suffix}"].append(np.round(Z_COEF * std_r / np.sqrt(n), decimals=3))
    return pd.DataFrame(results_dict) if to_df else results_dict


# ----------------------------------------------------------------------------
# statistics straight from 256-bin histograms (numpy, exact integers)
# ----------------------------------------------------------------------------
def expand_hist(h: np.ndarray) -> np.ndarray:
    return np.repeat(np.arange(256, dtype=np.int64), np.asarray(h, np.int64))


def stats_from_hist(h: np.ndarray, ddof: int = 1, percentiles: Sequence[float] = ()) -> Dict[str, float]:
    """min/max/median/mean/count/std (+ numpy-style percentiles) of the multiset a
    histogram describes, by expanding it -- slow and obviously right."""
    v = expand_hist(h)
    n = int(v.size)
    if n == 0:
        out = {k: float("nan") for k in ("min", "max", "median", "mean", "std")}
        out["count"] = 0
        for q in percentiles:
            out[f"percentile_{q:g}"] = float("nan")
        return out
    out = {
        "min": int(v.min()), "max": int(v.max()), "median": float(np.median(v)),
        "mean": float(v.mean()), "count": n,
        "std": float(v.std(ddof=ddof)) if n > ddof else float("nan"),
    }
    for q in percentiles:
        out[f"percentile_{q:g}"] = float(np.percentile(v, q))
    return out


def apply_nodata_convention(hist: np.ndarray, n_allzero: np.ndarray, mode: str, min_zero=None) -> np.ndarray:
    """Turn raw in-mask histograms (R, C, 256) into the multiset get_pixel_values returns
    (SURVEY.md A.4).  mode 'N': tile nodata None -> rows with every band 0 dropped.
    mode 'Z': tile nodata 0 -> per band the zeros are dropped, then the band is padded with
    zeros up to the longest band OF THE SAME (road, tile) CALL (fct_misc.py:95-111); the calls of a road are
    concatenated afterwards (statistical_analysis.py:187-193).  ``min_zero`` (R,) = sum over the road's calls of
    min over bands of the call's zero count (oracle.raster.zonal_accumulate(want_min_zero=True)); without it the
    road is taken as ONE call (exact for single-tile roads only).  mode 'raw': unchanged."""
    h = np.array(hist, dtype=np.int64, copy=True)
    if mode == "raw":
        return h
    if mode == "N":
        h[:, :, 0] -= np.asarray(n_allzero, np.int64)[:, None]
        return h
    if mode == "Z" and min_zero is not None:
        h[:, :, 0] -= np.asarray(min_zero, np.int64)[:, None]
        return h
    if mode == "Z":
        nonzero = h[:, :, 1:].sum(axis=2)                 # (R, C) = L_b
        longest = nonzero.max(axis=1, keepdims=True)
        h[:, :, 0] = longest - nonzero
        return h
    raise ValueError(mode)


def road_stats_table(hist: np.ndarray, n_allzero: np.ndarray, road_ids: Sequence, mode: str = "N",
                     bands: Sequence[int] = (1, 2, 3)) -> pd.DataFrame:
    """statistical_analysis.py:235-246 from accumulators: one row per road with pixels,
    columns min_b,max_b,median_b,mean_b,std_b,margin_b (rounded as fct_statistics.py:61-63) + count."""
    h = apply_nodata_convention(hist, n_allzero, mode)
    rows = []
    for r, rid in enumerate(road_ids):
        if h[r, 0].sum() == 0:
            continue
        row = {"road_id": rid}
        for ci, b in enumerate(bands):
            s = stats_from_hist(h[r, ci], ddof=1)
            margin = Z_COEF * s["std"] / np.sqrt(s["count"])
            row[f"min_{b}"] = s["min"]
            row[f"max_{b}"] = s["max"]
            row[f"median_{b}"] = s["median"]
            row[f"mean_{b}"] = np.round(s["mean"], 2)
            row[f"std_{b}"] = np.round(s["std"], 2)
            row[f"margin_{b}"] = np.round(margin, 2)
            if ci == 0:
                count = s["count"]
        row["count"] = count
        rows.append(row)
    return pd.DataFrame(rows)


def filter_roads(stats_df: pd.DataFrame, bands: Sequence[int], count_threshold=10, max_moe=12.5) -> pd.DataFrame:
    """statistical_analysis.py:264-270."""
    ok = np.zeros(len(stats_df), bool)
    for b in bands:
        ok |= (stats_df[f"margin_{b}"] < max_moe).to_numpy()
    keep = (stats_df["count"] > count_threshold).to_numpy() & ok
    return stats_df[keep].drop(columns=[f"margin_{b}" for b in bands] + ["count"])


# ----------------------------------------------------------------------------
# rasterstats 0.17.0 zonal_stats -- parity unpinned
# ----------------------------------------------------------------------------
def zonal_stats(vectors: Sequence[Sequence[np.ndarray]], raster: np.ndarray, affine, stats=("min", "max", "mean", "count"),
                nodata=None, percentiles: Sequence[float] = ()) -> List[dict]:
    """rasterstats.zonal_stats(vectors, array, affine=..., stats=..., nodata=...) on one band.

    raster: 2-D array; window = rowcol(bounds) floor/ceil, boundless (padded with nodata);
    mask = (value == nodata) | ~rasterized; std ddof=0."""
    import math
    H, W = raster.shape
    a, b, c, d, e, f = affine
    out = []
    for rings in vectors:
        allxy = np.concatenate([np.asarray(r, np.float64) for r in rings])
        w_, s_, e_, n_ = allxy[:, 0].min(), allxy[:, 1].min(), allxy[:, 0].max(), allxy[:, 1].max()
        inv = gdal_fill.affine_invert((a, b, c, d, e, f))

        def rowcol(x, y, op):
            fc, fr = gdal_fill.affine_apply(inv, x, y)
            return int(op(fr)), int(op(fc))
        r0, c0 = rowcol(w_, n_, math.floor)
        r1, c1 = rowcol(e_, s_, math.ceil)
        h, w = max(r1 - r0, 0), max(c1 - c0, 0)
        fill = nodata if nodata is not None else -999
        win = np.full((h, w), fill, dtype=np.float64)
        rr0, rr1, cc0, cc1 = max(r0, 0), min(r1, H), max(c0, 0), min(c1, W)
        if rr1 > rr0 and cc1 > cc0:
            win[rr0 - r0:rr1 - r0, cc0 - c0:cc1 - c0] = raster[rr0:rr1, cc0:cc1]
        wt = gdal_fill.affine_mul_translation((a, b, c, d, e, f), c0, r0)
        rv = gdal_fill.rasterize(rings, (h, w), wt).astype(bool)
        # rasterstats io.Raster: nodata None becomes -999 (with a warning) and the comparison below still
        # runs, so the padding of a boundless window is always masked
        masked = ~rv | np.isnan(win) | (win == fill)
        vals = win[~masked]
        res = {}
        if vals.size == 0:
            for s in stats:
                res[s] = 0 if s == "count" else None
            for q in percentiles:
                res[f"percentile_{q:g}"] = None
        else:
            for s in stats:
                if s == "min":
                    res[s] = float(vals.min())
                elif s == "max":
                    res[s] = float(vals.max())
                elif s == "mean":
                    res[s] = float(vals.mean())
                elif s == "count":
                    res[s] = int(vals.size)
                elif s == "std":
                    res[s] = float(vals.std())
                elif s == "median":
                    res[s] = float(np.median(vals))
                elif s == "sum":
                    res[s] = float(vals.sum())
                else:
                    raise ValueError(s)
            for q in percentiles:
                res[f"percentile_{q:g}"] = float(np.percentile(vals, q))
        out.append(res)
    return out


RATIO_NAMES = {'1/2': 'R/G', '1/3': 'R/B', '1/4': 'R/NIR', '2/3': 'G/B', '2/4': 'G/NIR', '3/4': 'B/NIR'}


def band_ratios(pixels_per_band: pd.DataFrame, BANDS: Sequence[int] = range(1, 5)) -> pd.DataFrame:
    """scripts/statistical_analysis/statistical_analysis.py:279-293: ratios between bands (float64 division, round(3), NaN -> 0,
    then non-finite -> 1) and VgNIR-BI = (band2 - band4) / (band2 + band4), round(5), NaN kept.  Same pandas statements,
    on a copy."""
    df = pixels_per_band.copy()
    for band in BANDS:
        for sec_band in range(band + 1, max(BANDS) + 1):
            name = RATIO_NAMES[f'{band}/{sec_band}']
            with np.errstate(divide='ignore', invalid='ignore'):
                df[name] = df[f'band{band}'].astype('float64') / df[f'band{sec_band}'].astype('float64')
            df[name] = df[name].round(3)
            df.loc[np.isnan(df[name]), name] = 0
            df.loc[~np.isfinite(df[name]), name] = 1
    with np.errstate(divide='ignore', invalid='ignore'):
        df['VgNIR-BI'] = (df['band2'].astype('float64') - df['band4'].astype('float64')) / (
            df['band2'].astype('float64') + df['band4'].astype('float64'))
    df['VgNIR-BI'] = df['VgNIR-BI'].round(5)
    return df
