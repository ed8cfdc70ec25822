"""Mirror of the statistics helpers of scripts/functions/fct_statistics.py (:44-105).

The reference groups a table of pixel rows with pandas; here the grouped column is folded into 256-bin
histograms on the GPU (rs_group_hist_host) and the statistics come from the histogram kernel
(rs_finalize_stats_host, ddof = 1 like pandas ``std``).  Rounding and column naming follow the reference.
``road_stats_from_accumulators`` builds the same table straight from the fused kernel's per-road
histograms, without ever materialising pixel rows (what bench.py times).
The plotting / PCA helpers of the reference file are presentation code and are not part of this path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import pandas as pd

from .._native import STAT_COLS
from ..engine import default_engine

Z = 2       # "Coefficient of 1.96 rounded up" (fct_statistics.py:58, :98)
_COL = {k: i for i, k in enumerate(STAT_COLS)}


def _uint8_column(series: pd.Series, col: str) -> np.ndarray:
    a = series.to_numpy()
    if a.dtype == np.uint8:
        return a
    if np.issubdtype(a.dtype, np.integer) or np.issubdtype(a.dtype, np.floating):
        if len(a) == 0 or (np.all(a == np.floor(a)) and a.min() >= 0 and a.max() <= 255):
            return a.astype(np.uint8)
    raise TypeError(f"column {col!r}: the GPU statistics path takes 8-bit pixel values (uint8 band columns)")


def _stats_table(values: np.ndarray, codes: np.ndarray, n_groups: int, engine=None) -> np.ndarray:
    eng = engine or default_engine()
    hist = eng.group_hist_host(values, codes, n_groups)
    return eng.finalize_stats_host(hist[:, None, :], None, nodata_mode="raw", ddof=1)[:, 0, :]


def get_df_stats_groupby(dataframe, col, groups, suffix=''):
    """min, max, median, mean, count, std (ddof 1) of ``col`` per group + margin = Z*std/sqrt(count);
    mean, std, margin rounded to 2 decimals; columns suffixed -- fct_statistics.py:44-70."""
    keys = dataframe[groups[0]] if len(groups) == 1 else pd.MultiIndex.from_frame(dataframe[groups])
    codes, uniques = pd.factorize(keys, sort=True)
    st = _stats_table(_uint8_column(dataframe[col], col), codes.astype(np.int32), len(uniques))
    index = pd.Index(uniques, name=groups[0]) if len(groups) == 1 else pd.MultiIndex.from_tuples(list(uniques), names=groups)
    src_dtype = dataframe[col].dtype
    stats_df = pd.DataFrame({
        'min': st[:, _COL['min']].astype(src_dtype), 'max': st[:, _COL['max']].astype(src_dtype),
        'median': st[:, _COL['median']], 'mean': st[:, _COL['mean']], 'count': st[:, _COL['count']].astype(np.int64),
        'std': st[:, _COL['std']],
    }, index=index)
    stats_df[f'margin{suffix}'] = Z * stats_df['std'] / (stats_df['count'] ** (1 / 2))
    stats_df['mean'] = stats_df['mean'].round(2)
    stats_df['std'] = stats_df['std'].round(2)
    stats_df[f'margin{suffix}'] = stats_df[f'margin{suffix}'].round(2)
    if suffix != '':
        stats_df.rename(columns={k: f'{k}{suffix}' for k in ('min', 'max', 'median', 'mean', 'count', 'std')}, inplace=True)
    return stats_df


def get_df_stats_no_group(dataframe, col, results_dict=None, suffix='', to_df=False):
    """Same statistics over the whole column, appended to a dict of lists -- fct_statistics.py:72-105."""
    if results_dict is None:
        results_dict = {f'min{suffix}': [], f'max{suffix}': [], f'mean{suffix}': [], f'median{suffix}': [],
                        f'std{suffix}': [], f'count{suffix}': [], f'margin{suffix}': []}
    values = _uint8_column(dataframe[col], col)
    st = _stats_table(values, np.zeros(len(values), np.int32), 1)[0]
    results_dict[f'min{suffix}'].append(int(st[_COL['min']]))
    results_dict[f'max{suffix}'].append(int(st[_COL['max']]))
    results_dict[f'mean{suffix}'].append(np.float64(st[_COL['mean']]).round(2))
    results_dict[f'median{suffix}'].append(float(st[_COL['median']]))
    results_dict[f'std{suffix}'].append(np.float64(st[_COL['std']]).round(2))
    results_dict[f'count{suffix}'].append(int(st[_COL['count']]))
    results_dict[f'margin{suffix}'].append(np.round(Z * results_dict[f'std{suffix}'][-1] / (results_dict[f'count{suffix}'][-1] ** (1 / 2)),
                                                    decimals=3))
    if to_df:
        return pd.DataFrame(results_dict)
    return results_dict


# ------------------------------------------------------------------------------------------
# batched forms on the fused kernel's accumulators
# ------------------------------------------------------------------------------------------
def road_stats_from_accumulators(stats: np.ndarray, road_ids: Sequence, BANDS: Sequence[int] = (1, 2, 3)) -> pd.DataFrame:
    """statistical_analysis.py:235-246 from the statistics table of rs_zonal_stats / rs_finalize_stats
    (shape (R, C, RS_NSTAT [+ percentiles])): one row per road that has pixels, columns
    min_b, max_b, median_b, mean_b, std_b, margin_b per band, then ``count`` (= count of band 1)."""
    stats = np.asarray(stats)
    keep = stats[:, 0, _COL['count']] > 0
    out = {'road_id': np.asarray(road_ids)[keep]}
    for ci, b in enumerate(BANDS):
        s = stats[keep, ci]
        margin = Z * s[:, _COL['std']] / np.sqrt(s[:, _COL['count']])
        out[f'min_{b}'] = s[:, _COL['min']].astype(np.uint8)
        out[f'max_{b}'] = s[:, _COL['max']].astype(np.uint8)
        out[f'median_{b}'] = s[:, _COL['median']]
        out[f'mean_{b}'] = np.round(s[:, _COL['mean']], 2)
        out[f'std_{b}'] = np.round(s[:, _COL['std']], 2)
        out[f'margin_{b}'] = np.round(margin, 2)
    out['count'] = stats[keep, 0, _COL['count']].astype(np.int64)
    return pd.DataFrame(out)


def filter_roads(roads_stats_df: pd.DataFrame, BANDS: Sequence[int], COUNT_THRESHOLD=10, MAX_MOE=12.5) -> pd.DataFrame:
    """statistical_analysis.py:264-270: keep count > threshold and any margin below MAX_MOE."""
    ok = np.zeros(len(roads_stats_df), bool)
    for b in BANDS:
        ok |= (roads_stats_df[f'margin_{b}'] < MAX_MOE).to_numpy()
    keep = (roads_stats_df['count'] > COUNT_THRESHOLD).to_numpy() & ok
    return roads_stats_df[keep].drop(columns=[f'margin_{b}' for b in BANDS] + ['count'])


def ks_test_from_hists(hist_band: np.ndarray, road_type: Sequence, engine=None) -> pd.DataFrame:
    """statistical_analysis.py:441-461 without the pixel table: for every road, scipy.stats.kstest(pixels of the road,
    all pixels of the road's type) on one band.  hist_band (R, 256): per-road histogram of the band (what
    get_pixel_values would have returned for the road); road_type (R,): cover of each road.  The reference compares a
    road with the pooled pixels of its type, the road itself included.  Returns columns ks_D (round 3) and ks_p
    ('{:0.3e}'-rounded), D exact from the histograms, p from scipy's large-sample branch
    kstwo.sf(D, round(m n / (m + n))) (ks_2samp mode 'asymp', which 'auto' takes when a sample exceeds 10 000)."""
    from scipy.stats import distributions
    eng = engine or default_engine()
    h = np.ascontiguousarray(hist_band, np.uint32)
    types, code = np.unique(np.asarray(road_type), return_inverse=True)
    ref = np.zeros((len(types), 256), np.uint64)
    np.add.at(ref, code, h.astype(np.uint64))
    D, m = eng.ks_hist_host(h, ref, code.astype(np.int32))
    n = ref.sum(axis=1).astype(np.float64)[code]
    with np.errstate(invalid="ignore", divide="ignore"):
        en = np.round(m * n / (m + n))
    p = np.array([distributions.kstwo.sf(d, e) if e >= 1 and d == d else np.nan for d, e in zip(D, en)])
    return pd.DataFrame({"ks_D": np.round(D, 3), "ks_p": [float("{:0.3e}".format(v)) if v == v else np.nan for v in p]})


def cover_stats_from_accumulators(hist: np.ndarray, n_allzero: np.ndarray, road_type: Sequence, BANDS: Sequence[int] = (1, 2, 3),
                                  nodata_mode: str = "none", engine=None) -> pd.DataFrame:
    """statistical_analysis.py:296-316 from the per-road accumulators: the statistics of all pixels of each road type
    (cover) per band -- what get_df_stats_no_group returns on ``pixels_per_band[road_type == cover]`` -- with the
    reference's rounding (mean / std to 2 decimals and margin to 3 inside get_df_stats_no_group, then 1 decimal in the
    table, :314-316).  hist (R, C, 256), n_allzero (R,), road_type (R,)."""
    eng = engine or default_engine()
    h = np.asarray(hist)
    types, code = np.unique(np.asarray(road_type), return_inverse=True)
    pooled = np.zeros((len(types),) + h.shape[1:], np.uint64)
    np.add.at(pooled, code, h.astype(np.uint64))
    nz = np.zeros(len(types), np.uint64)
    np.add.at(nz, code, np.asarray(n_allzero, np.uint64))
    if pooled.max(initial=0) >= 2 ** 32:
        raise OverflowError("a pooled histogram bin exceeds 32 bits")
    st = eng.finalize_stats_host(pooled.astype(np.uint32), nz.astype(np.uint32), nodata_mode=nodata_mode, ddof=1)
    rows = {'cover': [], 'band': [], 'min': [], 'max': [], 'mean': [], 'median': [], 'std': [], 'margin': [], 'count': []}
    for ti, cover in enumerate(types):
        for ci, b in enumerate(BANDS):
            s_ = st[ti, b - 1]
            std2 = np.float64(s_[_COL['std']]).round(2)
            n = int(s_[_COL['count']])
            rows['cover'].append(cover)
            rows['band'].append(b)
            rows['min'].append(int(s_[_COL['min']]))
            rows['max'].append(int(s_[_COL['max']]))
            rows['mean'].append(np.float64(s_[_COL['mean']]).round(2))
            rows['median'].append(float(s_[_COL['median']]))
            rows['std'].append(std2)
            rows['count'].append(n)
            rows['margin'].append(np.round(Z * std2 / (n ** (1 / 2)), decimals=3))
    df = pd.DataFrame(rows)
    df['mean'] = df['mean'].round(1)
    df['std'] = df['std'].round(1)
    df['margin'] = df['margin'].round(1)
    return df


RATIO_NAMES = {'1/2': 'R/G', '1/3': 'R/B', '1/4': 'R/NIR', '2/3': 'G/B', '2/4': 'G/NIR', '3/4': 'B/NIR'}


def add_band_ratios(pixels_per_band: pd.DataFrame, BANDS: Sequence[int] = range(1, 5), engine=None) -> pd.DataFrame:
    """The derived per-pixel columns of scripts/statistical_analysis/statistical_analysis.py:279-293, added in place and
    returned: 'R/G', 'R/B', 'R/NIR', 'G/B', 'G/NIR', 'B/NIR' (band_a / band_b rounded to 3 decimals, NaN -> 0, inf -> 1) and
    'VgNIR-BI' (rounded to 5 decimals), one kernel over the uint8 band table (rs_band_ratios_host).  BANDS must be the
    consecutive bands 1..max(BANDS) of the reference's loops; like the reference, VgNIR-BI needs band2 and band4
    (KeyError otherwise)."""
    BANDS = list(BANDS)
    if BANDS != list(range(1, len(BANDS) + 1)) or not 2 <= len(BANDS) <= 4:
        raise ValueError("BANDS must be range(1, n+1) with n in 2..4")
    if len(BANDS) < 4:
        raise KeyError('band4')                      # statistical_analysis.py:289-291 reads band2 and band4 unconditionally
    values = np.stack([_uint8_column(pixels_per_band[f'band{b}'], f'band{b}') for b in BANDS], axis=1)
    out = (engine or default_engine()).band_ratios_host(values)
    k = 0
    for band in BANDS:
        for sec_band in range(band + 1, max(BANDS) + 1):
            pixels_per_band[RATIO_NAMES[f'{band}/{sec_band}']] = out[k]
            k += 1
    pixels_per_band['VgNIR-BI'] = out[k]
    return pixels_per_band
