"""Mirror of the metric helpers of scripts/road_segmentation/final_metrics.py (:22-105, :277-316).

``get_tag`` and ``get_metrics`` keep the reference's table interface; the counting and the P / R / F1
arithmetic run on the GPU (rs_confusion_metrics_host).  ``threshold_sweep`` is the batched form of the
20-threshold loop (:277-316) on raster accumulators.
"""
from __future__ import annotations

import sys
from typing import Sequence

import numpy as np
import pandas as pd

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("roadsurf_b200")

from .._native import METRIC_COLS
from ..engine import default_engine
from . import determine_class

_M = {k: i for i, k in enumerate(METRIC_COLS)}
COVER_CODE = {"artificial": 0, "natural": 1, "undetermined": 2, "undetected": 3}


def get_tag(row):
    'final_metrics.py:91-105: FN for undetermined / undetected, TP when the classes agree, else wrong class.'
    det_class = row.cover_type
    gt_class = row.CATEGORY
    if det_class == 'undetermined' or det_class == 'undetected':
        return 'FN'
    elif det_class == gt_class:
        return 'TP'
    elif det_class != gt_class:
        return 'wrong class'
    else:
        logger.error(f'Unexpected configuration: prediction class is {det_class} and ground truth class is {gt_class}.')
        sys.exit(1)


def tags_from_codes(cover: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """vectorised get_tag on cover / ground-truth codes"""
    out = np.where(cover >= 2, 'FN', np.where(cover == gt, 'TP', 'wrong class'))
    return out.astype(object)


def get_metrics(comparison_df, CLASSES):
    """Per-class TP / FP / FN / Pk / Rk / f1k / count and the global Pw, Rw, f1w, Pb, Rb, f1b
    (balanced metrics divide by the literal 2 of final_metrics.py:78-79).  The reference reads the ``tag``
    column; tags are a function of (cover_type, CATEGORY), which is what is counted here."""
    if list(CLASSES) != ['artificial', 'natural']:
        raise ValueError("the GPU metrics path is built for CLASSES == ['artificial', 'natural']")
    cover = comparison_df['cover_type'].map(COVER_CODE).fillna(-1).to_numpy().astype(np.int8)
    gt = comparison_df['CATEGORY'].map(determine_class.CLASS_CODE).fillna(-1).to_numpy().astype(np.int8)
    if 'tag' in comparison_df.columns:              # honour a tag column that disagrees with get_tag (it never does)
        expect = tags_from_codes(cover, gt)
        if not np.array_equal(expect[(gt >= 0) & (cover >= 0)], comparison_df['tag'].to_numpy()[(gt >= 0) & (cover >= 0)]):
            raise ValueError("tag column is not get_tag(cover_type, CATEGORY)")
    conf, met = default_engine().confusion_metrics_host(cover[None, :], gt)
    return metrics_frames(conf[0], met[0], CLASSES)


def metrics_frames(conf: np.ndarray, met: np.ndarray, CLASSES=('artificial', 'natural')):
    """(2, 4) confusion counts + 12 metrics -> the reference's two DataFrames."""
    rows = {'cover_class': [], 'TP': [], 'FP': [], 'FN': [], 'Pk': [], 'Rk': [], 'f1k': [], 'count': []}
    for k, name in enumerate(CLASSES):
        tp = int(conf[k, k])
        rows['cover_class'].append(name)
        rows['TP'].append(tp)
        rows['FP'].append(int(conf[1 - k, k]))
        rows['FN'].append(int(conf[k, 2] + conf[k, 3] + conf[k, 1 - k]))
        # the reference stores integer 0 when TP == 0, floats otherwise
        rows['Pk'].append(0 if tp == 0 else float(met[_M[f'P{k}']]))
        rows['Rk'].append(0 if tp == 0 else float(met[_M[f'R{k}']]))
        rows['f1k'].append(0 if tp == 0 else float(met[_M[f'F{k}']]))
        rows['count'].append(int(conf[k].sum()))
    metrics_df = pd.DataFrame(rows)
    pw, rw, pb, rb = (float(met[_M[k]]) for k in ('Pw', 'Rw', 'Pb', 'Rb'))
    global_metrics_df = pd.DataFrame({'Pw': [pw], 'Rw': [rw], 'f1w': [0 if (pw == 0 and rw == 0) else float(met[_M['f1w']])],
                                      'Pb': [pb], 'Rb': [rb], 'f1b': [0 if (pb == 0 and rb == 0) else float(met[_M['f1b']])]})
    return metrics_df, global_metrics_df


CLASSES = ['artificial', 'natural']          # final_metrics.py:186 (module-level constant of the reference script)


def show_metrics(metrics_by_class, global_metrics):
    """final_metrics.py:107-123: log the by-class precision / recall and the balanced f1, precision and recall."""
    for metric in metrics_by_class.itertuples():
        logger.info(f"The {metric.cover_class} roads have a precision of {round(metric.Pk, 2)}"
                    + f" and a recall of {round(metric.Rk, 2)}.")
    logger.info(f"The final f1-score is {round(global_metrics.f1b[0], 2)}"
                + f" with a precision of {round(global_metrics.Pb[0], 2)} and a recall of"
                + f" {round(global_metrics.Rb[0], 2)}.")


def from_preds_to_metrics(predictions, ground_truth, by_class_metrics, global_metrics, dataset_name, threshold=0, show=False):
    """final_metrics.py:126-159: detected class of every road at ``threshold`` (determine_detected_class), tags, metrics,
    appended to the running by-class and global tables with their ``dataset`` and ``threshold`` columns.
    Returns (comparison_df, by_class_metrics, global_metrics)."""
    comparison_df = determine_class.determine_detected_class(predictions, ground_truth, threshold)
    cover = comparison_df['cover_type'].map(COVER_CODE).fillna(-1).to_numpy().astype(np.int8)
    gt = comparison_df['CATEGORY'].map(determine_class.CLASS_CODE).fillna(-1).to_numpy().astype(np.int8)
    comparison_df['tag'] = tags_from_codes(cover, gt)                    # get_tag row by row in the reference
    dst_metrics_by_class, dst_global_metrics = get_metrics(comparison_df, CLASSES)
    if show:
        show_metrics(dst_metrics_by_class, dst_global_metrics)
    dst_metrics_by_class['dataset'] = dataset_name
    dst_metrics_by_class['threshold'] = threshold
    by_class_metrics = pd.concat([by_class_metrics, dst_metrics_by_class], ignore_index=True)
    dst_global_metrics['dataset'] = dataset_name
    dst_global_metrics['threshold'] = threshold
    global_metrics = pd.concat([global_metrics, dst_global_metrics], ignore_index=True)
    return comparison_df, by_class_metrics, global_metrics


def best_threshold(f1b: Sequence[float], Pb: Sequence[float], thresholds: Sequence[float]):
    """final_metrics.py:295-312: maximise f1b, a tie goes to the larger Pb, otherwise the earlier threshold;
    the first threshold is reported as 0, later ones as round(threshold, 2)."""
    best, max_f1, max_p = 0, f1b[0], Pb[0]
    for i in range(1, len(thresholds)):
        if f1b[i] > max_f1 or (f1b[i] == max_f1 and Pb[i] > max_p):
            best, max_f1, max_p = i, f1b[i], Pb[i]
    return best, (0 if best == 0 else round(float(thresholds[best]), 2))


def threshold_sweep(joint_hist: np.ndarray, gt_class: np.ndarray, thresholds=None, rule: str = "count",
                    min_area_frac: float = 0.0, *, return_scores: bool = False):
    """The 20-threshold loop of final_metrics.py:277-316 on raster accumulators, one GPU launch.
    Returns (all_metrics_by_class, all_global_metrics, best_index, best_threshold, cover[, scores (T, R, 3) = artificial
    index, natural index, |difference| when return_scores])."""
    thresholds = np.arange(0, 1., 0.05) if thresholds is None else np.asarray(thresholds, float)
    cover, scores, conf, met = determine_class.raster_vote(joint_hist, gt_class, thresholds, rule, min_area_frac)
    by_class, glob = [], []
    for i, thr in enumerate(thresholds):
        a, b = metrics_frames(conf[i], met[i])
        a['threshold'] = thr
        b['threshold'] = thr
        by_class.append(a)
        glob.append(b)
    all_by_class = pd.concat(by_class, ignore_index=True)
    all_global = pd.concat(glob, ignore_index=True)
    bi, bt = best_threshold(all_global['f1b'].tolist(), all_global['Pb'].tolist(), thresholds)
    if return_scores:
        return all_by_class, all_global, bi, bt, cover, scores
    return all_by_class, all_global, bi, bt, cover


def diff_score_sweep(comparison_df, thresholds=None, CLASSES=('artificial', 'natural')):
    """final_metrics.py:429-478: for every threshold, roads whose diff_score is below it become 'undetermined', tags
    and metrics are recomputed (all thresholds in one GPU call); the best threshold maximises f1b (strictly greater
    wins, the first threshold is reported as 0).  Returns (metrics_by_class, global_metrics, best_threshold, best_results)
    where best_results is the comparison table at the best threshold."""
    thresholds = np.arange(0, 1., 0.05) if thresholds is None else np.asarray(thresholds, float)
    cover = comparison_df['cover_type'].map(COVER_CODE).fillna(-1).to_numpy().astype(np.int8)
    gt = comparison_df['CATEGORY'].map(determine_class.CLASS_CODE).fillna(-1).to_numpy().astype(np.int8)
    diff = comparison_df['diff_score'].to_numpy(float)
    cover_t = np.where(diff[None, :] < thresholds[:, None], np.int8(2), cover[None, :]).astype(np.int8)
    conf, met = default_engine().confusion_metrics_host(cover_t, gt)
    by_class, glob = [], []
    best, max_f1 = 0, None
    for i, thr in enumerate(thresholds):
        a, b = metrics_frames(conf[i], met[i], CLASSES)
        a['threshold'] = thr
        b['threshold'] = thr
        by_class.append(a)
        glob.append(b)
        f1 = b['f1b'][0]
        if i == 0 or f1 > max_f1:
            best, max_f1 = i, f1
    best_results = comparison_df.copy()
    best_results['cover_type'] = determine_class.COVER_NAMES[cover_t[best]]
    best_results['tag'] = tags_from_codes(cover_t[best], gt)
    return (pd.concat(by_class, ignore_index=True), pd.concat(glob, ignore_index=True),
            0 if best == 0 else round(float(thresholds[best]), 2), best_results)


BIN_ACCURACY_PARAM = {'artificial': ['art_score', 'artificial', 'artifical score'],
                      'natural': ['nat_score', 'natural', 'natural score'],
                      'artificial_diff': ['diff_score', 'artificial', 'score diff in artificial roads'],
                      'naturall_diff': ['diff_score', 'natural', 'score diff in natural roads']}


def bin_accuracy(best_comparison_df, thresholds_bins=None):
    """final_metrics.py:541-571: the calibration tables.  For every gt_type (in order of first appearance) and every entry of
    BIN_ACCURACY_PARAM, the share of roads with cover_type == the entry's class among the roads of that CATEGORY whose
    score lies in (threshold - 0.5, threshold] -- the 0.5 is the reference's (:557) -- for the thresholds
    np.arange(0, 1.05, 0.05) with a non-empty bin.  Returns the list of DataFrames (threshold, accuracy), each with the
    reference's ``name``; all counts come from one rs_bin_counts_host call."""
    thresholds_bins = np.arange(0, 1.05, 0.05) if thresholds_bins is None else np.asarray(thresholds_bins, float)
    df = best_comparison_df
    gt_types = list(pd.unique(df['gt_type']))
    group = df['gt_type'].map({g: i for i, g in enumerate(gt_types)}).to_numpy().astype(np.int32)
    params = list(BIN_ACCURACY_PARAM.values())
    values = np.stack([df[p[0]].to_numpy(float) for p in params])
    sel = np.stack([(df['CATEGORY'] == p[1]).to_numpy() for p in params]).astype(np.int8)
    hit = np.stack([(df['cover_type'] == p[1]).to_numpy() for p in params]).astype(np.int8)
    counts = default_engine().bin_counts_host(values, sel, hit, group, max(len(gt_types), 1), thresholds_bins - 0.5, thresholds_bins)
    tables = []
    for gi, gt_type in enumerate(gt_types):
        for ki, p in enumerate(params):
            keep = counts[gi, ki, :, 0] > 0
            t = pd.DataFrame({'threshold': [float(x) for x in thresholds_bins[keep]],
                              'accuracy': [int(h) / int(c) for c, h in counts[gi, ki][keep]]})
            t.name = p[2] + ' for ' + gt_type
            tables.append(t)
    return tables
