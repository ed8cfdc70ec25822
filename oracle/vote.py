"""Oracle: per-road class vote, tags, confusion counts and class-balanced F1.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  scripts/road_segmentation/determine_class.py:122-190   determine_detected_class
  scripts/road_segmentation/final_metrics.py:22-89       get_metrics
  scripts/road_segmentation/final_metrics.py:91-105      get_tag
  scripts/road_segmentation/final_metrics.py:277-316     threshold sweep
Pinned: the table functions are checked against the reference's own code executed
with geopandas/plotly stubbed (tests/golden/vote_*.json, tests/golden/metrics.json).

The raster restatement (below table a14) is this project's
definition, not reference code: detections arrive as a class plane (0 none,
1 artificial, 2 natural) and a uint8 score plane; each road accumulates the joint
histogram h[k][s] of (class, score) over its pixels.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import pandas as pd

CLASSES = ["artificial", "natural"]
COVER = ["artificial", "natural", "undetermined", "undetected"]   # codes 0..3
THRESHOLDS = np.arange(0, 1.0, 0.05)                               # final_metrics.py:277


# ----------------------------------------------------------------------------
# vector form -- reference-shaped
# ----------------------------------------------------------------------------
def determine_detected_class(predictions: pd.DataFrame, roads: pd.DataFrame, threshold=0) -> pd.DataFrame:
    """predictions: OBJECTID, score, det_class_name, weighted_score, area_pred_in_label."""
    valid = predictions[predictions["score"] >= threshold]
    seen = set(valid["OBJECTID"].unique().tolist())
    rec = {"road_id": [], "cover_type": [], "nat_score": [], "art_score": [], "diff_score": []}
    for rid in roads["OBJECTID"].unique().tolist():
        rec["road_id"].append(rid)
        if rid not in seen:
            rec["cover_type"].append("undetected")
            rec["nat_score"].append(0)
            rec["art_score"].append(0)
            rec["diff_score"].append(0)
            continue
        sums = valid[valid["OBJECTID"] == rid].groupby("det_class_name")[["weighted_score", "area_pred_in_label"]].sum()
        idx = {}
        for k in ("natural", "artificial"):
            if k in sums.index and sums.loc[k, "weighted_score"] != 0:
                idx[k] = sums.loc[k, "weighted_score"] / sums.loc[k, "area_pred_in_label"]
            else:
                idx[k] = 0
        a, n = idx["artificial"], idx["natural"]
        if a == n:
            rec["cover_type"].append("undetermined")
            rec["diff_score"].append(0)
        else:
            rec["cover_type"].append("artificial" if a > n else "natural")
            rec["diff_score"].append(abs(a - n))
        rec["art_score"].append(round(a, 3))
        rec["nat_score"].append(round(n, 3))
    keep = ["OBJECTID"] + [c for c in ("CATEGORY", "gt_type") if c in roads.columns and "gt_type" in roads.columns]
    return pd.DataFrame(rec).merge(roads[keep], how="inner", left_on="road_id", right_on="OBJECTID")


def get_tag(cover_type: str, category: str) -> str:
    if cover_type in ("undetermined", "undetected"):
        return "FN"
    return "TP" if cover_type == category else "wrong class"


def get_metrics(comparison_df: pd.DataFrame, classes: Sequence[str] = CLASSES):
    rows = []
    for k in classes:
        is_k = comparison_df["CATEGORY"] == k
        tp = int(((comparison_df["tag"] == "TP") & is_k).sum())
        fp = int(((comparison_df["tag"] == "wrong class") & (comparison_df["cover_type"] == k)).sum())
        fn = int(((comparison_df["tag"] == "FN") & is_k).sum()) + int(((comparison_df["tag"] == "wrong class") & is_k).sum())
        if tp == 0:
            pk = rk = f1k = 0
        else:
            pk = tp / (tp + fp)
            rk = tp / (tp + fn)
            f1k = 2 * pk * rk / (pk + rk)
        rows.append({"cover_class": k, "TP": tp, "FP": fp, "FN": fn, "Pk": pk, "Rk": rk, "f1k": f1k, "count": int(is_k.sum())})
    by_class = pd.DataFrame(rows)
    return by_class, global_from_by_class(by_class)


def global_from_by_class(by_class: pd.DataFrame) -> pd.DataFrame:
    total = by_class["count"].sum()
    pw = (by_class["Pk"] * by_class["count"]).sum() / total
    rw = (by_class["Rk"] * by_class["count"]).sum() / total
    f1w = 0 if (pw == 0 and rw == 0) else 2 * pw * rw / (pw + rw)
    pb = by_class["Pk"].sum() / 2          # literal 2 (final_metrics.py:78-79)
    rb = by_class["Rk"].sum() / 2
    f1b = 0 if (pb == 0 and rb == 0) else 2 * pb * rb / (pb + rb)
    return pd.DataFrame({"Pw": [pw], "Rw": [rw], "f1w": [f1w], "Pb": [pb], "Rb": [rb], "f1b": [f1b]})


# ----------------------------------------------------------------------------
# raster form
# ----------------------------------------------------------------------------
def score_cutoffs(thresholds: Sequence[float] = THRESHOLDS) -> np.ndarray:
    """Smallest uint8 score s with s/255.0 >= thr, per threshold (256 if none)."""
    s = np.arange(256) / 255.0
    return np.array([int(np.argmax(s >= t)) if (s >= t).any() else 256 for t in thresholds], np.int32)


def raster_vote(joint_hist: np.ndarray, cutoff: int, rule: str = "count", min_area_frac: float = 0.0):
    """joint_hist (R, 3, 256) = h[k][s]; returns (cover_code (R,), art_score, nat_score, diff_score).

    rule 'count'  (north star): argmax of pixel counts n_k = sum_{s>=cutoff} h_k[s]
    rule 'score'  (reference-shaped, determine_class.py:149-176): compare the mean scores
                  S_k/(255 n_k), S_k = sum s*h_k[s]; exact through S_a*n_n vs S_n*n_a.
    Classes whose area fraction round(n_k/n_inside, 2) <= min_area_frac are ignored
    (per-class stand-in for determine_class.py:118, which filters per detection).
    """
    h = np.asarray(joint_hist, np.int64)
    R = h.shape[0]
    s = np.arange(256, dtype=np.int64)
    sel = s >= cutoff
    n = (h[:, :, sel]).sum(axis=2)                      # (R, 3)
    S = (h[:, :, sel] * s[sel]).sum(axis=2)
    n_inside = h.sum(axis=(1, 2))
    if min_area_frac > 0:
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = np.round(n / np.maximum(n_inside, 1)[:, None], 2)
        drop = frac <= min_area_frac
        n = np.where(drop, 0, n)
        S = np.where(drop, 0, S)
    na, nn = n[:, 1], n[:, 2]
    Sa, Sn = S[:, 1], S[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        ia = np.where((na > 0) & (Sa > 0), Sa / (255.0 * np.maximum(na, 1)), 0.0)
        inn = np.where((nn > 0) & (Sn > 0), Sn / (255.0 * np.maximum(nn, 1)), 0.0)
    if rule == "count":
        left, right = na, nn
    elif rule == "score":
        left, right = Sa * np.maximum(nn, 1) * (na > 0), Sn * np.maximum(na, 1) * (nn > 0)
    else:
        raise ValueError(rule)
    cover = np.full(R, 2, np.int32)
    cover[left > right] = 0
    cover[left < right] = 1
    cover[(na + nn) == 0] = 3
    diff = np.abs(ia - inn)
    return cover, ia, inn, diff


def confusion(cover: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """(2, 4) counts: rows gt class (0 artificial, 1 natural), columns COVER codes."""
    m = np.zeros((2, 4), np.int64)
    np.add.at(m, (np.asarray(gt), np.asarray(cover)), 1)
    return m


def metrics_from_confusion(m: np.ndarray) -> Dict[str, float]:
    """final_metrics.get_metrics from the confusion counts."""
    out: Dict[str, float] = {}
    P, Rr, cnt = [], [], []
    for k in (0, 1):
        tp = int(m[k, k])
        fp = int(m[1 - k, k])
        fn = int(m[k, 2] + m[k, 3] + m[k, 1 - k])
        if tp == 0:
            pk = rk = f1k = 0.0
        else:
            pk = tp / (tp + fp)
            rk = tp / (tp + fn)
            f1k = 2 * pk * rk / (pk + rk)
        out.update({f"TP_{k}": tp, f"FP_{k}": fp, f"FN_{k}": fn, f"P_{k}": pk, f"R_{k}": rk, f"f1_{k}": f1k})
        P.append(pk)
        Rr.append(rk)
        cnt.append(int(m[k].sum()))
    total = sum(cnt)
    pw = (P[0] * cnt[0] + P[1] * cnt[1]) / total if total else float("nan")
    rw = (Rr[0] * cnt[0] + Rr[1] * cnt[1]) / total if total else float("nan")
    out["Pw"], out["Rw"] = pw, rw
    out["f1w"] = 0.0 if (pw == 0 and rw == 0) else 2 * pw * rw / (pw + rw)
    pb, rb = (P[0] + P[1]) / 2, (Rr[0] + Rr[1]) / 2
    out["Pb"], out["Rb"] = pb, rb
    out["f1b"] = 0.0 if (pb == 0 and rb == 0) else 2 * pb * rb / (pb + rb)
    return out


def sweep(joint_hist: np.ndarray, gt: np.ndarray, thresholds: Sequence[float] = THRESHOLDS, rule: str = "count",
          min_area_frac: float = 0.0):
    """final_metrics.py:277-316 on raster accumulators: per threshold the confusion counts
    and metrics; best = max f1b, ties broken by larger Pb, first threshold wins otherwise."""
    cuts = score_cutoffs(thresholds)
    rows = []
    best = None
    for i, (t, c) in enumerate(zip(thresholds, cuts)):
        cover, *_ = raster_vote(joint_hist, int(c), rule, min_area_frac)
        m = confusion(cover, gt)
        met = metrics_from_confusion(m)
        met["threshold"] = float(t)
        met["confusion"] = m
        rows.append(met)
        if i == 0 or met["f1b"] > best[1] or (met["f1b"] == best[1] and met["Pb"] > best[2]):
            best = (i, met["f1b"], met["Pb"])
    return rows, best[0]


# ----------------------------------------------------------------------------
# instance-faithful raster form of get_weighted_scores (determine_class.py:97-120)
# ----------------------------------------------------------------------------
def weighted_scores_raster(inside_masks, instance_tiles, road_of_pair, pair_tile, road_ids, inst_score, inst_class_name,
                           min_area: float = 0.05) -> pd.DataFrame:
    """inside_masks[p] (H, W) bool mask of pair p; instance_tiles (T, H, W) uint16 instance ids (0 = no detection).
    area_label = pixels of the road (all its pairs); per (road, instance): area_pred_in_label =
    round(n_pixels / area_label, 2), weighted_score = area_pred_in_label * score, kept when area_pred_in_label > min_area
    -- the reference's overlay areas (determine_class.py:107-118) counted in pixels."""
    R = len(road_ids)
    area_label = np.zeros(R, np.int64)
    counts: Dict[Tuple[int, int], int] = {}
    for p, m in enumerate(inside_masks):
        r = int(road_of_pair[p])
        area_label[r] += int(m.sum())
        ids, c = np.unique(instance_tiles[int(pair_tile[p])][m], return_counts=True)
        for i, k in zip(ids.tolist(), c.tolist()):
            if i:
                counts[(r, i)] = counts.get((r, i), 0) + k
    rows = []
    for (r, i), k in sorted(counts.items()):
        frac = round(k / area_label[r], 2)
        if frac > min_area:
            rows.append({"OBJECTID": road_ids[r], "instance": i, "score": float(inst_score[i]), "det_class_name": inst_class_name[i],
                         "area_pred_in_label": frac, "weighted_score": frac * float(inst_score[i])})
    return pd.DataFrame(rows, columns=["OBJECTID", "instance", "score", "det_class_name", "area_pred_in_label", "weighted_score"])


def bin_accuracy(best_comparison_df: pd.DataFrame) -> list:
    """scripts/road_segmentation/final_metrics.py:541-571 (script body): calibration tables, one per (gt_type, parameter) --
    roads of the parameter's CATEGORY with threshold - 0.5 < score <= threshold (sic, :557), accuracy = share of them whose
    cover_type is the parameter's class; thresholds np.arange(0, 1.05, 0.05); empty bins are skipped."""
    params = {'artificial': ['art_score', 'artificial', 'artifical score'],
              'natural': ['nat_score', 'natural', 'natural score'],
              'artificial_diff': ['diff_score', 'artificial', 'score diff in artificial roads'],
              'naturall_diff': ['diff_score', 'natural', 'score diff in natural roads']}
    out = []
    for gt_type in best_comparison_df['gt_type'].unique():
        sub = best_comparison_df[best_comparison_df['gt_type'] == gt_type]
        for col, cls, label in params.values():
            acc, thr = [], []
            for threshold in np.arange(0, 1.05, 0.05):
                in_bin = sub[(sub[col] > threshold - 0.5) & (sub[col] <= threshold) & (sub['CATEGORY'] == cls)]
                if not in_bin.empty:
                    acc.append(in_bin[in_bin['cover_type'] == cls].shape[0] / in_bin.shape[0])
                    thr.append(threshold)
            df = pd.DataFrame({'threshold': thr, 'accuracy': acc})
            df.name = label + ' for ' + gt_type
            out.append(df)
    return out


# ------------------------------------------------------------------------------------------
# 'within' (determine_class.py:57: gpd.sjoin(roads, buffered_quarries, predicate='within'))
# ------------------------------------------------------------------------------------------
def _orient(p, q, r):
    """sign of the cross product (q - p) x (r - p), exact (rational arithmetic on the binary64 inputs)"""
    from fractions import Fraction as F
    v = (F(q[0]) - F(p[0])) * (F(r[1]) - F(p[1])) - (F(q[1]) - F(p[1])) * (F(r[0]) - F(p[0]))
    return (v > 0) - (v < 0)


def _edges(rings):
    for ring in rings:
        n = len(ring)
        for k in range(n):
            a, b = ring[k], ring[(k + 1) % n]
            if a[0] != b[0] or a[1] != b[1]:
                yield a, b


def point_in_rings(rings, p) -> int:
    """0 outside, 1 inside, 2 on the boundary; even-odd over all rings"""
    inside = 0
    for a, b in _edges(rings):
        o = _orient(a, b, p)
        if o == 0 and min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1]):
            return 2
        if (a[1] <= p[1]) != (b[1] <= p[1]):
            if (o > 0) == (b[1] > a[1]):
                inside ^= 1
    return inside


def polygon_within(a_rings, b_rings) -> bool:
    """Polygon a within polygon b (GEOS / DE-9IM 'within' for areal geometries: no point of a in the exterior of b; touching
    boundaries allowed): every vertex of a in or on b, no proper crossing between their edges, no vertex of b strictly
    inside a.  Exact predicates (GEOS evaluates orientation robustly too)."""
    if not a_rings or not b_rings:
        return False
    for ring in a_rings:
        for p in ring:
            if point_in_rings(b_rings, p) == 0:
                return False
    for p, q in _edges(a_rings):
        for c, d in _edges(b_rings):
            o1, o2, o3, o4 = _orient(p, q, c), _orient(p, q, d), _orient(c, d, p), _orient(c, d, q)
            if o1 * o2 < 0 and o3 * o4 < 0:
                return False
    for ring in b_rings:
        for p in ring:
            if point_in_rings(a_rings, p) == 1:
                return False
    return True
