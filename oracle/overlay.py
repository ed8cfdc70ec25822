"""Test oracle: area of the intersection of two polygons, the quantity behind
scripts/road_segmentation/determine_class.py:107-118 get_weighted_scores
    all_intersections = gpd.overlay(ground_truth, predictions, how='intersection', keep_geom_type=True)
    joined_area = all_intersections.area ;  area_pred_in_label = round(joined_area / area_label, 2)

geopandas / shapely (GEOS) are not installable here, so this is a restatement: **parity unpinned** against
GEOS; it is exact for the polygons as given (rational arithmetic), GEOS' overlay nodes the edges in floating point and
agrees to ~1e-12 relative.  The algorithm is deliberately NOT the one of the CUDA kernel (which integrates x dy along the
clipped boundaries): here the plane is cut into horizontal strips at every vertex and every edge crossing; inside a strip
the edges do not cross, so the width of (A and B) is linear in y and the strip area is a trapezoid.
Test infrastructure only.
"""
from __future__ import annotations

from fractions import Fraction as F
from typing import List, Sequence

import numpy as np


def _edges(rings: Sequence[np.ndarray]):
    out = []
    for ring in rings:
        n = len(ring)
        for k in range(n):
            a, b = ring[k], ring[(k + 1) % n]
            ax, ay, bx, by = F(float(a[0])), F(float(a[1])), F(float(b[0])), F(float(b[1]))
            if ay != by:                               # horizontal edges bound no area in a horizontal strip decomposition
                out.append((ax, ay, bx, by))
    return out


def polygon_area(rings: Sequence[np.ndarray]) -> float:
    """even-odd area of a ring set (|shoelace| of every ring with the sign of its nesting parity), through the strips"""
    return intersection_area(rings, None)


def _x_at(e, y):
    ax, ay, bx, by = e
    return ax + (bx - ax) * (y - ay) / (by - ay)


def _inside_intervals(edges, y_mid, y):
    """sorted crossing abscissae at height y of the edges that span the open strip around y_mid"""
    span = sorted((e for e in edges if min(e[1], e[3]) < y_mid < max(e[1], e[3])), key=lambda e: _x_at(e, y_mid))
    xs = [_x_at(e, y) for e in span]                   # the order inside the strip, evaluated at the strip's end
    return [(xs[i], xs[i + 1]) for i in range(0, len(xs) - 1, 2)]


def _overlap(ia, ib):
    tot = F(0)
    for a0, a1 in ia:
        for b0, b1 in ib:
            lo, hi = max(a0, b0), min(a1, b1)
            if hi > lo:
                tot += hi - lo
    return tot


def intersection_area(a_rings: Sequence[np.ndarray], b_rings) -> float:
    ea = _edges(a_rings)
    eb = _edges(b_rings) if b_rings is not None else None
    ys = {e[1] for e in ea} | {e[3] for e in ea}
    allE = list(ea)
    if eb is not None:
        ys |= {e[1] for e in eb} | {e[3] for e in eb}
        allE += eb
    # crossings between any two edges (including edges of the same polygon: self-touching rings)
    for i in range(len(allE)):
        ax, ay, bx, by = allE[i]
        for j in range(i + 1, len(allE)):
            cx, cy, dx, dy = allE[j]
            den = (bx - ax) * (dy - cy) - (by - ay) * (dx - cx)
            if den == 0:
                continue
            t = ((cx - ax) * (dy - cy) - (cy - ay) * (dx - cx)) / den
            u = ((cx - ax) * (by - ay) - (cy - ay) * (bx - ax)) / den
            if 0 < t < 1 and 0 < u < 1:
                ys.add(ay + t * (by - ay))
    ys = sorted(ys)
    area = F(0)
    for y0, y1 in zip(ys[:-1], ys[1:]):
        ym = (y0 + y1) / 2
        w = []
        for y in (y0, y1):
            ia = _inside_intervals(ea, ym, y)
            if eb is None:
                w.append(sum((b - a for a, b in ia), F(0)))
            else:
                w.append(_overlap(ia, _inside_intervals(eb, ym, y)))
        area += (w[0] + w[1]) / 2 * (y1 - y0)
    return float(area)


def get_weighted_scores(labels: List[Sequence[np.ndarray]], preds: List[Sequence[np.ndarray]], score: Sequence[float]):
    """rows (label index, prediction index, joined_area, area_pred_in_label, weighted_score) of determine_class.py:107-118:
    every (label, prediction) pair with a positive intersection area, in (label, prediction) order, filtered to
    area_pred_in_label > 0.05 with area_pred_in_label = round(joined_area / area_label, 2)"""
    def box(rings):
        v = np.concatenate([np.asarray(r, float) for r in rings])
        return v[:, 0].min(), v[:, 1].min(), v[:, 0].max(), v[:, 1].max()
    pb = [box(b) for b in preds]
    rows = []
    for i, a in enumerate(labels):
        area_label = polygon_area(a)
        ab = box(a)
        for j, b in enumerate(preds):
            if ab[0] > pb[j][2] or ab[2] < pb[j][0] or ab[1] > pb[j][3] or ab[3] < pb[j][1]:
                continue                                   # disjoint boxes: no intersection
            joined = intersection_area(a, b)
            if joined <= 0.0:
                continue
            frac = round(joined / area_label, 2)
            if frac > 0.05:
                rows.append((i, j, joined, frac, frac * float(score[j])))
    return rows


def polygon_intersects_rect(rings: Sequence[np.ndarray], rect) -> bool:
    """Exact (rational arithmetic) 'intersects' of a polygon given by its rings (even-odd) and the CLOSED rectangle
    rect = (xmin, ymin, xmax, ymax): the predicate of gpd.sjoin(tiles, roads) (statistical_analysis.py:170-171); touching counts.
    True iff an edge of the polygon meets the rectangle, or the rectangle lies inside the polygon."""
    x0, y0, x1, y1 = (F(float(v)) for v in rect)
    edges = []
    for ring in rings:
        n = len(ring)
        for k in range(n):
            a, b = ring[k], ring[(k + 1) % n]
            edges.append((F(float(a[0])), F(float(a[1])), F(float(b[0])), F(float(b[1]))))
    for ax, ay, bx, by in edges:
        # clip the segment a + t (b - a), t in [0, 1], against the four half-planes (Liang-Barsky, exact)
        t0, t1 = F(0), F(1)
        ok = True
        for p, q in ((-(bx - ax), ax - x0), (bx - ax, x1 - ax), (-(by - ay), ay - y0), (by - ay, y1 - ay)):
            if p == 0:
                if q < 0:
                    ok = False
                    break
            else:
                r = q / p
                if p < 0:
                    t0 = max(t0, r)
                else:
                    t1 = min(t1, r)
                if t0 > t1:
                    ok = False
                    break
        if ok:
            return True
    # no edge meets the rectangle: it is wholly inside or wholly outside the polygon; test its lower-left corner (even-odd)
    inside = False
    for ax, ay, bx, by in edges:
        if (ay <= y0) != (by <= y0):
            xc = ax + (bx - ax) * (y0 - ay) / (by - ay)
            if xc > x0:
                inside = not inside
    return inside
