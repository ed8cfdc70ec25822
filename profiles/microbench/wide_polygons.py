"""BASELINE configs[4] (1024 px tiles, wide polygons with holes and 1 k - 10 k vertices): the wide-window kernel (rs_wide.cu)
against the fused kernel (RS_ZONAL_WIDE=0) on the same inputs; run on a B200:
    python profiles/microbench/wide_polygons.py [--grid 64 --polys 1024 --once]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--polys", type=int, default=1024)
ap.add_argument("--once", action="store_true", help="one launch of the wide kernel only (for ncu)")
args = ap.parse_args()
eng = Engine(0)
g = synth.Grid(args.grid, args.grid, size=1024)
wp = synth.wide_polygons(g, args.polys)
t = eng.synth_tiles_dev(g.keys(), 1024, 1024, 3, kind=0, gt=g.transforms())
dr, dp = eng.upload_roads(wp.roads), eng.upload_pairs(wp.pairs)
res = {}
for name, flag in (("wide", "1"), ("fused", "0")):
    if args.once and name == "fused":
        break
    os.environ["RS_ZONAL_WIDE"] = flag
    l0 = eng.launch_count
    out = eng.zonal_hist_dev(dr, t, dp)
    launches = eng.launch_count - l0
    if args.once:
        break
    for _ in range(2):
        eng.zonal_hist_dev(dr, t, dp, out=out, check=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.zonal_hist_dev(dr, t, dp, out=out, check=False)
    e1.record()
    torch.cuda.synchronize()
    eng.sync_status()
    ms = e0.elapsed_time(e1) / 5
    px = g.n_tiles * 1024 * 1024
    res[name] = {"ms": ms, "Gpixel/s": px / ms / 1e6, "frac_3B_6547.8": px * 3 / ms / 1e6 / 6547.8, "launches": launches,
                 "hist": out[0].cpu().numpy(), "nz": out[1].cpu().numpy()}
if not args.once:
    same = bool(np.array_equal(res["wide"]["hist"], res["fused"]["hist"]) and np.array_equal(res["wide"]["nz"], res["fused"]["nz"]))
    cov = float(res["wide"]["hist"][:, 0].astype(np.int64).sum()) / (g.n_tiles * 1024.0 * 1024.0)
    nv = np.diff(wp.roads.ring_off[wp.roads.road_ring_off])
    print(json.dumps({"tiles": g.n_tiles, "polygons": args.polys, "pairs": wp.pairs.n_pairs, "covered_fraction": cov,
                      "mean_vertices": float(nv.mean()), "max_vertices": int(nv.max()), "identical_histograms": same,
                      **{k: {kk: vv for kk, vv in v.items() if kk not in ("hist", "nz")} for k, v in res.items()}}))
