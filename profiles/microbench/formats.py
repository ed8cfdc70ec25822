"""One tile format of the fused kernel at a chosen size, for ncu captures and quick comparisons; run on a B200:
    python profiles/microbench/formats.py --format u16lut --tiles 128 [--once]
formats: u8x3, u8x4, class_score, u16f64, u16f32 (RS_ZONAL_LUT=0), u16lut (RS_ZONAL_LUT=1, the default for the binary64 semantics)"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.engine import Engine, scale_params

ap = argparse.ArgumentParser()
ap.add_argument("--format", default="u16lut")
ap.add_argument("--tiles", type=int, default=128, help="tiles per side")
ap.add_argument("--once", action="store_true")
args = ap.parse_args()
eng = Engine(0)
g = synth.Grid(args.tiles, args.tiles)
rr = synth.ribbon_roads(g, args.tiles * args.tiles // 2)
dr, dp = eng.upload_roads(rr.roads), eng.upload_pairs(rr.pairs)
fmt = args.format
ch, dtype, kind, kw, bpp = {"u8x3": (3, "u8", 0, {}, 3), "u8x4": (4, "u8", 0, {}, 4), "class_score": (2, "u8", 2, {"hist_mode": "class_score"}, 2),
                            "u16f64": (4, "u16", 0, {"rescale": scale_params([0.0] * 4, [65535.0] * 4) + (False,)}, 8),
                            "u16f32": (4, "u16", 0, {"rescale": scale_params([0.0] * 4, [65535.0] * 4, True) + (True,)}, 8),
                            "u16lut": (4, "u16", 0, {"rescale": scale_params([0.0] * 4, [65535.0] * 4) + (False,)}, 8)}[fmt]
os.environ["RS_ZONAL_LUT"] = "1" if fmt == "u16lut" else "0"
t = eng.synth_tiles_dev(g.keys(), 256, 256, ch, dtype=dtype, kind=kind, gt=g.transforms())
out = eng.zonal_hist_dev(dr, t, dp, **kw)
if not args.once:
    for _ in range(2):
        eng.zonal_hist_dev(dr, t, dp, out=out, check=False, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.zonal_hist_dev(dr, t, dp, out=out, check=False, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    px = g.n_tiles * 65536
    print(json.dumps({"format": fmt, "tiles": g.n_tiles, "ms": ms, "Gpixel/s": px / ms / 1e6, "frac": px * bpp / ms / 1e6 / 6547.8}))
