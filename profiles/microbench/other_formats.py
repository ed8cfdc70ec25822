"""Full-scale timing of the other tile formats (class/score planes = BASELINE configs[1], 1- and 4-band uint8) and of the
vote + metrics kernels; run on a B200:  python profiles/microbench/other_formats.py"""
import sys, time, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from proj_roadsurf_b200 import synth
from proj_roadsurf_b200.engine import Engine, scale_params
from oracle import vote as ovote
eng = Engine(0)
g = synth.Grid(512, 512)
rr = synth.ribbon_roads(g, 131072)
dr, dp = eng.upload_roads(rr.roads), eng.upload_pairs(rr.pairs)
gt = g.transforms()
for name, ch, dtype, kind, kw, bpp in (("class_score_u8x2", 2, "u8", 2, {"hist_mode": "class_score"}, 2),
                                       ("bands_u8x4", 4, "u8", 0, {}, 4), ("bands_u8x1", 1, "u8", 0, {}, 1)):
    t2 = eng.synth_tiles_dev(g.keys(), 256, 256, ch, dtype=dtype, kind=kind, gt=gt)
    for _ in range(3):
        out2 = eng.zonal_hist_dev(dr, t2, dp, check=False, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.zonal_hist_dev(dr, t2, dp, out=out2, check=False, **kw)
    e1.record(); torch.cuda.synchronize(); eng.sync_status()
    ms = e0.elapsed_time(e1) / 10
    gpx = g.n_tiles * 65536 / (ms * 1e-3) / 1e9
    print(name, "ms", round(ms, 3), "Gpx/s", round(gpx, 1), "frac", round(gpx * bpp / 6547.8, 3))
    if name.startswith("class"):
        gtc = torch.from_numpy(rr.gt_class).cuda()
        cuts = ovote.score_cutoffs()
        for _ in range(3):
            r = eng.vote_metrics_dev(out2[0], gtc, cuts, rule="count", check=False)
        e0.record()
        for _ in range(10):
            r = eng.vote_metrics_dev(out2[0], gtc, cuts, rule="count", check=False)
        e1.record(); torch.cuda.synchronize()
        print("vote+metrics 20 thresholds ms", round(e0.elapsed_time(e1) / 10, 3), "f1b@0", float(r[3][0, 11]))
    del t2, out2
