"""Does the host link carry bulk DMA next to the in-place sector reads?  The in-place transport of bench.py's e2e leg is bound
by the number of read requests the SMs can keep in flight over PCIe (~24 GB/s of a ~54 GB/s link); this script runs the in-place
kernel on one part of the tiles and the streamed copies on the rest AT THE SAME TIME (two contexts, two host threads) and prints
the rate for several splits.  python profiles/microbench/hybrid_e2e.py [tile_rows]"""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proj_roadsurf_b200 import synth                              # noqa: E402
from proj_roadsurf_b200.engine import Engine                      # noqa: E402
from proj_roadsurf_b200.geometry import TileBatch                 # noqa: E402

H = W = 256
C = 3


def sub(roads, pairs, lo, hi):
    p = pairs.restrict_tiles(lo, hi)
    idx = np.nonzero(np.diff(p.road_pair_off) > 0)[0]
    return roads.subset(idx), p.take_roads(idx)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    tx = 512
    grid = synth.Grid(tx, rows)
    rr = synth.ribbon_roads(grid, 256 * rows)
    n = grid.n_tiles
    gt = grid.transforms()
    e_a, e_b = Engine(0), Engine(0)
    dev = e_a.synth_tiles_dev(grid.keys(np.arange(n)), H, W, C, kind=0, gt=gt)
    host = torch.empty((n, H, W, C), dtype=torch.uint8, pin_memory=True)
    host.copy_(dev.pixels)
    torch.cuda.synchronize()
    del dev
    hp = host.numpy()

    def run(f_stream):
        n_s = int(round(rows * f_stream)) * tx                   # the LAST n_s tiles are streamed, the first n - n_s read in place
        jobs = []
        if n - n_s > 0:
            ra, pa = sub(rr.roads, rr.pairs, 0, n - n_s)
            jobs.append((e_a, ra, TileBatch(hp[:n - n_s], gt[:n - n_s], H, W, C), pa, dict(mapped=True)))
        if n_s > 0:
            rb, pb = sub(rr.roads, rr.pairs, n - n_s, n)
            jobs.append((e_b, rb, TileBatch(hp[n - n_s:], gt[n - n_s:], H, W, C), pb, dict(tiles_per_chunk=8192)))

        def work(j, out, k):
            t0 = time.perf_counter()
            j[0].zonal_stats_host(j[1], j[2], j[3], **j[4])
            out[k] = time.perf_counter() - t0

        best, parts = None, None
        for _ in range(3):
            out = [0.0] * len(jobs)
            th = [threading.Thread(target=work, args=(j, out, k)) for k, j in enumerate(jobs)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, parts = dt, list(out)
        print(f"streamed fraction {f_stream:.2f}: {n * H * W / best / 1e9:7.1f} Gpixel/s  ({best * 1e3:7.1f} ms; parts "
              + ", ".join(f"{p * 1e3:.1f}" for p in parts) + " ms)", flush=True)

    for f in (0.0, 1.0, 0.25, 0.3125, 0.375, 0.4375, 0.5):
        run(f)


if __name__ == "__main__":
    main()
