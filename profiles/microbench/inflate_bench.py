#!/usr/bin/env python
"""Device inflate alone: N low-entropy tiles as 8-row deflate strips through rs_decode_segments_host and through
rs_zonal_stats_compressed_host; run under `ncu --metrics gpu__time_duration.sum` for the kernel times.
  python profiles/microbench/inflate_bench.py [tiles] [zlib level]"""
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from proj_roadsurf_b200 import synth  # noqa: E402
from proj_roadsurf_b200.engine import Engine  # noqa: E402

n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(0)
g = synth.Grid(64, n_tiles // 64)
t = eng.synth_tiles_dev(g.keys(), 256, 256, 3, kind=1)
host = t.pixels.cpu().numpy()
flat = host.reshape(n_tiles * 32, -1)
with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
    comp_l = list(ex.map(lambda i: zlib.compress(flat[i].tobytes(), level), range(len(flat))))
comp_off = np.zeros(len(comp_l) + 1, np.int64)
comp_off[1:] = np.cumsum([len(c) for c in comp_l])
raw_off = np.arange(len(comp_l) + 1, dtype=np.int64) * flat.shape[1]
comp = np.frombuffer(b"".join(comp_l), np.uint8)
for warp in ("lut", "warp", "bits"):
    os.environ["RS_INFLATE"] = warp
    eng.decode_segments_host(comp, comp_off, 8, raw_off)
    t0 = time.perf_counter()
    out = eng.decode_segments_host(comp, comp_off, 8, raw_off)
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(out, host.reshape(-1)))
    print(f"RS_INFLATE={warp}: {n_tiles} tiles, {len(comp_l)} strips, ratio {host.nbytes / comp.nbytes:.2f}: decode_segments_host {dt * 1e3:.1f} ms "
          f"(upload + decode + download), equal={ok}")
