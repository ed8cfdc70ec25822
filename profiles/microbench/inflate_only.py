#!/usr/bin/env python
"""One decoder (RS_INFLATE from the environment), kernel time by CUDA events around rs_decode_segments_dev.
  python profiles/microbench/inflate_only.py [tiles] [zlib level]"""
import ctypes as C
import os
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from proj_roadsurf_b200 import synth  # noqa: E402
from proj_roadsurf_b200 import _native as N  # noqa: E402
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
comp = torch.from_numpy(np.frombuffer(b"".join(comp_l), np.uint8).copy()).cuda()
co, ro = torch.from_numpy(comp_off).cuda(), torch.from_numpy(raw_off).cuda()
raw = torch.zeros(host.size, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def run():
    rc = eng.lib.rs_decode_segments_dev(eng._ctx, comp.data_ptr(), co.data_ptr(), len(comp_l), 8, raw.data_ptr(), ro.data_ptr(), C.c_void_p(st))
    assert rc == 0, rc


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
ok = bool(np.array_equal(raw.cpu().numpy(), host.reshape(-1)))
print(f"RS_INFLATE={os.environ.get('RS_INFLATE', 'default')} lib={os.path.basename(os.environ.get('ROADSURF_B200_LIB', 'default'))}: {n_tiles} tiles, "
      f"{len(comp_l)} strips, level {level}, ratio {host.nbytes / comp.numel():.2f}: {ms:.2f} ms = {host.nbytes / ms / 1e6:.1f} GB/s out, equal={ok}")
