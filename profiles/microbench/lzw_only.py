#!/usr/bin/env python
"""The TIFF LZW decoder alone: 64 distinct 8-row strips of low-entropy 256 x 256 x 3 tiles encoded by tests/tiff_util.lzw_encode
(libtiff's convention), replicated to N strips; kernel time by CUDA events around rs_decode_segments_dev.
  python profiles/microbench/lzw_only.py [strips]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tiff_util import lzw_encode  # noqa: E402
from proj_roadsurf_b200.engine import Engine  # noqa: E402

n_strips = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
rng = np.random.default_rng(7)
base = [np.clip(rng.normal(110, 6, 6144), 0, 255).astype(np.uint8).tobytes() for _ in range(64)]
enc = [lzw_encode(b) for b in base]
idx = np.arange(n_strips) % 64
comp_l = [enc[i] for i in idx]
comp_off = np.zeros(n_strips + 1, np.int64)
comp_off[1:] = np.cumsum([len(c) for c in comp_l])
raw_off = np.arange(n_strips + 1, dtype=np.int64) * 6144
want = np.frombuffer(b"".join(base[i] for i in idx), np.uint8)
eng = Engine(0)
comp = torch.from_numpy(np.frombuffer(b"".join(comp_l), np.uint8).copy()).cuda()
co, ro = torch.from_numpy(comp_off).cuda(), torch.from_numpy(raw_off).cuda()
raw = torch.zeros(want.size, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def run():
    rc = eng.lib.rs_decode_segments_dev(eng._ctx, comp.data_ptr(), co.data_ptr(), n_strips, 5, raw.data_ptr(), ro.data_ptr(), C.c_void_p(st))
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
ok = bool(np.array_equal(raw.cpu().numpy(), want))
print(f"LZW: {n_strips} strips of 6144 bytes, ratio {want.size / comp.numel():.2f}: {ms:.2f} ms = {want.size / ms / 1e6:.1f} GB/s out, equal={ok}")
