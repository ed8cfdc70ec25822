"""ctypes front-end of the plain-C oracle (oracle/c/roadsurf_oracle.c).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); **parity unpinned** like the file it wraps.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "c", "roadsurf_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libroadsurf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    if force or not os.path.exists(_OUT) or (os.path.exists(_SRC) and os.path.getmtime(_OUT) < os.path.getmtime(_SRC)):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_zonal_accumulate.restype = ctypes.c_int
        _lib.orc_rasterize.restype = ctypes.c_int
        _lib.orc_pair_mask_full.restype = ctypes.c_int
        _lib.orc_geometry_window.restype = ctypes.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def flatten_rings(rings: Sequence[np.ndarray]):
    sizes = np.array([len(r) for r in rings], np.int32)
    xy = np.ascontiguousarray(np.concatenate([np.asarray(r, np.float64).reshape(-1, 2) for r in rings]) if len(rings) else np.zeros((0, 2)))
    return xy, sizes


def rasterize(rings, out_shape, transform=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0)) -> np.ndarray:
    H, W = out_shape
    xy, sizes = flatten_rings(rings)
    mask = np.zeros((H, W), np.uint8)
    t = np.asarray(transform, np.float64)
    rc = lib().orc_rasterize(_p(t), ctypes.c_int(len(sizes)), _p(sizes), _p(xy), ctypes.c_int(W), ctypes.c_int(H), _p(mask))
    if rc != 0:
        raise ValueError("rotated/degenerate transform")
    return mask


def pair_mask_full(transform, rings, W, H) -> np.ndarray:
    xy, sizes = flatten_rings(rings)
    mask = np.zeros((H, W), np.uint8)
    t = np.asarray(transform, np.float64)
    rc = lib().orc_pair_mask_full(_p(t), ctypes.c_int(len(sizes)), _p(sizes), _p(xy), ctypes.c_int(W), ctypes.c_int(H), _p(mask))
    if rc != 0:
        raise ValueError("rotated/degenerate transform")
    return mask


def geometry_window(transform, rings, W, H):
    xy, _ = flatten_rings(rings)
    t = np.asarray(transform, np.float64)
    win = np.zeros(4, np.int32)
    ok = lib().orc_geometry_window(_p(t), ctypes.c_int(len(xy)), _p(xy), ctypes.c_int(W), ctypes.c_int(H), _p(win))
    return tuple(int(v) for v in win) if ok else None


def zonal_accumulate(xy, ring_off, road_ring_off, road_pair_off, pair_tile, tiles, tile_gt,
                     scale_k=None, scale_off=None, rescale_f32=False, joint=False, threads: int = 1, want_min_zero: bool = False):
    """Per-road histograms over a road-major pair list.  tiles (T,H,W,C) uint8|uint16.
    want_min_zero: also return, per road, the sum over its (road, tile) calls of min over bands of the call's
    zero-valued in-mask pixels (what the per-call zero padding of fct_misc.py:95-111 leaves out).
    threads > 1 splits the road range over a thread pool (the C call releases the GIL);
    every road is owned by one thread, so results do not depend on the split."""
    xy = np.ascontiguousarray(xy, np.float64)
    ring_off = np.ascontiguousarray(ring_off, np.int32)
    road_ring_off = np.ascontiguousarray(road_ring_off, np.int32)
    road_pair_off = np.ascontiguousarray(road_pair_off, np.int32)
    pair_tile = np.ascontiguousarray(pair_tile, np.int32)
    tiles = np.ascontiguousarray(tiles)
    tile_gt = np.ascontiguousarray(tile_gt, np.float64)
    T, H, W, C = tiles.shape
    eb = tiles.dtype.itemsize
    assert tiles.dtype in (np.uint8, np.uint16)
    R = len(road_ring_off) - 1
    HC = 3 if joint else C
    hist = np.zeros((R, HC, 256), np.uint64)
    nzero = np.zeros(R, np.uint64)
    minz = np.zeros(R, np.uint64) if want_min_zero else None
    k = None if scale_k is None else np.ascontiguousarray(scale_k, np.float64)
    o = None if scale_off is None else np.ascontiguousarray(scale_off, np.float64)
    if eb == 2:
        assert k is not None and o is not None

    def run(lo, hi):
        return lib().orc_zonal_accumulate(
            _p(xy), _p(ring_off), _p(road_ring_off), _p(road_pair_off), _p(pair_tile), _p(tiles), _p(tile_gt),
            ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(C), ctypes.c_int(eb), _p(k), _p(o),
            ctypes.c_int(int(rescale_f32)), ctypes.c_int(int(joint)), ctypes.c_int(lo), ctypes.c_int(hi),
            _p(hist), _p(nzero), _p(minz))

    if threads <= 1 or R < 2 * threads:
        rcs = [run(0, R)]
    else:
        # split by pair count so threads get similar work
        cuts = np.searchsorted(road_pair_off, np.linspace(0, road_pair_off[-1], threads * 4 + 1)[1:-1])
        bounds = sorted(set([0, R] + [int(c) for c in cuts]))
        with ThreadPoolExecutor(threads) as ex:
            rcs = list(ex.map(lambda ab: run(*ab), zip(bounds[:-1], bounds[1:])))
    if any(rc != 0 for rc in rcs):
        raise ValueError("rotated/degenerate transform")
    return (hist, nzero, minz) if want_min_zero else (hist, nzero)
