#!/usr/bin/env python
"""Headline benchmark: Gpixel/s of fused rasterize + per-road zonal statistics (BASELINE.json).

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the CPU path (oracle port, all host cores)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU

Workload (config.workload): the canton-scale case of BASELINE.json configs[3] -- ~1 M buffered road
polygons over ~2 M 256x256 3-band uint8 zoom-18 tiles on 8 GPUs -- sharded by tile: every GPU owns a
512 x 512 tile band (262 144 tiles, 17.2 Gpx, 51.5 GB of pixels resident in HBM) and 131 072 roads start
in it, so N = 8 is exactly the named configuration and N = 1 is its single-GPU shard (weak scaling).
A step = one pass over all tiles: fused rasterize+histogram kernel over the rank's (road, tile) pairs,
all-reduce of the boundary-road table (N > 1), statistics kernel over the rank's roads.
Pixels counted = n_tiles * H * W (every tile pixel once).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 256
C = 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles-x", type=int, default=512)
    ap.add_argument("--tiles-y", type=int, default=512, help="tile rows per GPU")
    ap.add_argument("--roads-per-gpu", type=int, default=131072)
    ap.add_argument("--kind", default="uniform", choices=["uniform", "asphalt"])
    ap.add_argument("--e2e-rows", type=int, default=0, help="tile rows of each rank's shard in the host-buffer (e2e) leg; 0 = the whole "
                    "shard when half of the free host memory per rank can hold it pinned")
    ap.add_argument("--e2e-chunk", type=int, default=8192, help="tiles per streamed chunk of the e2e leg")
    ap.add_argument("--e2e-mode", default="both", choices=["both", "stream", "mapped"], help="host-buffer transport(s) to time; the "
                    "faster one is reported as e2e")
    ap.add_argument("--cpu-rows", type=int, default=32, help="tile rows of the CPU baseline sample")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: tiles-y rows and roads-per-gpu roads PER GPU "
                    "(N = 8 is the canton configuration); strong: that much IN TOTAL, sharded over the GPUs (efficiency = T1 / (N TN))")
    ap.add_argument("--balance", default="pairs", choices=["tiles", "pairs"], help="shard cuts: equal tile counts, or tile ranges "
                    "holding equal numbers of (road, tile) pairs")
    ap.add_argument("--shard-shift", type=int, default=0, help="rank r works on shard (r + shift) mod N: tells a slow shard from a slow GPU")
    ap.add_argument("--plan-world", type=int, default=0, help="N = 1 only: plan the shards of a PLAN_WORLD-GPU weak-scaling run and time shard "
                    "--shard-shift alone on this GPU (boundary rows stay partial: no merge, no oracle check)")
    ap.add_argument("--overlap", action="store_true", help="N > 1: the statistics of a rank's own rows run under the all-reduce of the "
                    "boundary rows on a second stream (measured at N = 2: no gain, 9.46 vs 9.45 ms per step -- off by default)")
    ap.add_argument("--no-legs", action="store_true", help="skip the full-size legs of the other BASELINE configurations (N = 1)")
    ap.add_argument("--leg-steps", type=int, default=5)
    ap.add_argument("--wide-grid", type=int, default=64, help="configs[4] leg: tiles per side of the 1024 px lattice (64 -> 4096 tiles)")
    ap.add_argument("--wide-polys", type=int, default=1024, help="configs[4] leg: polygons of the dense variant (the sparse one has a quarter)")
    ap.add_argument("--pageable-rows", type=int, default=96, help="tile rows of the pageable-caller e2e sample")
    ap.add_argument("--no-pageable", action="store_true")
    ap.add_argument("--compressed-rows", type=int, default=32, help="tile rows of the compressed-tiles e2e sample")
    ap.add_argument("--no-compressed", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML in-process, 10 ms period)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


class PcieRxSampler:
    """NVML PCIe RX throughput (host -> device, KB/s over the driver's 20 ms window) sampled during an e2e leg: the measured
    count of the bytes that crossed the host link, next to the bytes counted from the buffers."""

    def __init__(self, index: int):
        self.vals, self._stop, self._thr = [], threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv.nvmlDeviceGetPcieThroughput(self.h, self.nv.NVML_PCIE_UTIL_RX_BYTES)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.vals.append(float(self.nv.nvmlDeviceGetPcieThroughput(self.h, self.nv.NVML_PCIE_UTIL_RX_BYTES)))
            except Exception:  # noqa: BLE001
                self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def stop(self, seconds: float):
        """bytes received over `seconds`, or None when NVML gave no samples"""
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if not self.vals:
            return None
        return float(np.mean(self.vals)) * 1024.0 * seconds


def build_inputs(args, world, rank=0, dist=None, dev=None):
    """the GLOBAL tile lattice and road set (weak scaling: tiles_y rows and roads_per_gpu roads per GPU; strong scaling: that
    much in total, whatever the number of GPUs).  With several ranks (one node) rank 0 generates the set once and hands it
    to the others through /dev/shm instead of every rank repeating 40 s of numpy at N = 8."""
    from proj_roadsurf_b200 import synth
    from proj_roadsurf_b200.geometry import PairList, RoadSet
    mult = 1 if args.scaling == "strong" else world
    grid = synth.Grid(args.tiles_x, args.tiles_y * mult)
    n_roads = args.roads_per_gpu * mult
    if world == 1 or dist is None:
        return grid, synth.ribbon_roads(grid, n_roads)
    import shutil
    import torch
    d = f"/dev/shm/roadsurf_bench_{os.environ.get('MASTER_PORT', '0')}_{grid.nx}x{grid.ny}_{n_roads}"
    names = ("xy", "ring_off", "road_ring_off", "bbox", "ids", "road_pair_off", "pair_tile", "width_m", "n_centre", "gt_class")
    ok = torch.zeros(1, dtype=torch.int32, device=dev)
    rr = None
    if rank == 0:
        rr = synth.ribbon_roads(grid, n_roads)
        try:
            shutil.rmtree(d, ignore_errors=True)
            os.makedirs(d)
            arrs = (rr.roads.xy, rr.roads.ring_off, rr.roads.road_ring_off, rr.roads.bbox, rr.roads.ids, rr.pairs.road_pair_off,
                    rr.pairs.pair_tile, rr.width_m, rr.n_centre, rr.gt_class)
            for n, a in zip(names, arrs):
                np.save(os.path.join(d, n + ".npy"), a)
            ok += 1
        except OSError:
            pass
    dist.broadcast(ok, 0)
    shared = bool(ok.item())
    if rank != 0:
        if shared:
            a = {n: np.load(os.path.join(d, n + ".npy")) for n in names}
            rr = synth.RibbonRoads(RoadSet(a["xy"], a["ring_off"], a["road_ring_off"], a["bbox"], a["ids"]),
                                   PairList(a["road_pair_off"], a["pair_tile"]), a["width_m"], a["n_centre"], a["gt_class"])
        else:
            rr = synth.ribbon_roads(grid, n_roads)
    dist.barrier()
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    return grid, rr


def sub_problem(shard_roads, shard_pairs, n_tiles_sub):
    """roads / pairs of the first n_tiles_sub local tiles (the e2e and CPU samples)"""
    p = shard_pairs.restrict_tiles(0, n_tiles_sub)
    idx = np.nonzero(np.diff(p.road_pair_off) > 0)[0]
    return shard_roads.subset(idx), p.take_roads(idx), idx


def cpu_cores():
    return len(os.sched_getaffinity(0))


def workload_name(args, world):
    if args.scaling == "strong":
        return (f"strong scaling: {args.tiles_x}x{args.tiles_y} zoom-18 tiles of {H}x{W}x{C} uint8 and {args.roads_per_gpu} buffered "
                f"road polygons IN TOTAL, sharded by tile over {world} GPU(s)")
    return (f"canton-scale shard (BASELINE configs[3] at 8 GPUs): {args.tiles_x}x{args.tiles_y} zoom-18 tiles of "
            f"{H}x{W}x{C} uint8 and {args.roads_per_gpu} buffered road polygons per GPU, x{world} GPU(s), sharded by tile")


def load_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """DRAM bytes per launch measured under ncu for the legs of this file (profiles/traffic.json; see profiles/README.md)"""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(tpath)) if os.path.exists(tpath) else {}


# ---------------------------------------------------------------------------------------------
def oracle_rows(eng, grid, roads, pairs, road_idx, resident=None, tile_lo=0, tile_hi=0, channels=C, dtype="u8", kind=0, size=H,
                joint=False, scale=None):
    """Per-road histograms of the roads `road_idx` of (roads, pairs) from the plain-C oracle.  pairs holds GLOBAL tile indices of
    `grid`; the pixels of a tile come from `resident` (this rank's DeviceTiles, tiles [tile_lo, tile_hi)) when the rank owns it
    and are regenerated with the deterministic generator otherwise (tiles another rank owns: boundary roads)."""
    import torch
    from oracle import cport
    road_idx = np.asarray(road_idx, np.int64)
    r_s, p_s = roads.subset(road_idx), pairs.take_roads(road_idx)
    tiles_g, inv = np.unique(p_s.pair_tile.astype(np.int64), return_inverse=True)
    own = (tiles_g >= tile_lo) & (tiles_g < tile_hi) if resident is not None else np.zeros(len(tiles_g), bool)
    npdt = np.uint16 if dtype == "u16" else np.uint8
    host = np.empty((len(tiles_g), size, size, channels), npdt)
    if own.any():
        sel = torch.from_numpy(tiles_g[own] - tile_lo).to(resident.pixels.device)
        got = resident.pixels.index_select(0, sel).cpu().numpy()
        host[own] = got.view(np.uint16) if dtype == "u16" else got
    if (~own).any():
        far = eng.synth_tiles_dev(grid.keys(tiles_g[~own]), size, size, channels, dtype=dtype, kind=kind)
        got = far.pixels.cpu().numpy()
        host[~own] = got.view(np.uint16) if dtype == "u16" else got
        del far
    gt = grid.transforms(tiles_g)
    kw = {}
    if scale is not None:
        kw = dict(scale_k=scale[0], scale_off=scale[1], rescale_f32=bool(scale[2]))
    return cport.zonal_accumulate(r_s.xy, r_s.ring_off, r_s.road_ring_off, p_s.road_pair_off, inv.astype(np.int32), host, gt,
                                  joint=joint, threads=cpu_cores(), **kw)


def rows_match(gpu_hist, gpu_nz, rows, oh, onz):
    import torch
    sel = torch.as_tensor(np.asarray(rows, np.int64), device=gpu_hist.device)
    gh = gpu_hist.index_select(0, sel).cpu().numpy().view(np.uint32).astype(np.uint64)
    gz = gpu_nz.index_select(0, sel).cpu().numpy().view(np.uint32).astype(np.uint64)
    return bool(np.array_equal(gh, oh) and np.array_equal(gz, onz))


def u16_scale_range():
    """(smin, smax) per band of the fused 16 -> 8 bit rescale for the synthetic uint16 tiles (uniform 0 .. 65535, 1 % zeros):
    tif2cog.py:224-236 summarize_stats on the population statistics."""
    mean, std = 0.99 * 32767.5, float(np.sqrt(0.99 * (65535.0 ** 2 / 3.0) - (0.99 * 32767.5) ** 2))
    lo, hi = max(mean - 2.0 * std, 0.0), min(mean + 2.0 * std, 65535.0)          # identical for every band: the std over bands is 0
    return [lo] * 4, [hi] * 4


def spread(n, k):
    """k indices spread over range(n)"""
    return np.unique(np.linspace(0, n - 1, min(n, k)).astype(np.int64)) if n > 0 else np.zeros(0, np.int64)


def time_loop(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU path, restated (oracle port: GDAL scanline fill + rasterio window + masked
    extraction + per-road histograms, plain C, one thread per core), on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport
    from proj_roadsurf_b200 import synth
    from proj_roadsurf_b200.distributed import plan_shards
    world = 1          # the sample is a slice of rank 0's shard, whatever --gpus says
    grid, rr = build_inputs(args, world)
    sh = plan_shards(rr.roads, rr.pairs, grid.n_tiles, 1)[0]
    n_sub = args.tiles_x * min(args.cpu_rows, args.tiles_y)
    roads, pairs, _ = sub_problem(sh.roads, sh.pairs, n_sub)
    tiles = synth.host_tiles(synth.Grid(args.tiles_x, min(args.cpu_rows, args.tiles_y)), C, args.kind)
    gt = grid.transforms(np.arange(n_sub))
    cores = cpu_cores()
    cport.build()

    def step():
        return cport.zonal_accumulate(roads.xy, roads.ring_off, roads.road_ring_off, pairs.road_pair_off, pairs.pair_tile,
                                      tiles, gt, threads=cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    px = n_sub * H * W
    val = px * args.steps / dt / 1e9
    sample = f"first {min(args.cpu_rows, args.tiles_y)} tile rows of the shard: {n_sub} tiles, {roads.n_roads} roads, {pairs.n_pairs} pairs per step"
    print(json.dumps({
        "impl": "reference", "metric": "Gpixel/s rasterize+per-road zonal stats", "value": val, "unit": "Gpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Gpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
def config_legs(args, eng, torch, dev, grid, sh, rr_gt_class, dr, dp, peak, traffic):
    """The other BASELINE.json configurations, each at full size on this GPU, each with its kernel time, roofline fraction,
    measured DRAM bytes (when profiles/traffic.json holds them) and its own oracle flag.  Runs on one GPU after the main leg
    released its tiles."""
    from oracle import vote as ovote
    from proj_roadsurf_b200 import synth
    from proj_roadsurf_b200.engine import scale_params
    legs = {}
    n_tiles = sh.tile_hi - sh.tile_lo
    tile_idx = np.arange(sh.tile_lo, sh.tile_hi)
    gt = grid.transforms(tile_idx)
    g_pairs = sh.pairs                                # world == 1: local tile index == global tile index
    sample = spread(sh.roads.n_roads, 96)

    def leg_entry(ms, px, bpp, name_key, ok, **extra):
        gpx = px / (ms * 1e-3) / 1e9
        e = {"Gpixel/s": gpx, "kernel_ms": ms, "bytes_per_pixel": bpp, "roofline_frac": gpx * bpp / peak,
             "dram_bytes": traffic.get("legs", {}).get(name_key), "gpu_matches_oracle_on_sample": ok}
        e.update(extra)
        return e

    # ---- configs[0]/[3] pixel values, low-entropy variant: N(110, 6) values, same-bin histogram contention ----
    if args.kind != "asphalt":
        t = eng.synth_tiles_dev(grid.keys(tile_idx), H, W, C, kind=1, gt=gt)
        out = eng.zonal_hist_dev(dr, t, dp, check=False)
        ms = time_loop(torch, lambda: eng.zonal_hist_dev(dr, t, dp, out=out, check=False), args.leg_steps, 3)
        eng.sync_status()
        oh, onz = oracle_rows(eng, grid, sh.roads, g_pairs, sample, t, 0, n_tiles, kind=1)
        legs["asphalt_u8x3"] = leg_entry(ms, n_tiles * H * W, 3, "asphalt_u8x3", rows_match(out[0], out[1], sample, oh, onz),
                                         config="configs[0]/[3] shapes with low-entropy N(110, 6) pixel values", tiles=n_tiles)
        del t, out
        torch.cuda.empty_cache()

    # ---- configs[1]: class/score planes -> per-road joint histogram -> vote + confusion + F1 for 20 thresholds ----
    t = eng.synth_tiles_dev(grid.keys(tile_idx), H, W, 2, kind=2, gt=gt)
    out = eng.zonal_hist_dev(dr, t, dp, hist_mode="class_score", check=False)
    ms = time_loop(torch, lambda: eng.zonal_hist_dev(dr, t, dp, hist_mode="class_score", out=out, check=False), args.leg_steps, 3)
    eng.sync_status()
    gtc_h = np.ascontiguousarray(rr_gt_class[sh.road_global], np.int8)
    gtc = torch.from_numpy(gtc_h).to(dev)
    cuts = ovote.score_cutoffs()
    vm = [None]

    def vote():
        vm[0] = eng.vote_metrics_dev(out[0], gtc, cuts, rule="count", check=False)
    ms_vote = time_loop(torch, vote, args.leg_steps, 3)
    oh, onz = oracle_rows(eng, grid, sh.roads, g_pairs, sample, t, 0, n_tiles, channels=2, kind=2, joint=True)
    ok = rows_match(out[0], out[1], sample, oh, onz)
    # the vote on the sample rows against the oracle sweep (confusion counts exact, balanced F1 to 1e-6)
    sel = torch.as_tensor(sample, device=dev)
    _, _, conf_s, met_s = eng.vote_metrics_dev(out[0].index_select(0, sel).contiguous(), gtc.index_select(0, sel).contiguous(), cuts,
                                               rule="count")
    rows, best = ovote.sweep(oh.astype(np.uint32), gtc_h[sample], rule="count")
    conf_s, met_s = conf_s.cpu().numpy(), met_s.cpu().numpy()
    ok_vote = all(np.array_equal(conf_s[i], rows[i]["confusion"]) and abs(met_s[i, 11] - rows[i]["f1b"]) <= 1e-6 for i in range(len(cuts)))
    met_full = vm[0][3].cpu().numpy()
    legs["class_score_vote_u8x2"] = leg_entry(ms, n_tiles * H * W, 2, "class_score_vote_u8x2", bool(ok and ok_vote),
                                              config="configs[1]: synthetic class + score planes, pixel-count argmax vote, confusion, F1",
                                              vote_metrics_ms=ms_vote, thresholds=len(cuts), tiles=n_tiles,
                                              best_f1b=float(met_full[:, 11].max()))
    del t, out, vm
    torch.cuda.empty_cache()

    # ---- configs[2]: 4-band uint16 tiles, tif2cog 16 -> 8 bit rescale fused with the zonal statistics ----
    free_b, _ = torch.cuda.mem_get_info(dev)
    rows16 = int(min(args.tiles_y, (free_b * 0.88) // (args.tiles_x * H * W * 8)))
    if rows16 >= 8:
        n16 = args.tiles_x * rows16
        if n16 == n_tiles:
            roads16, pairs16, dr16, dp16 = sh.roads, g_pairs, dr, dp
        else:
            roads16, pairs16, _ = sub_problem(sh.roads, sh.pairs, n16)
            dr16, dp16 = eng.upload_roads(roads16), eng.upload_pairs(pairs16)
        t = eng.synth_tiles_dev(grid.keys(tile_idx[:n16]), H, W, 4, dtype="u16", kind=0, gt=gt[:n16])
        # scale ranges from the population, the way tif2cog.py:224-236 derives them (mean -+ 2 std of every band, then
        # mean -+ std over the bands, clamped to [0, 65535]): the synthetic uint16 values are uniform, so the ranges are the
        # whole 16 bits and the 8-bit outputs spread over all 256 bins
        smin16, smax16 = u16_scale_range()
        k, off = scale_params(smin16, smax16)
        # float64 semantics twice: the launcher's own choice (it verifies on all 65 536 inputs per band that the float32 evaluation
        # gives the same bytes for this range and then runs that policy), and the integer-threshold form it uses for ranges where
        # the two precisions differ somewhere (RS_ZONAL_F32EQ=0)
        for f32, forced in ((False, None), (False, "0"), (True, None)):
            rs = (k, off, f32) if not f32 else scale_params(smin16, smax16, True) + (True,)
            if forced is not None:
                os.environ["RS_ZONAL_F32EQ"] = forced
            out = eng.zonal_hist_dev(dr16, t, dp16, rescale=rs, check=False)
            ms = time_loop(torch, lambda: eng.zonal_hist_dev(dr16, t, dp16, rescale=rs, out=out, check=False), args.leg_steps, 3)
            eng.sync_status()
            os.environ.pop("RS_ZONAL_F32EQ", None)
            smp = spread(roads16.n_roads, 64)
            oh, onz = oracle_rows(eng, grid, roads16, pairs16, smp, t, 0, n16, channels=4, dtype="u16", scale=rs)
            name = "rescale_u16x4_" + ("f32" if f32 else "f64") + ("_thresholds" if forced is not None else "")
            how = ("float32 working precision" if f32 else
                   "float64 working precision, " + ("integer-threshold form" if forced is not None else
                                                    "evaluated in float32 after the launcher verified every input gives the same byte"))
            legs[name] = leg_entry(ms, n16 * H * W, 8, name, rows_match(out[0], out[1], smp, oh, onz),
                                   config="configs[2]: RGB+NIR uint16 tiles, gdal.Translate -scale fused into the accumulation, " + how,
                                   tiles=n16)
            del out
        del t
        torch.cuda.empty_cache()

    # ---- configs[4]: 1024 px tiles, wide polygons with holes and 1 k - 10 k vertices (long edge lists) ----
    # two densities of the same polygons: "dense" = the tiles are covered 1.25 times over (every pixel is read and histogrammed
    # once per polygon above it, so the work is 1.25 x the tile bytes), "sparse" = a third of the surface, closer to a road network
    g5 = synth.Grid(args.wide_grid, args.wide_grid, size=1024)
    t5 = eng.synth_tiles_dev(g5.keys(), 1024, 1024, 3, kind=0, gt=g5.transforms())
    for name, n_poly in (("wide_polygons_1024px", args.wide_polys), ("wide_polygons_1024px_sparse", max(1, args.wide_polys // 4))):
        wp = synth.wide_polygons(g5, n_poly)
        d5r, d5p = eng.upload_roads(wp.roads), eng.upload_pairs(wp.pairs)
        o5 = eng.zonal_hist_dev(d5r, t5, d5p, check=False)
        ms5 = time_loop(torch, lambda: eng.zonal_hist_dev(d5r, t5, d5p, out=o5, check=False), args.leg_steps, 2)
        eng.sync_status()
        smp = spread(wp.roads.n_roads, 6)
        oh, onz = oracle_rows(eng, g5, wp.roads, wp.pairs, smp, t5, 0, g5.n_tiles, size=1024)
        cov5 = float(o5[0][:, 0].sum().item()) / (g5.n_tiles * 1024.0 * 1024.0)
        nv5 = np.diff(wp.roads.ring_off[wp.roads.road_ring_off])
        legs[name] = leg_entry(ms5, g5.n_tiles * 1024 * 1024, 3, name, rows_match(o5[0], o5[1], smp, oh, onz),
                               config="configs[4]: 10 cm 1024x1024 tiles, 100-300 px wide polygons with 1-8 holes",
                               kernel="zonal_wide_kernel (rs_wide.cu)", tiles=g5.n_tiles, polygons=int(wp.roads.n_roads),
                               pairs=int(wp.pairs.n_pairs), covered_fraction=cov5,
                               covered_Gpixel_s=cov5 * g5.n_tiles * 1024 * 1024 / (ms5 * 1e-3) / 1e9,
                               mean_vertices=float(nv5.mean()), max_vertices=int(nv5.max()))
        del o5, d5r, d5p
    del t5
    torch.cuda.empty_cache()
    return legs


def run_b200(args):
    import torch
    import torch.distributed as dist

    from proj_roadsurf_b200 import synth
    from proj_roadsurf_b200.distributed import merge_boundary, plan_shards
    from proj_roadsurf_b200.engine import Engine
    from proj_roadsurf_b200.geometry import TileBatch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    if world > 1:
        # one block of host cores per rank (GPU i sits behind the socket that holds cores [i*n/N, (i+1)*n/N) on HGX
        # boards): the pinned buffers of the e2e leg are then allocated on the NUMA node next to the GPU
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except Exception:  # noqa: BLE001
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pw = args.plan_world if (world == 1 and args.plan_world > 1) else world
    grid, rr = build_inputs(args, pw, rank, dist if world > 1 else None, dev)
    shard = (rank + args.shard_shift) % pw
    sh = plan_shards(rr.roads, rr.pairs, grid.n_tiles, pw, only_rank=shard, balance=args.balance)[shard]
    n_tiles = sh.tile_hi - sh.tile_lo
    tile_idx = np.arange(sh.tile_lo, sh.tile_hi)
    gt = grid.transforms(tile_idx)
    eng = Engine(local)
    if world > 1:
        eng.comm_init_from_torch()          # the merge goes through the C ABI (rs_allreduce_accumulators_dev), not torch.distributed
    kind = {"uniform": 0, "asphalt": 1}[args.kind]
    dt_ = eng.synth_tiles_dev(grid.keys(tile_idx), H, W, C, kind=kind, gt=gt)
    dr, dp = eng.upload_roads(sh.roads), eng.upload_pairs(sh.pairs)
    slot = torch.from_numpy(sh.slot).to(dev) if pw > 1 else None
    hist = torch.zeros((sh.n_rows, C, 256), dtype=torch.int32, device=dev)
    nz = torch.zeros((sh.n_rows,), dtype=torch.int32, device=dev)
    stats = torch.empty((sh.n_rows, C, 9), dtype=torch.float64, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    overlap = world > 1 and args.overlap and sh.n_boundary > 0 and sh.n_own > 0
    side = torch.cuda.Stream(device=dev) if overlap else None
    ev_k = torch.cuda.Event() if overlap else None

    def step(i=None):
        if world > 1:                       # boundary rows of roads this rank does not hold must be zero before the sum
            hist[sh.n_own:].zero_()
            nz[sh.n_own:].zero_()
        if i is not None:
            ev[i][0].record()
        eng.zonal_hist_dev(dr, dt_, dp, road_slot=slot, out=(hist, nz), check=False)
        if i is not None:
            ev[i][1].record()
        if overlap:
            # the rows of this rank's own roads are final when the kernel ends: their statistics run while the boundary rows
            # are summed over the ranks on a second stream; the statistics of the boundary rows follow the sum
            main = torch.cuda.current_stream()
            ev_k.record(main)
            side.wait_event(ev_k)
            with torch.cuda.stream(side):
                merge_boundary(hist, nz, sh.n_own, engine=eng)
            eng.finalize_stats_dev(hist[:sh.n_own], nz[:sh.n_own], nodata_mode="none", ddof=1, out=stats[:sh.n_own], check=False)
            main.wait_stream(side)
            eng.finalize_stats_dev(hist[sh.n_own:], nz[sh.n_own:], nodata_mode="none", ddof=1, out=stats[sh.n_own:], check=False)
            return
        merge_boundary(hist, nz, sh.n_own, engine=eng)
        eng.finalize_stats_dev(hist, nz, nodata_mode="none", ddof=1, out=stats, check=False)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    eng.sync_status()
    clocks = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    l0 = eng.launch_count
    clocks.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step(i)
    t_end.record()
    torch.cuda.synchronize()
    barrier()
    clk = clocks.stop()
    launches = eng.launch_count - l0
    eng.sync_status()
    ms = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # every rank's own figures next to the maxima: kernel ms, step ms, pairs and road pixels of its shard
        mine = torch.tensor([float(kms.item()), float(ms.item()) / args.steps, float(sh.pairs.n_pairs), float(hist[:sh.n_own, 0].sum().item())],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"shard": [(r + args.shard_shift) % world for r in range(world)], "kernel_ms": [round(float(t[0]), 4) for t in allr], "step_ms": [round(float(t[1]), 4) for t in allr],
                    "pairs": [int(t[2]) for t in allr], "own_road_pixels": [int(t[3]) for t in allr]}
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(ms.item()), float(kms.item())
    px_total = float(grid.n_tiles) * H * W
    value = px_total * args.steps / (ms_total * 1e-3) / 1e9

    # ---- parity of the timed result, every rank, AFTER the merge: a spread of this rank's own roads and of the boundary
    #      roads it touches (their pixels live on several ranks: the oracle sees all their tiles) against the plain-C oracle ----
    parity, parity_note = None, None
    if not args.no_cpu:
        from oracle import cport
        cport.build()
        own_s = spread(sh.n_own, 48)
        oh, onz = oracle_rows(eng, grid, rr.roads, rr.pairs, sh.road_global[own_s], dt_, sh.tile_lo, sh.tile_hi, kind=kind)
        ok = rows_match(hist, nz, sh.slot[own_s], oh, onz)
        n_b_checked = 0
        if world > 1 and sh.n_boundary > 0:
            mine = np.arange(sh.n_own, len(sh.road_global))                   # local roads that are boundary roads
            b_s = mine[spread(len(mine), 64)]
            oh, onz = oracle_rows(eng, grid, rr.roads, rr.pairs, sh.road_global[b_s], dt_, sh.tile_lo, sh.tile_hi, kind=kind)
            ok = ok and rows_match(hist, nz, sh.slot[b_s], oh, onz)
            n_b_checked = len(b_s)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        cnt = torch.tensor([len(own_s), n_b_checked], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        parity = bool(flag.item())
        parity_note = (f"{int(cnt[0].item())} own roads and {int(cnt[1].item())} boundary roads (after the NCCL merge; all their tiles, "
                       f"also those of other ranks) over {world} rank(s) against the plain-C oracle, AND over ranks")

    # ---- roofline of the dominant kernel (fused rasterize + histogram), this rank's launch ----
    nv_road = (sh.roads.ring_off[sh.roads.road_ring_off[1:]] - sh.roads.ring_off[sh.roads.road_ring_off[:-1]]).astype(np.int64)
    edge_bytes = 16 * int((nv_road * np.diff(sh.pairs.road_pair_off)).sum())
    alg_bytes = n_tiles * H * W * C + edge_bytes + sh.n_rows * (C * 1024 + 4)
    peak, peak_src = load_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = load_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if args.scaling == "weak" and args.tiles_y == 512 else None,
                "kernel": "zonal_kernel (fused rasterize + per-road histograms)", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "algorithmic bytes = every tile byte once + 16 B per edge per pair + output rows (SURVEY 8d); the kernel is "
                        "span-driven and only touches the sectors under road pixels, so DRAM traffic is BELOW this figure"}

    # ---- e2e: host buffers through the C ABI (rs_zonal_stats_host), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        rows = min(args.e2e_rows, args.tiles_y) if args.e2e_rows > 0 else args.tiles_y
        rows = min(rows, n_tiles // args.tiles_x)
        try:                                            # the pinned copy of the tiles must fit comfortably in host memory
            import psutil
            budget = psutil.virtual_memory().available * 0.5 / world
            rows = max(1, min(rows, int(budget // (args.tiles_x * H * W * C))))
        except Exception:  # noqa: BLE001
            rows = min(rows, 64)
        n_sub = args.tiles_x * rows
        roads_s, pairs_s, _ = sub_problem(sh.roads, sh.pairs, n_sub)
        t0 = time.perf_counter()
        host_px = torch.empty((n_sub, H, W, C), dtype=torch.uint8, pin_memory=True)
        pin_s = time.perf_counter() - t0
        host_px.copy_(dt_.pixels[:n_sub])
        torch.cuda.synchronize()
        tb = TileBatch(host_px.numpy(), gt[:n_sub], H, W, C)
        chunk = min(args.e2e_chunk, n_sub)
        h2d_meta = roads_s.xy.nbytes + roads_s.bbox.nbytes + roads_s.ring_off.nbytes + roads_s.road_ring_off.nbytes + \
            pairs_s.road_pair_off.nbytes + pairs_s.pair_tile.nbytes + gt[:n_sub].nbytes
        modes = {"stream": dict(tiles_per_chunk=chunk), "mapped": dict(mapped=True)}
        if args.e2e_mode != "both":
            modes = {args.e2e_mode: modes[args.e2e_mode]}
        n_e2e, legs_e, st_ref, rx = 3, {}, None, {}
        for name, kw in modes.items():
            eng.zonal_stats_host(roads_s, tb, pairs_s, **kw)                       # warm-up (allocates the staging buffers)
            barrier()
            pcie = PcieRxSampler(local).start()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                st_host = eng.zonal_stats_host(roads_s, tb, pairs_s, **kw)
            t1 = time.perf_counter()
            got = pcie.stop(t1 - t0)
            rx[name] = None if got is None else int(got / n_e2e)
            et = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(et, op=dist.ReduceOp.MAX)
            legs_e[name] = world * n_sub * H * W * n_e2e / float(et.item()) / 1e9
            if st_ref is None:
                st_ref = st_host
            elif not np.array_equal(st_ref, st_host, equal_nan=True):
                raise SystemExit("e2e transports disagree")
        best = max(legs_e, key=legs_e.get)
        # bytes that cross the host link per step: every tile byte when streamed (counted from the buffer); when the kernel reads
        # the page-locked tiles in place, what NVML's PCIe RX counter saw over the timed passes of THIS run (fallback: the
        # ncu pcie__read_bytes ratio kept in profiles/traffic.json)
        how = {"stream": f"rs_zonal_stats_stream_host ({chunk}-tile chunks through two device buffers, copy overlapped with compute)",
               "mapped": "rs_zonal_stats_mapped_host (zonal_kernel reads the page-locked tiles in place; only the sectors under "
                         "road pixels cross the host link)"}[best]
        if best == "stream":
            h2d_px, h2d_src = host_px.numel(), "counted from the copied buffers"
        elif rx.get("mapped"):
            h2d_px, h2d_src = max(0, rx["mapped"] - h2d_meta), "NVML PCIe RX counter over the timed passes of this run"
        else:
            frac = traffic.get("mapped_host_tiles", {}).get("pcie_read_bytes_per_tile_byte", 1.0)
            h2d_px, h2d_src = int(host_px.numel() * min(1.0, frac)), "ncu pcie__read_bytes ratio of profiles/traffic.json (NVML gave no samples)"
        e2e = {"value": legs_e[best], "unit": "Gpixel/s", "h2d_bytes_per_step": int(h2d_px + h2d_meta),
               "d2h_bytes_per_step": int(st_host.nbytes), "steps": n_e2e, "transport": best,
               "transports_Gpixel_s": legs_e, "host_tile_bytes": int(host_px.numel()), "h2d_bytes_source": h2d_src,
               "pcie_rx_bytes_per_step_nvml": rx, "pinned_alloc_s": pin_s,
               "sample": f"first {rows} of {n_tiles // args.tiles_x} tile rows of each rank's shard ({n_sub} tiles, {roads_s.n_roads} roads) in "
                         f"pinned host memory through {how}; statistics table read back"}
        # what a caller with an ordinary (pageable) buffer gets: the streamed copies run from pageable memory, or the buffer is
        # page-locked first (rs_host_register: the cost is reported, it is a one-off per buffer)
        if world == 1 and not args.no_pageable:
            prow = max(1, min(rows, args.pageable_rows))
            n_p = args.tiles_x * prow
            roads_p, pairs_p, _ = sub_problem(sh.roads, sh.pairs, n_p)
            pageable = np.empty((n_p, H, W, C), np.uint8)
            pageable[:] = host_px.numpy()[:n_p]
            tbp = TileBatch(pageable, gt[:n_p], H, W, C)
            eng.zonal_stats_host(roads_p, tbp, pairs_p, tiles_per_chunk=chunk)
            t0 = time.perf_counter()
            st_p = eng.zonal_stats_host(roads_p, tbp, pairs_p, tiles_per_chunk=chunk)
            t1 = time.perf_counter()
            eng.pin_host(pageable)
            t2 = time.perf_counter()
            st_m = eng.zonal_stats_host(roads_p, tbp, pairs_p, mapped=True)
            t3 = time.perf_counter()
            eng.unpin_host(pageable)
            e2e["pageable_caller"] = {
                "streamed_from_pageable_Gpixel_s": n_p * H * W / (t1 - t0) / 1e9,
                "register_s": t2 - t1, "register_GB": pageable.nbytes / 1e9,
                "register_then_in_place_first_call_Gpixel_s": n_p * H * W / (t3 - t1) / 1e9,
                "results_equal": bool(np.array_equal(st_p, st_m, equal_nan=True)),
                "sample": f"first {prow} tile rows ({n_p} tiles) in an ordinary numpy buffer"}
            del pageable
        # tiles that are still compressed, as GeoTIFF strips are on disk (deflate): only the compressed bytes cross the host link,
        # decompression + assembly + statistics on the device (rs_zonal_stats_compressed_host).  Low-entropy "asphalt" tiles,
        # the kind deflate can shrink; the compression itself is file preparation and is not timed.
        if world == 1 and not args.no_compressed:
            import zlib
            from concurrent.futures import ThreadPoolExecutor
            crow = max(1, min(rows, args.compressed_rows))
            n_c = args.tiles_x * crow
            roads_c, pairs_c, _ = sub_problem(sh.roads, sh.pairs, n_c)
            tc = eng.synth_tiles_dev(grid.keys(tile_idx[:n_c]), H, W, C, kind=1)
            host_c = tc.pixels.cpu().numpy()
            del tc
            strips = 32                                                 # 8-row strips: 32 segments per tile, a decoder thread each
            flat = host_c.reshape(n_c * strips, -1)
            with ThreadPoolExecutor(max_workers=cpu_cores()) as ex:
                comp_l = list(ex.map(lambda i: zlib.compress(flat[i].tobytes(), 1), range(len(flat))))
            comp_off = np.zeros(len(comp_l) + 1, np.int64)
            comp_off[1:] = np.cumsum([len(c) for c in comp_l])
            raw_off = np.arange(len(comp_l) + 1, dtype=np.int64) * flat.shape[1]
            comp = np.frombuffer(b"".join(comp_l), np.uint8)
            del comp_l
            eng.zonal_stats_compressed_host(roads_c, gt[:n_c], H, W, C, pairs_c, comp, comp_off, raw_off)
            t0 = time.perf_counter()
            st_c = eng.zonal_stats_compressed_host(roads_c, gt[:n_c], H, W, C, pairs_c, comp, comp_off, raw_off)
            t1 = time.perf_counter()
            raw_dec = eng.decode_segments_host(comp[:int(comp_off[strips * 64])], comp_off[:strips * 64 + 1], 8, raw_off[:strips * 64 + 1])
            st_u = eng.zonal_stats_host(roads_c, TileBatch(host_c, gt[:n_c], H, W, C), pairs_c, tiles_per_chunk=chunk)
            e2e["compressed_tiles"] = {
                "Gpixel_s": n_c * H * W / (t1 - t0) / 1e9, "codec": "deflate (zlib level 1), 8-row strips",
                "compressed_bytes": int(comp.nbytes), "raw_bytes": int(host_c.nbytes), "ratio": float(host_c.nbytes / max(1, comp.nbytes)),
                "segments": int(len(comp_off) - 1), "results_equal_uncompressed_path": bool(np.array_equal(st_c, st_u, equal_nan=True)),
                "decoded_bytes_match": bool(np.array_equal(raw_dec, host_c.reshape(-1)[:len(raw_dec)])),
                "sample": f"first {crow} tile rows ({n_c} low-entropy tiles) as deflate strips through rs_zonal_stats_compressed_host"}
            del host_c, comp
        del host_px

    # ---- CPU baseline beside it (rank 0, N = 1 only) + parity of the same sample ----
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import cport
        rows = min(args.cpu_rows, args.tiles_y)
        n_sub = args.tiles_x * rows
        roads_s, pairs_s, idx = sub_problem(sh.roads, sh.pairs, n_sub)
        tiles_h = dt_.pixels[:n_sub].cpu().numpy()
        cores = cpu_cores()
        t0 = time.perf_counter()
        oh, onz = cport.zonal_accumulate(roads_s.xy, roads_s.ring_off, roads_s.road_ring_off, pairs_s.road_pair_off,
                                         pairs_s.pair_tile, tiles_h, gt[:n_sub], threads=cores)
        t1 = time.perf_counter()
        dts = eng.upload_tiles(TileBatch(tiles_h, gt[:n_sub], H, W, C))
        drs, dps = eng.upload_roads(roads_s), eng.upload_pairs(pairs_s)
        gh, gz = eng.zonal_hist_dev(drs, dts, dps)
        gpu_ms = time_loop(torch, lambda: eng.zonal_hist_dev(drs, dts, dps, out=(gh, gz), check=False), 5, 1)
        ok = bool(np.array_equal(gh.cpu().numpy().view(np.uint32).astype(np.uint64), oh) and
                  np.array_equal(gz.cpu().numpy().view(np.uint32).astype(np.uint64), onz))
        cpu = {"value": n_sub * H * W / (t1 - t0) / 1e9, "unit": "Gpixel/s", "cores": cores, "kind": "port",
               "sample": f"first {rows} tile rows of the shard ({n_sub} tiles, {roads_s.n_roads} roads, {pairs_s.n_pairs} pairs), "
                         "one pass of the plain-C oracle, one thread per core",
               "gpu_matches_oracle_on_sample": ok, "gpu_kernel_ms_same_sample": gpu_ms}
        del dts, gh, gz

    # ---- reference-shaped CPU variant: per-pair Python loop + DataFrame concat + groupby, 1 core ----
    if cpu is not None:
        from oracle import raster as oraster, stats as ostats
        import pandas as pd
        gA = synth.Grid(4, 4)
        rrA = synth.ribbon_roads(gA, 8, seed=3)
        tA = synth.host_tiles(gA, 3)
        gtA = gA.transforms()
        road_of = rrA.pairs.road_of_pair()
        t0 = time.perf_counter()
        pix = pd.DataFrame()
        for pA in range(rrA.pairs.n_pairs):
            tile = {"data": tA[rrA.pairs.pair_tile[pA]], "transform": tuple(gtA[rrA.pairs.pair_tile[pA]]), "nodata": None}
            try:
                pix = oraster.get_pixel_values(rrA.roads.rings(int(road_of[pA])), tile, range(1, 4), pix, road_id=int(road_of[pA]))
            except ValueError:
                pass
        for b in (1, 2, 3):
            ostats.get_df_stats_groupby(pix, f"band{b}", ["road_id"], f"_{b}")
        t1 = time.perf_counter()
        cpu["reference_shaped"] = {"value": gA.n_tiles * H * W / (t1 - t0) / 1e9, "unit": "Gpixel/s", "cores": 1,
                                   "sample": f"4x4 tiles, 8 roads, {rrA.pairs.n_pairs} pairs: per-pair Python loop + DataFrame concat + "
                                             "pandas groupby (statistical_analysis.py:180-246 shape), pure-Python oracle"}

    # ---- the other BASELINE configurations at full size (one GPU) ----
    legs = None
    if world == 1 and not args.no_legs:
        del dt_, hist, nz, stats
        torch.cuda.empty_cache()
        legs = config_legs(args, eng, torch, dev, grid, sh, rr.gt_class, dr, dp, peak, traffic)

    if rank == 0:
        line = {
            "metric": "Gpixel/s rasterize+per-road zonal stats", "value": value, "unit": "Gpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "tiles_rank0": n_tiles, "roads_rank0": int(sh.roads.n_roads),
                       "pairs_rank0": int(sh.pairs.n_pairs), "boundary_roads": int(sh.n_boundary), "tile_kind": args.kind,
                       "shard_cuts": args.balance,
                       "merge": "rs_allreduce_accumulators_dev: one grouped NCCL all-reduce of the boundary rows" if world > 1 else "none",
                       "l2": f"inputs ({n_tiles * H * W * C / 1e9:.1f} GB per GPU) are larger than L2; no flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "gpu_matches_oracle_on_sample": parity, "parity_sample": parity_note,
        }
        if legs is not None:
            line["configs"] = legs
        if per_rank is not None:
            line["per_rank"] = per_rank
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
