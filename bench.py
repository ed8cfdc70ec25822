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
    ap.add_argument("--extras", action="store_true", help="also time the class/score (config 2) and uint16 rescale (config 3) kernels")
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


def build_inputs(args, world):
    from proj_roadsurf_b200 import synth
    grid = synth.Grid(args.tiles_x, args.tiles_y * world)
    rr = synth.ribbon_roads(grid, args.roads_per_gpu * world)
    return grid, rr


def sub_problem(shard_roads, shard_pairs, n_tiles_sub):
    """roads / pairs of the first n_tiles_sub local tiles (the e2e and CPU samples)"""
    p = shard_pairs.restrict_tiles(0, n_tiles_sub)
    idx = np.nonzero(np.diff(p.road_pair_off) > 0)[0]
    return shard_roads.subset(idx), p.take_roads(idx), idx


def cpu_cores():
    return len(os.sched_getaffinity(0))


def workload_name(args, world):
    return (f"canton-scale shard (BASELINE configs[3] at 8 GPUs): {args.tiles_x}x{args.tiles_y} zoom-18 tiles of "
            f"{H}x{W}x{C} uint8 and {args.roads_per_gpu} buffered road polygons per GPU, x{world} GPU(s), sharded by tile")


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
    n_sub = args.tiles_x * args.cpu_rows
    roads, pairs, _ = sub_problem(sh.roads, sh.pairs, n_sub)
    tiles = synth.host_tiles(synth.Grid(args.tiles_x, args.cpu_rows), C, args.kind)
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
    sample = f"first {args.cpu_rows} tile rows of the shard: {n_sub} tiles, {roads.n_roads} roads, {pairs.n_pairs} pairs per step"
    print(json.dumps({
        "impl": "reference", "metric": "Gpixel/s rasterize+per-road zonal stats", "value": val, "unit": "Gpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Gpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
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

    grid, rr = build_inputs(args, world)
    sh = plan_shards(rr.roads, rr.pairs, grid.n_tiles, world, only_rank=rank)[rank]
    del rr
    n_tiles = sh.tile_hi - sh.tile_lo
    tile_idx = np.arange(sh.tile_lo, sh.tile_hi)
    gt = grid.transforms(tile_idx)
    eng = Engine(local)
    kind = {"uniform": 0, "asphalt": 1}[args.kind]
    dt_ = eng.synth_tiles_dev(grid.keys(tile_idx), H, W, C, kind=kind, gt=gt)
    dr, dp = eng.upload_roads(sh.roads), eng.upload_pairs(sh.pairs)
    slot = torch.from_numpy(sh.slot).to(dev) if world > 1 else None
    hist = torch.zeros((sh.n_rows, C, 256), dtype=torch.int32, device=dev)
    nz = torch.zeros((sh.n_rows,), dtype=torch.int32, device=dev)
    stats = torch.empty((sh.n_rows, C, 9), dtype=torch.float64, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def step(i=None):
        if world > 1:
            hist[sh.n_own:].zero_()
            nz[sh.n_own:].zero_()
        if i is not None:
            ev[i][0].record()
        eng.zonal_hist_dev(dr, dt_, dp, road_slot=slot, out=(hist, nz), check=False)
        if i is not None:
            ev[i][1].record()
        merge_boundary(hist, nz, sh.n_own)
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
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(ms.item()), float(kms.item())
    px_total = float(grid.n_tiles) * H * W
    value = px_total * args.steps / (ms_total * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (fused rasterize + histogram), this rank's launch ----
    nv_road = (sh.roads.ring_off[sh.roads.road_ring_off[1:]] - sh.roads.ring_off[sh.roads.road_ring_off[:-1]]).astype(np.int64)
    edge_bytes = 16 * int((nv_road * np.diff(sh.pairs.road_pair_off)).sum())
    alg_bytes = n_tiles * H * W * C + edge_bytes + sh.n_rows * (C * 1024 + 4)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "zonal_kernel (fused rasterize + per-road histograms)", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "algorithmic bytes = every tile byte once + 16 B per edge per pair + output rows (SURVEY 8d); the kernel is "
                        "span-driven and only touches the sectors under road pixels, so DRAM traffic is BELOW this figure"}

    # ---- e2e: host buffers through the C ABI (rs_zonal_stats_host), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        rows = min(args.e2e_rows, args.tiles_y) if args.e2e_rows > 0 else args.tiles_y
        try:                                            # the pinned copy of the tiles must fit comfortably in host memory
            import psutil
            budget = psutil.virtual_memory().available * 0.5 / world
            rows = max(1, min(rows, int(budget // (args.tiles_x * H * W * C))))
        except Exception:  # noqa: BLE001
            rows = min(rows, 64)
        n_sub = args.tiles_x * rows
        roads_s, pairs_s, _ = sub_problem(sh.roads, sh.pairs, n_sub)
        host_px = torch.empty((n_sub, H, W, C), dtype=torch.uint8, pin_memory=True)
        host_px.copy_(dt_.pixels[:n_sub])
        torch.cuda.synchronize()
        tb = TileBatch(host_px.numpy(), gt[:n_sub], H, W, C)
        chunk = min(args.e2e_chunk, n_sub)
        h2d_meta = roads_s.xy.nbytes + roads_s.bbox.nbytes + roads_s.ring_off.nbytes + roads_s.road_ring_off.nbytes + \
            pairs_s.road_pair_off.nbytes + pairs_s.pair_tile.nbytes + gt[:n_sub].nbytes
        modes = {"stream": dict(tiles_per_chunk=chunk), "mapped": dict(mapped=True)}
        if args.e2e_mode != "both":
            modes = {args.e2e_mode: modes[args.e2e_mode]}
        n_e2e, legs, st_ref, rx = 3, {}, None, {}
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
            legs[name] = world * n_sub * H * W * n_e2e / float(et.item()) / 1e9
            if st_ref is None:
                st_ref = st_host
            elif not np.array_equal(st_ref, st_host, equal_nan=True):
                raise SystemExit("e2e transports disagree")
        best = max(legs, key=legs.get)
        # bytes that cross the host link per step: every tile byte when streamed; when the kernel reads the page-locked tiles
        # in place, the PCIe read bytes ncu counted for that kernel per tile byte (profiles/traffic.json, mapped_host_tiles)
        h2d_px = host_px.numel() if best == "stream" else None
        how = {"stream": f"rs_zonal_stats_stream_host ({chunk}-tile chunks through two device buffers, copy overlapped with compute)",
               "mapped": "rs_zonal_stats_mapped_host (zonal_kernel reads the page-locked tiles in place; only the sectors under "
                         "road pixels cross the host link)"}[best]
        if h2d_px is None:
            frac = 1.0
            if os.path.exists(tpath):            # ncu pcie__read_bytes of the in-place kernel per tile byte (same synthetic roads)
                frac = json.load(open(tpath)).get("mapped_host_tiles", {}).get("pcie_read_bytes_per_tile_byte", 1.0)
            h2d_px = int(host_px.numel() * min(1.0, frac))
        e2e = {"value": legs[best], "unit": "Gpixel/s", "h2d_bytes_per_step": int(h2d_px + h2d_meta),
               "d2h_bytes_per_step": int(st_host.nbytes), "steps": n_e2e, "transport": best,
               "transports_Gpixel_s": legs, "host_tile_bytes": int(host_px.numel()),
               "pcie_rx_bytes_per_step_nvml": rx,
               "sample": f"first {rows} of {args.tiles_y} tile rows of each rank's shard ({n_sub} tiles, {roads_s.n_roads} roads) in pinned "
                         f"host memory through {how}; statistics table read back"}
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
        cport.build()
        t0 = time.perf_counter()
        oh, onz = cport.zonal_accumulate(roads_s.xy, roads_s.ring_off, roads_s.road_ring_off, pairs_s.road_pair_off,
                                         pairs_s.pair_tile, tiles_h, gt[:n_sub], threads=cores)
        t1 = time.perf_counter()
        dts = eng.upload_tiles(TileBatch(tiles_h, gt[:n_sub], H, W, C))
        gh, gz = eng.zonal_hist_dev(eng.upload_roads(roads_s), dts, eng.upload_pairs(pairs_s))
        ok = bool(np.array_equal(gh.cpu().numpy().view(np.uint32).astype(np.uint64), oh) and
                  np.array_equal(gz.cpu().numpy().view(np.uint32).astype(np.uint64), onz))
        cpu = {"value": n_sub * H * W / (t1 - t0) / 1e9, "unit": "Gpixel/s", "cores": cores, "kind": "port",
               "sample": f"first {rows} tile rows of the shard ({n_sub} tiles, {roads_s.n_roads} roads, {pairs_s.n_pairs} pairs), "
                         "one pass of the plain-C oracle, one thread per core",
               "gpu_matches_oracle_on_sample": ok}

    # ---- optional: the other tile formats of BASELINE.json configs[1] / configs[2] on a 128 x 128 tile block ----
    extras = None
    if args.extras and world == 1:
        from proj_roadsurf_b200.engine import scale_params
        g2 = synth.Grid(128, 128)
        rr2 = synth.ribbon_roads(g2, 8192)
        dr2, dp2 = eng.upload_roads(rr2.roads), eng.upload_pairs(rr2.pairs)
        gt2 = g2.transforms()
        extras = {}
        for name, ch, dtype, kind, kw, bpp in (("class_score_u8x2", 2, "u8", 2, {"hist_mode": "class_score"}, 2),
                                               ("rescale_u16x4", 4, "u16", 0, {"rescale": scale_params([150.0] * 4, [9000.0] * 4) + (False,)}, 8),
                                               ("bands_u8x3", 3, "u8", 0, {}, 3)):
            t2 = eng.synth_tiles_dev(g2.keys(), H, W, ch, dtype=dtype, kind=kind, gt=gt2)
            for _ in range(3):
                out2 = eng.zonal_hist_dev(dr2, t2, dp2, check=False, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.zonal_hist_dev(dr2, t2, dp2, out=out2, check=False, **kw)
            e1.record()
            torch.cuda.synchronize()
            eng.sync_status()
            ms2 = e0.elapsed_time(e1) / 10
            gpx = g2.n_tiles * H * W / (ms2 * 1e-3) / 1e9
            extras[name] = {"Gpixel/s": gpx, "kernel_ms": ms2, "bytes_per_pixel": bpp,
                            "roofline_frac": gpx * bpp / peak, "tiles": g2.n_tiles}
            del t2, out2
        # BASELINE configs[4]: 1024 px tiles, wide polygons with holes and 1 k - 10 k vertices (long edge lists)
        g5 = synth.Grid(32, 32, size=1024)
        wp = synth.wide_polygons(g5, 384)
        t5 = eng.synth_tiles_dev(g5.keys(), 1024, 1024, 3, kind=0, gt=g5.transforms())
        d5r, d5p = eng.upload_roads(wp.roads), eng.upload_pairs(wp.pairs)
        for _ in range(2):
            o5 = eng.zonal_hist_dev(d5r, t5, d5p, check=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.zonal_hist_dev(d5r, t5, d5p, out=o5, check=False)
        e1.record()
        torch.cuda.synchronize()
        eng.sync_status()
        ms5 = e0.elapsed_time(e1) / 5
        cov5 = float(o5[0][:, 0].sum().item()) / (g5.n_tiles * 1024.0 * 1024.0)
        gpx5 = g5.n_tiles * 1024 * 1024 / (ms5 * 1e-3) / 1e9
        extras["wide_polygons_1024px"] = {"Gpixel/s": gpx5, "kernel_ms": ms5, "bytes_per_pixel": 3, "roofline_frac": gpx5 * 3 / peak,
                                          "tiles": g5.n_tiles, "polygons": 384, "pairs": wp.pairs.n_pairs, "covered_fraction": cov5,
                                          "mean_vertices": float(np.diff(wp.roads.ring_off[wp.roads.road_ring_off]).mean())}
        del t5, o5
        # the materialising 16 -> 8 bit pass (tif2cog.py:260-270): 8 B read + 4 B written per pixel, pure HBM streaming
        n_t = 16384
        t16 = eng.synth_tiles_dev(np.arange(n_t, dtype=np.int64), H, W, 4, dtype="u16", kind=0)
        o8 = torch.empty((n_t, H, W, 4), dtype=torch.uint8, device=dev)
        for f32 in (False, True):
            for _ in range(3):
                eng.rescale_u16_dev(t16.pixels, [150.0] * 4, [9000.0] * 4, bidx=[1, 2, 3, 0], f32=f32, out=o8)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.rescale_u16_dev(t16.pixels, [150.0] * 4, [9000.0] * 4, bidx=[1, 2, 3, 0], f32=f32, out=o8)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / 10
            gbs = n_t * H * W * 12 / (ms2 * 1e-3) / 1e9
            extras["rescale_u16_to_u8_" + ("f32" if f32 else "f64")] = {"Gpixel/s": n_t * H * W / (ms2 * 1e-3) / 1e9, "kernel_ms": ms2,
                                                                      "bytes_per_pixel": 12, "GB/s": gbs, "roofline_frac": gbs / peak}
        del t16, o8

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

    if rank == 0:
        line = {
            "metric": "Gpixel/s rasterize+per-road zonal stats", "value": value, "unit": "Gpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "tiles_per_gpu": n_tiles, "roads_rank0": int(sh.roads.n_roads),
                       "pairs_rank0": int(sh.pairs.n_pairs), "boundary_roads": int(sh.n_boundary), "tile_kind": args.kind,
                       "l2": f"inputs ({n_tiles * H * W * C / 1e9:.1f} GB per GPU) are larger than L2; no flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        if extras is not None:
            line["extras"] = extras
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
