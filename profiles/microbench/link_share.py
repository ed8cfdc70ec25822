"""Does the host link have spare capacity while zonal_kernel reads page-locked tiles in place?
Thread A: rs_zonal_stats_mapped_host over the whole benchmark shard.  Thread B: bulk H2D copies on another stream.
Prints the mapped rate alone, the copy rate alone, and both when they run together."""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from proj_roadsurf_b200 import _build, synth
if os.environ.get("RS_LIB"):                                   # experiment builds of the library
    _build.LIB_PATH = os.path.join(ROOT, os.environ["RS_LIB"])
    _build.is_stale = lambda: False
from proj_roadsurf_b200.engine import Engine
from proj_roadsurf_b200.geometry import TileBatch

TX, TY, R = 512, int(os.environ.get("TY", "256")), int(os.environ.get("ROADS", "65536"))
H = W = 256; C = 3
eng = Engine(0)
if os.environ.get("L2GRAN"):                                    # cudaLimitMaxL2FetchGranularity (0x05): 32 / 64 / 128
    import ctypes, glob
    cand = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = ctypes.CDLL(cand[0])
    torch.zeros(1, device="cuda")
    v = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2GRAN"])))
    v2 = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(v2), 5)
    print("L2 fetch granularity", v.value, "->", v2.value, "rc", rc)
grid = synth.Grid(TX, TY)
rr = synth.ribbon_roads(grid, R)
gt = grid.transforms()
dev = torch.device("cuda", 0)
host_px = torch.empty((grid.n_tiles, H, W, C), dtype=torch.uint8, pin_memory=True)
step = 8192
for lo in range(0, grid.n_tiles, step):                      # synthetic tiles generated on the device, chunk by chunk
    hi = min(grid.n_tiles, lo + step)
    host_px[lo:hi].copy_(eng.synth_tiles_dev(grid.keys()[lo:hi], H, W, C).pixels)
torch.cuda.synchronize()
tb = TileBatch(host_px.numpy(), gt, H, W, C)
px = grid.n_tiles * H * W

def mapped():
    t0 = time.perf_counter(); eng.zonal_stats_host(rr.roads, tb, rr.pairs, mapped=True); return time.perf_counter() - t0

mapped()
t_alone = min(mapped() for _ in range(2))
print(f"mapped alone: {t_alone*1e3:.1f} ms  {px/t_alone/1e9:.1f} Gpixel/s")

if os.environ.get("MAPPED_ONLY"):
    sys.exit(0)
src = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
cs = torch.cuda.Stream()
def copy_for(seconds, out):
    n = 0; t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        while time.perf_counter() - t0 < seconds:
            dst.copy_(src, non_blocking=True); cs.synchronize(); n += 1
    out.append(n * src.numel() / (time.perf_counter() - t0) / 1e9)
o = []; copy_for(1.0, o); print(f"copy alone: {o[0]:.1f} GB/s")
o = []
th = threading.Thread(target=copy_for, args=(t_alone * 1.5, o)); th.start()
time.sleep(0.05)
t_both = mapped(); th.join()
print(f"together: mapped {t_both*1e3:.1f} ms ({px/t_both/1e9:.1f} Gpixel/s), copy {o[0]:.1f} GB/s (over a window 1.5x the mapped-alone time)")
