"""roadsurf-b200: B200-native raster-vector overlay hot path of proj-roadsurf.

Layout
  csrc/                hand-written sm_100a CUDA kernels + the C ABI (include/roadsurf_b200.h)
  engine.py            batched Python face of the C ABI (host-buffer and device-tensor families)
  geometry.py          road polygon soup / tile batch / pair list containers
  synth.py             deterministic synthetic inputs of the named shapes
  functions/           drop-in mirrors of the reference's scripts/functions helpers
  road_segmentation/   drop-in mirrors of determine_class.py / final_metrics.py helpers
  distributed.py       tile sharding over the GPUs of one box + NCCL merge of per-road accumulators
There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
