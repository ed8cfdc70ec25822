"""Tile ingest: GeoTIFF tiles -> the pixel-interleaved batches the overlay kernels read.

The reference reads every tile with ``rasterio.open(tile).read()`` inside its pair loop (scripts/functions/fct_misc.py:76-77)
and gets the tiles from a tile server that selects and reorders bands (``bidx=2&bidx=3&bidx=4&bidx=1``,
config/config_stats.yaml:39) after tif2cog's 16 -> 8 bit rescale (scripts/preprocessing/tif2cog.py:260-270).  Here the
host only parses the TIFF directories; the compressed segments go to the device as they are in the files and are decoded
there (rs_decode_segments: DEFLATE through table-driven decoders stepped warp-wide, TIFF LZW a thread per segment), and everything per pixel -- TIFF predictor 2, byte
order, band-sequential -> interleaved, band selection, rescale -- is one kernel over the whole batch (rs_assemble_tiles_*).
``load_tiles(..., device_decode=False)`` keeps the earlier host inflate (zlib, a thread per file) as the comparison path.

Supported: classic TIFF and BigTIFF, little / big endian, strips or one internal tile row per image width, compression
none (1), LZW (5, device path) and deflate (8, 32946), predictor 1 / 2, PlanarConfiguration 1 / 2, unsigned 8 / 16 bit samples, 1-4 bands,
GeoTIFF ModelPixelScale + ModelTiepoint or ModelTransformation, GDAL_NODATA.  Anything else raises ``UnsupportedTiff``
(fct_misc.open_tile then falls back to rasterio / PIL where installed).
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .geometry import TileBatch


class UnsupportedTiff(ValueError):
    pass


@dataclass
class TiffInfo:
    width: int
    height: int
    channels: int
    sample_bytes: int
    planar: int
    predictor: int
    compression: int
    big_endian: bool
    seg_offsets: Tuple[int, ...]
    seg_counts: Tuple[int, ...]
    rows_per_seg: int
    seg_width: int            # width of a stored segment row (image width for strips, TileWidth for internal tiles)
    transform: Tuple[float, float, float, float, float, float]
    nodata: Optional[float]
    tiled: bool = False       # internal tiles (the last one may be padded below the image)

    @property
    def layout(self):
        return (self.height, self.width, self.channels, self.sample_bytes, self.planar, self.predictor, self.big_endian)


_TYPE = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2), 9: ("i", 4),
         10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}


def parse_tiff(buf: bytes) -> TiffInfo:
    """First image file directory of a classic TIFF."""
    if len(buf) < 8:
        raise UnsupportedTiff("not a TIFF")
    if buf[:2] == b"II":
        e = "<"
    elif buf[:2] == b"MM":
        e = ">"
    else:
        raise UnsupportedTiff("not a TIFF")
    (magic,) = struct.unpack(e + "H", buf[2:4])
    if magic == 42:                                         # classic: 4-byte offsets, 12-byte directory entries
        (ifd,) = struct.unpack(e + "I", buf[4:8])
        (n,) = struct.unpack(e + "H", buf[ifd:ifd + 2])
        first, esz, head, inline, ofmt = ifd + 2, 12, "HHI", 4, "I"
    elif magic == 43:                                       # BigTIFF: 8-byte offsets, 20-byte directory entries
        osz, zero, ifd = struct.unpack(e + "HHQ", buf[4:16])
        if osz != 8 or zero != 0:
            raise UnsupportedTiff("BigTIFF with an offset size other than 8")
        (n,) = struct.unpack(e + "Q", buf[ifd:ifd + 8])
        first, esz, head, inline, ofmt = ifd + 8, 20, "HHQ", 8, "Q"
    else:
        raise UnsupportedTiff("unknown TIFF magic")
    tags = {}
    for i in range(n):
        ent = buf[first + esz * i: first + esz * (i + 1)]
        tag, typ, cnt = struct.unpack(e + head, ent[:esz - inline])
        raw = ent[esz - inline:]
        if typ not in _TYPE:
            continue
        fmt, size = _TYPE[typ]
        nbytes = size * cnt
        data = raw[:nbytes] if nbytes <= inline else buf[struct.unpack(e + ofmt, raw)[0]:][:nbytes]
        if typ == 2:
            tags[tag] = data.split(b"\x00")[0].decode("latin-1")
        elif typ in (5, 10):
            v = struct.unpack(e + fmt[0] * (2 * cnt), data)
            tags[tag] = tuple(v[2 * j] / v[2 * j + 1] if v[2 * j + 1] else 0.0 for j in range(cnt))
        else:
            tags[tag] = struct.unpack(e + fmt * cnt, data)

    def one(tag, default=None):
        v = tags.get(tag)
        return default if v is None else (v[0] if isinstance(v, tuple) else v)

    W, H = one(256), one(257)
    if W is None or H is None:
        raise UnsupportedTiff("no image size")
    C = int(one(277, 1))
    bits = tags.get(258, (1,))
    if len(set(bits)) != 1 or bits[0] not in (8, 16):
        raise UnsupportedTiff(f"BitsPerSample {bits}")
    if any(f != 1 for f in tags.get(339, (1,))):
        raise UnsupportedTiff("SampleFormat is not unsigned integer")
    if not 1 <= C <= 4:
        raise UnsupportedTiff(f"{C} samples per pixel")
    comp = int(one(259, 1))
    if comp not in (1, 5, 8, 32946):
        raise UnsupportedTiff(f"compression {comp}")
    pred = int(one(317, 1))
    if pred not in (1, 2):
        raise UnsupportedTiff(f"predictor {pred}")
    if comp == 1:
        pred = 1            # libtiff implements the predictor inside its LZW / deflate codecs: uncompressed data ignores the tag
    planar = int(one(284, 1))
    if 322 in tags:                                         # internal tiles
        tw, th = int(one(322)), int(one(323))
        if tw < W:
            raise UnsupportedTiff("internal tiles narrower than the image")
        offs, cnts, rps, sw = tags[324], tags[325], th, tw
    else:
        rps = min(int(one(278, H)), H)
        offs, cnts, sw = tags[273], tags[279], W
    scale, tie, mt = tags.get(33550), tags.get(33922), tags.get(34264)
    if scale is not None and tie is not None:
        i, j, _, x, y, _ = [float(v) for v in tie[:6]]
        sx, sy = float(scale[0]), float(scale[1])
        transform = (sx, 0.0, x - i * sx, 0.0, -sy, y + j * sy)
    elif mt is not None:
        m = [float(v) for v in mt]
        transform = (m[0], m[1], m[3], m[4], m[5], m[7])
    else:
        transform = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
    nodata = None
    nd = tags.get(42113)
    if nd is not None:
        try:
            nodata = float(str(nd).strip())
            nodata = int(nodata) if nodata == int(nodata) else nodata
        except ValueError:
            nodata = None
    return TiffInfo(int(W), int(H), C, bits[0] // 8, planar, pred, comp, e == ">", tuple(int(o) for o in offs),
                    tuple(int(c) for c in cnts), int(rps), int(sw), transform, nodata, 322 in tags)


def read_raw(path_or_bytes) -> Tuple[np.ndarray, TiffInfo]:
    """The decompressed samples of the first image, still in file byte order and still predictor-encoded, as a uint8 array
    laid out [H][W][C] (planar 1) or [C][H][W] (planar 2)."""
    buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    info = parse_tiff(buf)
    H, W, C, sb = info.height, info.width, info.channels, info.sample_bytes
    spp = C if info.planar == 1 else 1                      # samples per pixel inside one segment
    planes = 1 if info.planar == 1 else C
    segs_per_plane = (H + info.rows_per_seg - 1) // info.rows_per_seg
    if len(info.seg_offsets) < planes * segs_per_plane:
        raise UnsupportedTiff("segment table shorter than the image")
    out = np.empty((planes, H, W * spp * sb), np.uint8)
    row_bytes = info.seg_width * spp * sb
    for pl in range(planes):
        for s in range(segs_per_plane):
            k = pl * segs_per_plane + s
            data = buf[info.seg_offsets[k]: info.seg_offsets[k] + info.seg_counts[k]]
            if info.compression == 5:
                raise UnsupportedTiff("LZW is decoded on the device (load_tiles(device_decode=True))")
            if info.compression != 1:
                data = zlib.decompress(data)
            r0 = s * info.rows_per_seg
            nr = min(info.rows_per_seg, H - r0)
            a = np.frombuffer(data, np.uint8, count=min(len(data), info.rows_per_seg * row_bytes))
            if a.size < nr * row_bytes:
                raise UnsupportedTiff("segment shorter than its rows")
            out[pl, r0:r0 + nr] = a[: nr * row_bytes].reshape(nr, row_bytes)[:, : W * spp * sb]
    return (out[0].reshape(H, W * C * sb) if info.planar == 1 else out.reshape(C, H, W * sb)), info


def segments(buf: bytes, info: TiffInfo):
    """(compressed bytes, decoded size) of every segment of the first image, in sample-buffer order, or None when the segments do
    not tile the sample buffer directly (internal tiles wider than the image: padded rows)."""
    if info.tiled or info.seg_width != info.width:
        return None
    H, W, C, sb = info.height, info.width, info.channels, info.sample_bytes
    spp = C if info.planar == 1 else 1
    planes = 1 if info.planar == 1 else C
    segs_per_plane = (H + info.rows_per_seg - 1) // info.rows_per_seg
    if len(info.seg_offsets) < planes * segs_per_plane:
        raise UnsupportedTiff("segment table shorter than the image")
    out = []
    for pl in range(planes):
        for s_ in range(segs_per_plane):
            k = pl * segs_per_plane + s_
            nr = min(info.rows_per_seg, H - s_ * info.rows_per_seg)
            out.append((buf[info.seg_offsets[k]: info.seg_offsets[k] + info.seg_counts[k]], nr * W * spp * sb))
    return out


def load_tiles(paths: Sequence, bidx: Optional[Sequence[int]] = None, rescale: Optional[dict] = None, engine=None,
               threads: int = 8, ids: Optional[Sequence] = None, device_decode: bool = True) -> TileBatch:
    """Read equally shaped GeoTIFF tiles into one TileBatch (pixels (T, H, W, C_out) on the host, transforms, nodata).
    bidx: 1-based input bands of the output bands, the tile server's ``bidx=`` list (default: all, in order).
    rescale: {'smin': [...], 'smax': [...], 'f32': bool} per OUTPUT band -- the gdal.Translate scaleParams of
    tif2cog.py:260-270 (dst = src * k + off, k = 255 / (smax - smin), off = -smin * k, clamped and rounded) -> uint8.
    device_decode: the compressed segments are uploaded as they are and decompressed on the GPU (rs_ingest_tiles_host);
    False (or a layout whose segments are padded) inflates on the host, a thread per file."""
    from .engine import default_engine
    eng = engine or default_engine()
    paths = list(paths)
    if not paths:
        raise ValueError("no tiles")

    def read(p):
        buf = p if isinstance(p, (bytes, bytearray)) else open(p, "rb").read()
        return buf, parse_tiff(buf)
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        files = list(ex.map(read, paths))
    info0 = files[0][1]
    for _, inf in files:
        if inf.layout != info0.layout:
            raise ValueError("tiles of one batch must share shape, sample width and layout")
    gt = np.array([inf.transform for _, inf in files], np.float64)
    comps = {inf.compression for _, inf in files}
    segs = [segments(buf, inf) for buf, inf in files] if (device_decode and len(comps) == 1) else None
    if segs is not None and all(s_ is not None for s_ in segs):
        flat = [x for s_ in segs for x in s_]
        comp_off = np.zeros(len(flat) + 1, np.int64)
        raw_off = np.zeros(len(flat) + 1, np.int64)
        comp_off[1:] = np.cumsum([len(c) for c, _ in flat])
        raw_off[1:] = np.cumsum([n for _, n in flat])
        comp = np.frombuffer(b"".join(c for c, _ in flat), np.uint8)
        pixels = eng.ingest_tiles_host(comp, comp_off, info0.compression, raw_off, len(files), info0, bidx, rescale)
    else:
        if 5 in comps:
            raise UnsupportedTiff("LZW tiles with padded internal tiles are not supported")
        with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:         # zlib releases the GIL
            raw = np.stack(list(ex.map(lambda f: read_raw(f[0])[0], files)))
        pixels = eng.assemble_tiles_host(raw, info0, bidx, rescale)
    return TileBatch(pixels, gt, info0.height, info0.width, pixels.shape[3], info0.nodata, list(ids) if ids is not None else paths)
