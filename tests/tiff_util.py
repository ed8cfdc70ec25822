"""Minimal baseline-TIFF writer for the ingest tests (strips or one internal tile, none / deflate, predictor 1 / 2,
chunky / planar, 8 / 16 bit, either byte order, GeoTIFF scale + tiepoint + GDAL_NODATA tags)."""
import struct
import zlib

import numpy as np


def lzw_encode(data: bytes) -> bytes:
    """TIFF LZW (MSB-first, ClearCode 256, EOI 257, code width grows one code early -- libtiff's convention)."""
    out = bytearray()
    acc = nb = 0

    def put(code, width):
        nonlocal acc, nb
        acc = (acc << width) | code
        nb += width
        while nb >= 8:
            out.append((acc >> (nb - 8)) & 255)
            nb -= 8
        acc &= (1 << nb) - 1

    table = {bytes([i]): i for i in range(256)}
    nxt, width = 258, 9
    put(256, width)
    w = b""
    for byte in data:
        wc = w + bytes([byte])
        if wc in table:
            w = wc
            continue
        put(table[w], width)
        table[wc] = nxt
        nxt += 1
        if nxt >= (1 << width) and width < 12:           # the decoder's table lags one entry: it switches at the same code
            width += 1
        if nxt >= 4094:                                  # table full: clear
            put(256, width)
            table = {bytes([i]): i for i in range(256)}
            nxt, width = 258, 9
        w = bytes([byte])
    if w:
        put(table[w], width)
        nxt += 1
        if nxt >= (1 << width) and width < 12:
            width += 1
    put(257, width)
    if nb:
        out.append((acc << (8 - nb)) & 255)
    return bytes(out)


def write_tiff(path, arr, compression=1, predictor=1, planar=1, big_endian=False, rows_per_strip=None, tile=None,
               transform=None, nodata=None, bigtiff=False):
    arr = np.asarray(arr)
    if arr.ndim == 2:
        arr = arr[..., None]
    H, W, C = arr.shape
    bits = arr.dtype.itemsize * 8
    e = ">" if big_endian else "<"
    enc = arr.astype(np.int64)
    if predictor == 2:
        d = enc.copy()
        d[:, 1:, :] = enc[:, 1:, :] - enc[:, :-1, :]
        enc = d % (1 << bits)
    enc = enc.astype(np.dtype(f"{e}u{bits // 8}"))
    planes = [enc] if planar == 1 else [enc[..., c:c + 1] for c in range(C)]
    segs = []
    if tile is not None:                       # one internal tile (tw >= W, th >= H), padded
        tw, th = tile
        for p in planes:
            pad = np.zeros((th, tw, p.shape[2]), p.dtype)
            pad[:H, :W] = p
            segs.append(pad.tobytes())
    else:
        rps = rows_per_strip or H
        for p in planes:
            for r0 in range(0, H, rps):
                segs.append(p[r0:r0 + rps].tobytes())
    if compression == 5:
        segs = [lzw_encode(s) for s in segs]
    elif compression != 1:
        segs = [zlib.compress(s) for s in segs]
    tags = []                                  # (tag, type, count, values)

    def add(tag, typ, vals):
        tags.append((tag, typ, vals))

    add(256, 4, [W]); add(257, 4, [H]); add(258, 3, [bits] * C); add(259, 3, [compression])
    add(262, 3, [2 if C >= 3 else 1]); add(277, 3, [C]); add(284, 3, [planar]); add(339, 3, [1] * C)
    if C > 3:
        add(338, 3, [0] * (C - 3))
    if predictor != 1:
        add(317, 3, [predictor])
    if tile is not None:
        add(322, 4, [tile[0]]); add(323, 4, [tile[1]]); add(324, 4, None); add(325, 4, [len(s) for s in segs])
    else:
        add(278, 4, [rows_per_strip or H]); add(273, 4, None); add(279, 4, [len(s) for s in segs])
    if transform is not None:
        a, b, c, d, ee, f = transform
        add(33550, 12, [a, -ee, 0.0]); add(33922, 12, [0.0, 0.0, 0.0, c, f, 0.0])
    if nodata is not None:
        add(42113, 2, str(nodata).encode() + b"\x00")
    tags.sort(key=lambda t: t[0])
    fmt = {2: "c", 3: "H", 4: "I", 12: "d"}
    size = {2: 1, 3: 2, 4: 4, 12: 8}
    # classic TIFF: 4-byte offsets, 12-byte entries; BigTIFF: 8-byte offsets, 20-byte entries, 8-byte counts
    osz, inline = (8, 8) if bigtiff else (4, 4)
    ofmt = "Q" if bigtiff else "I"
    ifd_off = 16 if bigtiff else 8
    ifd_len = (8 if bigtiff else 2) + (20 if bigtiff else 12) * len(tags) + osz
    extra_off = ifd_off + ifd_len
    extra = b""
    # segment data goes after the extra block; compute the extra block size first
    def payload(typ, vals):
        if typ == 2:
            return vals
        return struct.pack(e + fmt[typ] * len(vals), *vals)
    sizes = []
    for tag, typ, vals in tags:
        n = len(segs) if vals is None else len(vals)
        nb = n * size[typ]
        sizes.append(nb if nb > inline else 0)
    data_off = extra_off + sum((s + 1) // 2 * 2 for s in sizes)
    seg_offs, o = [], data_off
    for s in segs:
        seg_offs.append(o)
        o += len(s)
    entries = b""
    for (tag, typ, vals), nb in zip(tags, sizes):
        if vals is None:
            vals = seg_offs
        pl = payload(typ, vals)
        cnt = len(vals)
        if nb == 0:
            entries += struct.pack(e + "HH" + ofmt, tag, typ, cnt) + pl.ljust(inline, b"\x00")
        else:
            entries += struct.pack(e + "HH" + ofmt + ofmt, tag, typ, cnt, extra_off + len(extra))
            extra += pl + (b"\x00" if len(pl) % 2 else b"")
    if bigtiff:
        head = (b"MM" if big_endian else b"II") + struct.pack(e + "HHHQ", 43, 8, 0, ifd_off)
        body = struct.pack(e + "Q", len(tags)) + entries + struct.pack(e + "Q", 0) + extra + b"".join(segs)
    else:
        head = (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, ifd_off)
        body = struct.pack(e + "H", len(tags)) + entries + struct.pack(e + "I", 0) + extra + b"".join(segs)
    with open(path, "wb") as fh:
        fh.write(head + body)
