"""Tile ingest (proj_roadsurf_b200/ingest.py): the TIFF directory parser and segment reader on the CPU, the oracle's
restatement of the post-decompression steps, PIL (libtiff) as the third-party decoder where it can read the file, and the
assemble kernel on the GPU."""
import itertools
import os

import numpy as np
import pytest

from oracle import raster as oraster
from proj_roadsurf_b200 import ingest
from tiff_util import write_tiff

T = (0.5971642834779395, 0.0, 815000.25, 0.0, -0.5971642834779395, 5935000.75)
VARIANTS = [dict(compression=c, predictor=p, planar=pl, big_endian=be, rows_per_strip=rps, tile=tl)
            for c, p, pl, be, (rps, tl) in itertools.product((1, 8), (1, 2), (1, 2), (False, True), ((None, None), (7, None), (None, (64, 48))))
            if not (c == 1 and p == 2)]          # libtiff ignores the predictor tag of uncompressed data (see below)


def _image(rng, C, dtype, H=40, W=52):
    hi = 256 if dtype == np.uint8 else 65536
    a = rng.integers(0, hi, (H, W, C)).astype(dtype)
    a[5:9, :, :] = hi - 1                      # runs of equal and extreme values: differences wrap
    a[9:12, ::2, :] = 0
    return a


def _decode(path, bidx=None):
    raw, info = ingest.read_raw(path)
    return oraster.tiff_assemble(raw, info.height, info.width, info.channels, info.sample_bytes, info.planar, info.predictor,
                                 info.big_endian, bidx), info


@pytest.mark.parametrize("C,dtype", [(1, np.uint8), (3, np.uint8), (4, np.uint8), (1, np.uint16), (4, np.uint16)])
def test_reader_and_oracle_round_trip_every_layout(tmp_path, C, dtype):
    rng = np.random.default_rng(C * 7 + np.dtype(dtype).itemsize)
    img = _image(rng, C, dtype)
    for i, v in enumerate(VARIANTS):
        p = str(tmp_path / f"t{i}.tif")
        write_tiff(p, img, transform=T, nodata=0, **v)
        got, info = _decode(p)
        assert np.array_equal(got, img), v
        assert info.transform == T and info.nodata == 0
    got, _ = _decode(p, bidx=[C, 1])
    assert np.array_equal(got, img[..., [C - 1, 0]])


def test_pil_libtiff_agrees_where_it_can_read(tmp_path):
    """PIL decodes through libtiff, the decoder under GDAL / rasterio: files written here are read identically by it, and
    files written by PIL are read identically here"""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    rgb, gray16 = _image(rng, 3, np.uint8), _image(rng, 1, np.uint16)[..., 0]
    n = 0
    for i, v in enumerate(VARIANTS):
        p = str(tmp_path / f"w{i}.tif")
        write_tiff(p, rgb, **v)
        try:
            with Image.open(p) as im:
                pil = np.asarray(im)
        except Exception:  # noqa: BLE001  (a layout this PIL build does not read)
            continue
        assert np.array_equal(pil, rgb), v
        n += 1
    assert n >= 8
    # uncompressed + predictor tag: libtiff (and this reader) ignore the tag; the samples are taken as stored
    p = str(tmp_path / "nopred.tif")
    write_tiff(p, rgb, compression=1, predictor=2)
    raw, info = ingest.read_raw(p)
    assert info.predictor == 1
    with Image.open(p) as im:
        assert np.array_equal(np.asarray(im), _decode(p)[0])
    for comp in ("raw", "tiff_adobe_deflate", "tiff_deflate"):
        p = str(tmp_path / f"pil_{comp}.tif")
        Image.fromarray(rgb).save(p, compression=comp)
        assert np.array_equal(_decode(p)[0], rgb), comp
        p16 = str(tmp_path / f"pil16_{comp}.tif")
        Image.fromarray(gray16).save(p16, compression=comp)
        assert np.array_equal(_decode(p16)[0][..., 0], gray16), comp


def test_unsupported_files_are_refused(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    p = str(tmp_path / "lzw.tif")
    Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(p, compression="tiff_lzw")
    with pytest.raises(ingest.UnsupportedTiff):
        ingest.read_raw(p)                                          # LZW has no host decoder here: it is decoded on the device
    assert ingest.parse_tiff(open(p, "rb").read()).compression == 5
    p = str(tmp_path / "jpeg.tif")
    Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(p, compression="jpeg")
    with pytest.raises(ingest.UnsupportedTiff):
        ingest.parse_tiff(open(p, "rb").read())
    with pytest.raises(ingest.UnsupportedTiff):
        ingest.parse_tiff(b"not a tiff at all")


def test_bigtiff_directories_parse_like_classic_ones(tmp_path):
    rng = np.random.default_rng(4)
    img = _image(rng, 3, np.uint8)
    for i, v in enumerate(VARIANTS[::3]):
        pc, pb = str(tmp_path / f"c{i}.tif"), str(tmp_path / f"b{i}.tif")
        write_tiff(pc, img, transform=T, nodata=0, **v)
        write_tiff(pb, img, transform=T, nodata=0, bigtiff=True, **v)
        assert open(pb, "rb").read()[2:4] in (b"+\x00", b"\x00+")
        gc, ic = _decode(pc)
        gb, ib = _decode(pb)
        assert np.array_equal(gc, gb) and np.array_equal(gb, img)
        assert ic.layout == ib.layout and ib.transform == T and ib.nodata == 0 and ic.seg_counts == ib.seg_counts


@pytest.mark.gpu
def test_device_decode_deflate_lzw_bigtiff(tmp_path):
    """the compressed strips are decoded ON THE DEVICE (rs_ingest_tiles_host): deflate and LZW, predictor 1 / 2, chunky / planar,
    both byte orders, classic and BigTIFF, files written here and by libtiff (PIL); equal to the host-inflate path and to the
    images; a corrupt strip is an error, never garbage"""
    from proj_roadsurf_b200._native import NativeError
    rng = np.random.default_rng(17)
    n_checked = 0
    for comp, pred, planar, be, rps, big in itertools.product((8, 5), (1, 2), (1, 2), (False, True), (5, None), (False, True)):
        if be and big:
            continue
        imgs, paths = [], []
        for k in range(4):
            img = _image(rng, 4, np.uint16 if (pred == 2 and planar == 2) else np.uint8, H=37, W=48)
            p = str(tmp_path / f"d{n_checked}_{k}.tif")
            write_tiff(p, img, compression=comp, predictor=pred, planar=planar, big_endian=be, rows_per_strip=rps, bigtiff=big,
                       transform=T, nodata=0)
            imgs.append(img); paths.append(p)
        tb = ingest.load_tiles(paths, device_decode=True)
        assert np.array_equal(tb.pixels, np.stack(imgs)), (comp, pred, planar, be, rps, big)
        if comp == 8:
            assert np.array_equal(ingest.load_tiles(paths, device_decode=False).pixels, tb.pixels)
        n_checked += 1
    assert n_checked == 48
    Image = pytest.importorskip("PIL.Image")
    rgb = np.clip(rng.normal(110, 6, (3, 256, 256, 3)), 0, 255).astype(np.uint8)            # three "asphalt" tiles
    for comp in ("tiff_lzw", "tiff_adobe_deflate", "raw"):
        paths = []
        for k in range(3):
            p = str(tmp_path / f"pil_{comp}_{k}.tif")
            Image.fromarray(rgb[k]).save(p, compression=comp)
            paths.append(p)
        assert np.array_equal(ingest.load_tiles(paths).pixels, rgb), comp
    # a damaged strip: RS_ERR_CODEC
    buf = bytearray(open(paths[0], "rb").read())
    p = str(tmp_path / "bad.tif")
    write_tiff(p, rgb[0], compression=8, rows_per_strip=16)
    buf = bytearray(open(p, "rb").read())
    info = ingest.parse_tiff(bytes(buf))
    buf[info.seg_offsets[3] + 10] ^= 0x55
    open(p, "wb").write(bytes(buf))
    with pytest.raises(NativeError) as ei:
        ingest.load_tiles([p])
    assert ei.value.status == -10


@pytest.mark.gpu
@pytest.mark.parametrize("C,dtype", [(3, np.uint8), (4, np.uint8), (1, np.uint16), (4, np.uint16)])
def test_load_tiles_on_the_gpu(tmp_path, C, dtype):
    rng = np.random.default_rng(11 + C)
    for vi, v in enumerate(VARIANTS[::5]):
        imgs, paths = [], []
        for k in range(3):
            img = _image(rng, C, dtype, H=33, W=61)
            p = str(tmp_path / f"g{vi}_{k}.tif")
            write_tiff(p, img, transform=(T[0], 0.0, T[2] + 10.0 * k, 0.0, T[4], T[5]), nodata=0, **v)
            imgs.append(img); paths.append(p)
        tb = ingest.load_tiles(paths)
        assert tb.pixels.dtype == dtype and np.array_equal(tb.pixels, np.stack(imgs)), v
        assert np.array_equal(tb.gt[:, 2], T[2] + 10.0 * np.arange(3)) and tb.nodata == 0
        if C == 4:
            bidx = [2, 3, 4, 1]                                     # config_stats.yaml:39
            assert np.array_equal(ingest.load_tiles(paths, bidx=bidx).pixels, np.stack(imgs)[..., [1, 2, 3, 0]])
            if dtype == np.uint16:                                  # tif2cog.py:260-270 on ingest
                smin, smax = [100.0, 100.0, 100.0, 300.0], [20000.0, 20000.0, 20000.0, 40000.0]
                for f32 in (False, True):
                    got = ingest.load_tiles(paths, bidx=bidx, rescale={"smin": smin, "smax": smax, "f32": f32}).pixels
                    exp = oraster.rescale_u16_to_u8(np.stack(imgs)[..., [1, 2, 3, 0]], smin, smax, f32)
                    assert got.dtype == np.uint8 and np.array_equal(got, exp), (v, f32)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lut", "warp", "bits"])
def test_device_inflate_matches_zlib(mode, monkeypatch):
    """rs_decode_segments_host on zlib streams of every compression level and strategy (stored, fixed and dynamic blocks, maximal
    matches at distance 1, window-length distances, tiny alphabets, segments far larger than a strip) in ONE batch: byte for
    byte what zlib gives -- the three decoders (tables, a thread per segment; a warp per segment; bit by bit); damaged
    streams are reported (RS_ERR_CODEC), the others of the batch untouched by them"""
    import zlib
    from test_codec_host import _payloads
    from proj_roadsurf_b200._native import NativeError
    from proj_roadsurf_b200.engine import Engine
    monkeypatch.setenv("RS_INFLATE", mode)
    raws, comps = [], []
    for raw in _payloads():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                c = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
                comps.append(c.compress(raw) + c.flush())
                raws.append(raw)
    rng = np.random.default_rng(23)
    for k in range(300):                                                   # many small strips of varied content and alignment
        n = int(rng.integers(1, 7000))
        kind = k % 3
        raw = (rng.integers(0, 256, n, dtype=np.uint8) if kind == 0 else np.clip(rng.normal(110, 6, n), 0, 255).astype(np.uint8)
               if kind == 1 else (np.arange(n) // int(rng.integers(1, 40)) % 7).astype(np.uint8)).tobytes()
        comps.append(zlib.compress(raw, int(rng.integers(0, 10))))
        raws.append(raw)
    comp = np.frombuffer(b"".join(comps), np.uint8)
    comp_off = np.concatenate([[0], np.cumsum([len(c) for c in comps])]).astype(np.int64)
    raw_off = np.concatenate([[0], np.cumsum([len(r) for r in raws])]).astype(np.int64)
    eng = Engine(0)
    out = eng.decode_segments_host(comp, comp_off, 8, raw_off)
    want = np.frombuffer(b"".join(raws), np.uint8)
    for i in range(len(raws)):
        assert np.array_equal(out[raw_off[i]:raw_off[i + 1]], want[raw_off[i]:raw_off[i + 1]]), (i, len(raws[i]))
    # one damaged stream in the batch
    bad = comp.copy()
    victim = len(comps) - 7
    bad[comp_off[victim] + len(comps[victim]) // 2] ^= 0x10
    with pytest.raises(NativeError) as ei:
        eng.decode_segments_host(bad, comp_off, 8, raw_off)
    assert ei.value.status == -10
    eng.close()
