"""The decoder cores of the device-side ingest (proj_roadsurf_b200/csrc/rs_codec_core.h) compiled for the HOST with g++ and held
to zlib (every compression level and strategy: stored, fixed and dynamic Huffman blocks, long matches, window-length distances)
and to libtiff-made LZW streams (PIL) -- the same source the GPU kernels compile, so the bit-level logic is checked without a
GPU; tests/test_ingest.py checks the kernels themselves with -m gpu."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcodec_host.so")
    src = os.path.join(HERE, "native", "codec_host.cpp")
    hdr = os.path.join(ROOT, "proj_roadsurf_b200", "csrc", "rs_codec_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.dirname(hdr), "-o", so, src])
    L = ctypes.CDLL(so)
    for f in (L.host_inflate, L.host_inflate_warp, L.host_inflate_lut, L.host_lzw):
        f.restype = ctypes.c_longlong
        f.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
    return L


def test_long_segments_sum_their_output_again(tmp_path):
    """segments longer than RS_T_INLINE_SUM_MAX bytes (2^28 in the product) leave the unreduced inline Adler-32 sums and sum
    their output again in the trailer: the same source built with the limit at 1 000 bytes"""
    so = str(tmp_path / "libcodec_small.so")
    src = os.path.join(HERE, "native", "codec_host.cpp")
    hdr = os.path.join(ROOT, "proj_roadsurf_b200", "csrc", "rs_codec_core.h")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DRS_T_INLINE_SUM_MAX=1000", "-I", os.path.dirname(hdr),
                           "-o", so, src])
    L = ctypes.CDLL(so)
    L.host_inflate_lut.restype = ctypes.c_longlong
    L.host_inflate_lut.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
    rng = np.random.default_rng(3)
    for n in (10, 999, 1000, 1001, 70000):
        raw = np.clip(rng.normal(110, 6, n), 0, 255).astype(np.uint8).tobytes()
        comp = bytearray(zlib.compress(raw, 6))
        got, out = _run(L.host_inflate_lut, bytes(comp), n)
        assert got == n and out == raw, n
        comp[-1] ^= 0x01                                                   # a wrong checksum is still caught
        assert _run(L.host_inflate_lut, bytes(comp), n)[0] == -1, n


def _run(fn, comp: bytes, cap: int):
    out = np.zeros(max(cap, 1), np.uint8)
    got = fn(comp, len(comp), out.ctypes.data, cap)
    return got, out[:max(got, 0)].tobytes()


def _payloads():
    rng = np.random.default_rng(5)
    yield b""
    yield b"a"
    yield bytes(70000)                                                     # one long run: maximal matches, distance 1
    yield rng.integers(0, 256, 50000, dtype=np.uint8).tobytes()            # incompressible: stored blocks at level 0 / 1
    yield np.clip(rng.normal(110, 6, (256, 256, 3)), 0, 255).astype(np.uint8).tobytes()      # "asphalt" tile
    yield (np.arange(200000) % 251).astype(np.uint8).tobytes()             # period 251: long-distance matches
    yield rng.integers(0, 4, 300000, dtype=np.uint8).tobytes()             # tiny alphabet: short codes, 32 KiB window in use
    img = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    d = img.astype(np.int16)
    d[:, 1:] -= img[:, :-1]
    yield (d % 256).astype(np.uint8).tobytes()                             # predictor-2 residuals


INFLATERS = ("host_inflate", "host_inflate_warp", "host_inflate_lut")      # bit by bit / a warp per segment (one lane here) / tables


@pytest.mark.parametrize("which", INFLATERS)
def test_inflate_matches_zlib(lib, which):
    inflate = getattr(lib, which)
    n = 0
    for raw in _payloads():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                c = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
                comp = c.compress(raw) + c.flush()
                got, out = _run(inflate, comp, len(raw))
                assert got == len(raw) and out == raw, (len(raw), level, strategy)
                n += 1
    assert n == 8 * 16


@pytest.mark.parametrize("which", INFLATERS)
def test_inflate_rejects_bad_streams(lib, which):
    inflate = getattr(lib, which)
    raw = bytes(range(256)) * 40
    comp = zlib.compress(raw)
    assert _run(inflate, comp, len(raw) - 1)[0] == -1             # output larger than the segment may hold
    assert _run(inflate, comp[: len(comp) // 2], len(raw))[0] == -1        # truncated
    assert _run(inflate, b"\x78\x9d" + comp[2:], len(raw))[0] == -1        # header check fails
    bad = bytearray(comp)
    bad[2] |= 0x06                                                         # block type 3
    assert _run(inflate, bytes(bad), len(raw))[0] == -1
    assert _run(inflate, b"", 10)[0] == -1


def test_lzw_matches_libtiff(lib, tmp_path):
    from tiff_util import lzw_encode
    rng = np.random.default_rng(9)
    for raw in _payloads():
        if not raw:
            continue
        got, out = _run(lib.host_lzw, lzw_encode(raw), len(raw))
        assert got == len(raw) and out == raw, len(raw)
    Image = pytest.importorskip("PIL.Image")
    from proj_roadsurf_b200 import ingest
    for k, img in enumerate((rng.integers(0, 256, (300, 400, 3), dtype=np.uint8), np.full((90, 130), 77, np.uint8),
                             rng.integers(0, 7, (513, 257), dtype=np.uint8))):
        p = str(tmp_path / f"pil{k}.tif")
        Image.fromarray(img).save(p, compression="tiff_lzw")               # libtiff's encoder
        buf = open(p, "rb").read()
        info = ingest.parse_tiff(buf)
        assert info.compression == 5
        dec = b""
        for comp, n in ingest.segments(buf, info):
            got, out = _run(lib.host_lzw, comp, n)
            assert got == n
            dec += out
        assert dec == img.tobytes()
    assert _run(lib.host_lzw, lzw_encode(bytes(1000)), 999)[0] == -1       # does not fit


@pytest.mark.parametrize("which", INFLATERS)
def test_inflate_checks_the_adler32_trailer(lib, which):
    inflate = getattr(lib, which)
    raw = np.random.default_rng(1).integers(0, 256, 20000, dtype=np.uint8).tobytes()
    comp = bytearray(zlib.compress(raw, 0))                                # stored blocks: a flipped payload byte still "decodes"
    comp[100] ^= 0x01
    assert _run(inflate, bytes(comp), len(raw))[0] == -1
    comp[100] ^= 0x01
    assert _run(inflate, bytes(comp), len(raw))[0] == len(raw)
    assert _run(inflate, bytes(comp[:-1]), len(raw))[0] == -1     # trailer cut short


def test_corrupt_streams_never_leave_their_buffers(tmp_path):
    """tests/native/codec_fuzz.cpp under AddressSanitizer + UBSan: the intact streams decode, ~20 000 corrupted ones (bit flips,
    truncation, wrong capacities) are decoded into exact-size heap buffers without a single out-of-bounds access"""
    import struct
    from tiff_util import lzw_encode
    exe = os.path.join(HERE, "_build", "codec_fuzz")
    src = os.path.join(HERE, "native", "codec_fuzz.cpp")
    hdr = os.path.join(ROOT, "proj_roadsurf_b200", "csrc", "rs_codec_core.h")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        try:
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                                   "-I", os.path.dirname(hdr), "-o", exe, src])
        except (subprocess.CalledProcessError, FileNotFoundError):
            pytest.skip("no sanitizer runtime for g++ here")
    blob = bytearray()
    n_streams = 0
    for raw in _payloads():
        if not raw:
            continue
        raw = raw[:40000]
        for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)):
            c = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
            comp = c.compress(raw) + c.flush()
            blob += struct.pack("<BII", 0, len(raw), len(comp)) + comp
            n_streams += 1
        comp = lzw_encode(raw)
        blob += struct.pack("<BII", 1, len(raw), len(comp)) + comp
        n_streams += 1
    path = tmp_path / "streams.bin"
    path.write_bytes(bytes(blob))
    rounds = 600
    res = subprocess.run([exe, str(path), str(rounds)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.returncode, res.stderr[-2000:])
    intact, mutated, accepted = map(int, res.stdout.split())
    assert intact == n_streams and mutated == n_streams * rounds
    assert accepted < mutated                                             # most corruptions are noticed (not all can be: LZW has no checksum)


def test_inflate_random_streams_property(lib):
    """hypothesis: byte strings of mixed structure (runs, repeated phrases, noise, tiny alphabets) at a random level, strategy and
    memLevel -- which moves the block boundaries and the code lengths -- decode to themselves through all three decoders, and a
    capacity one byte short is refused"""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    piece = st.one_of(
        st.binary(min_size=0, max_size=300),                                                  # noise
        st.builds(lambda b, n: bytes([b]) * n, st.integers(0, 255), st.integers(1, 2000)),    # runs: distance-1 matches
        st.builds(lambda w, n: w * n, st.binary(min_size=1, max_size=40), st.integers(1, 60)),    # phrases: short distances
        st.builds(lambda n, k, seed: bytes(np.random.default_rng(seed).integers(0, k, n, dtype=np.uint8)),
                  st.integers(1, 3000), st.integers(1, 6), st.integers(0, 2 ** 31)))          # tiny alphabets: short codes

    @hyp.settings(max_examples=120, deadline=None)
    @hyp.given(st.lists(piece, min_size=0, max_size=12), st.integers(0, 9),
               st.sampled_from([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED]), st.integers(1, 9))
    def check(pieces, level, strategy, mem):
        raw = b"".join(pieces)
        c = zlib.compressobj(level, zlib.DEFLATED, 15, mem, strategy)
        comp = c.compress(raw) + c.flush()
        for which in INFLATERS:
            got, out = _run(getattr(lib, which), comp, len(raw))
            assert got == len(raw) and out == raw, (which, len(raw), level, strategy, mem)
            if raw:
                assert _run(getattr(lib, which), comp, len(raw) - 1)[0] == -1, which

    check()
