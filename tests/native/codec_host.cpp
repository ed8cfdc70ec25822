// Host build of the decoder cores of proj_roadsurf_b200/csrc/rs_codec_core.h (the source the GPU kernels compile): lets the CPU
// test suite hold the DEFLATE and TIFF-LZW decoders to zlib- and libtiff-made streams without a GPU (tests/test_codec_host.py).
#include <stdint.h>
#include <stdlib.h>

#include "rs_codec_core.h"

extern "C" long long host_inflate(const uint8_t *src, long long n, uint8_t *dst, long long cap)
{
    uint16_t hot[rs::codec::RS_INFLATE_HOT], sym[rs::codec::RS_INFLATE_SYM];
    uint8_t lens[rs::codec::RS_INFLATE_LEN];
    return rs::codec::inflate_segment(src, n, dst, cap, true, hot, 1, sym, lens);
}

// the warp-per-segment decoder, compiled with one lane
extern "C" long long host_inflate_warp(const uint8_t *src, long long n, uint8_t *dst, long long cap)
{
    rs::codec::WTables *t = (rs::codec::WTables *)malloc(sizeof(rs::codec::WTables));
    const long long got = rs::codec::inflate_segment_warp(src, n, dst, cap, true, *t, 0);
    free(t);
    return got;
}

// the table-driven thread-per-segment decoder
extern "C" long long host_inflate_lut(const uint8_t *src, long long n, uint8_t *dst, long long cap)
{
    uint16_t tab[rs::codec::RS_T_SMEM], sym[rs::codec::RS_INFLATE_SYM];
    uint8_t lens[rs::codec::RS_INFLATE_LEN];
    return rs::codec::inflate_segment_lut(src, n, dst, cap, true, tab, 1, sym, lens);
}

extern "C" long long host_lzw(const uint8_t *src, long long n, uint8_t *dst, long long cap)
{
    uint32_t *tab = (uint32_t *)malloc(4096 * sizeof(uint32_t));
    uint16_t *len = (uint16_t *)malloc(4096 * sizeof(uint16_t));
    const long long got = rs::codec::lzw_segment(src, n, dst, cap, tab, len);
    free(tab);
    free(len);
    return got;
}
