// Host build of the decoder cores of proj_roadsurf_b200/csrc/rs_codec_core.h (the source the GPU kernels compile): lets the CPU
// test suite hold the DEFLATE and TIFF-LZW decoders to zlib- and libtiff-made streams without a GPU (tests/test_codec_host.py).
#include <stdint.h>
#include <stdlib.h>

#include "rs_codec_core.h"

extern "C" long long host_inflate(const uint8_t *src, long long n, uint8_t *dst, long long cap)
{
    static thread_local uint16_t tab16[rs::codec::RS_INFLATE_U16];
    static thread_local uint8_t lens[rs::codec::RS_INFLATE_U8];
    return rs::codec::inflate_segment(src, n, dst, cap, true, tab16, lens, 1);
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
