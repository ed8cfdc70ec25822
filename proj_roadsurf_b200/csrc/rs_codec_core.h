// Decoder cores of the device-side tile ingest (rs_codec.cu): zlib / DEFLATE (RFC 1950 / 1951) and TIFF LZW, written so that the
// same source compiles for the device (nvcc) and for the host (g++, tests/test_codec_host.py runs it on the CPU against zlib and
// libtiff-made streams, which needs no GPU).  See rs_codec.cu for the design.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define RS_HD __device__
#define RS_TABLE __device__ __constant__ const
#else
#define RS_HD
#define RS_TABLE static const
#endif

namespace rs {
namespace codec {

// ---------------------------------------------------------------------------------------------
// DEFLATE
// ---------------------------------------------------------------------------------------------
RS_TABLE uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
RS_TABLE uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
RS_TABLE uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
RS_TABLE uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
RS_TABLE uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Bits {                         // LSB-first bit reader over [p, p + n)
    const uint8_t *p;
    long long n, pos;
    uint32_t buf;
    int cnt;
    bool over;                        // read past the end of the input
    RS_HD inline uint32_t get(int k)
    {
        while (cnt < k) {
            uint32_t b = 0;
            if (pos < n) b = p[pos];
            else over = true;
            pos++;
            buf |= b << cnt;
            cnt += 8;
        }
        const uint32_t v = buf & ((1u << k) - 1u);
        buf >>= k;
        cnt -= k;
        return k ? v : 0u;
    }
};

// A decoder's tables live where its caller puts them.  The per-length COUNTS are read for every bit of every code: element i is
// at count[i * stride] -- on the device a column of the block's shared memory (stride = block size: thread t owns column t), on
// the host (tests) stride 1.  The symbol lists and the code lengths are touched once per code: plain per-decoder arrays.
struct Huff {                         // canonical code: symbols ordered by (length, value), count of codes per length
    uint16_t *count;                  // [16], strided
    uint16_t *symbol;                 // [288] (literal / length) or [30] (distance)
    int stride;
};
enum { RS_INFLATE_HOT = 48, RS_INFLATE_SYM = 288 + 30, RS_INFLATE_LEN = 320 };   // strided uint16 / plain uint16 / plain uint8 per decoder

// lengths[0 .. n) -> canonical code; returns false for an over-subscribed set (an incomplete set is accepted, as zlib accepts
// the single-code distance trees real encoders write).  offs: 16 uint16 of scratch (same stride)
RS_HD bool huff_build(Huff &h, const uint8_t *len, int n, uint16_t *offs)
{
    const int st = h.stride;
    for (int i = 0; i < 16; i++) h.count[i * st] = 0;
    for (int i = 0; i < n; i++) h.count[len[i] * st]++;
    int left = 1;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= h.count[l * st];
        if (left < 0) return false;
    }
    offs[1 * st] = 0;
    for (int l = 1; l < 15; l++) offs[(l + 1) * st] = offs[l * st] + h.count[l * st];
    for (int i = 0; i < n; i++) {
        const int l = len[i];
        if (l) h.symbol[offs[l * st]++] = (uint16_t)i;
    }
    return true;
}

RS_HD int huff_decode(Bits &b, const Huff &h)
{
    int code = 0, first = 0, index = 0;
    const int st = h.stride;
    for (int l = 1; l < 16; l++) {
        code |= (int)b.get(1);
        const int count = h.count[l * st];
        if (code - count < first) return h.symbol[index + (code - first)];
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// returns bytes written, or -1 (corrupt / unsupported stream, or output larger than cap)
// hot: RS_INFLATE_HOT uint16 elements, element i at hot[i * stride]; sym: RS_INFLATE_SYM uint16; lens: RS_INFLATE_LEN uint8
RS_HD long long inflate_segment(const uint8_t *src, long long n, uint8_t *dst, long long cap, bool zlib_wrapper, uint16_t *hot,
                                int stride, uint16_t *sym, uint8_t *lens)
{
    Bits b{src, n, 0, 0u, 0, false};
    if (zlib_wrapper) {
        if (n < 2) return -1;
        const uint32_t cmf = b.get(8), flg = b.get(8);
        if ((cmf & 15u) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) return -1;      // deflate, header check, no preset dictionary
    }
    long long out = 0;
    Huff lit{hot, sym, stride}, dist{hot + 16 * stride, sym + 288, stride};
    uint16_t *offs = hot + 32 * stride;
    for (;;) {
        const uint32_t last = b.get(1), type = b.get(2);
        if (type == 0) {
            b.buf = 0; b.cnt = 0;                                  // to the next byte boundary
            if (b.pos + 4 > n) return -1;
            const uint32_t len = src[b.pos] | (src[b.pos + 1] << 8), nlen = src[b.pos + 2] | (src[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ 0xffffu) != nlen || b.pos + len > n || out + len > cap) return -1;
            for (uint32_t i = 0; i < len; i++) dst[out + i] = src[b.pos + i];
            out += len;
            b.pos += len;
        } else if (type == 1 || type == 2) {
            if (type == 1) {
                for (int i = 0; i < 144; i++) lens[i] = 8;
                for (int i = 144; i < 256; i++) lens[i] = 9;
                for (int i = 256; i < 280; i++) lens[i] = 7;
                for (int i = 280; i < 288; i++) lens[i] = 8;
                huff_build(lit, lens, 288, offs);
                for (int i = 0; i < 30; i++) lens[i] = 5;
                huff_build(dist, lens, 30, offs);
            } else {
                const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
                if (nlen > 286 || ndist > 30) return -1;
                for (int i = 0; i < 19; i++) lens[i] = 0;
                for (int i = 0; i < ncode; i++) lens[CL_ORDER[i]] = (uint8_t)b.get(3);
                if (!huff_build(lit, lens, 19, offs)) return -1;          // the code-length code, kept in `lit` for a moment
                int idx = 0;
                while (idx < nlen + ndist) {
                    const int sy = huff_decode(b, lit);
                    if (sy < 0) return -1;
                    if (sy < 16) lens[idx++] = (uint8_t)sy;
                    else {
                        int rep, val = 0;
                        if (sy == 16) {
                            if (idx == 0) return -1;
                            val = lens[idx - 1];
                            rep = 3 + (int)b.get(2);
                        } else if (sy == 17) rep = 3 + (int)b.get(3);
                        else rep = 11 + (int)b.get(7);
                        if (idx + rep > nlen + ndist) return -1;
                        while (rep--) lens[idx++] = (uint8_t)val;
                    }
                }
                if (lens[256] == 0) return -1;                      // no end-of-block code
                // the distance lengths first: building the literal code overwrites nothing they need (separate arrays)
                if (!huff_build(dist, lens + nlen, ndist, offs)) return -1;
                if (!huff_build(lit, lens, nlen, offs)) return -1;
            }
            for (;;) {
                const int sy = huff_decode(b, lit);
                if (sy < 0 || b.over) return -1;
                if (sy < 256) {
                    if (out >= cap) return -1;
                    dst[out++] = (uint8_t)sy;
                } else if (sy == 256) break;
                else {
                    const int ls = sy - 257;
                    if (ls >= 29) return -1;
                    const int len = LEN_BASE[ls] + (int)b.get(LEN_EXTRA[ls]);
                    const int ds = huff_decode(b, dist);
                    if (ds < 0 || ds >= 30) return -1;
                    const long long d = DIST_BASE[ds] + (long long)b.get(DIST_EXTRA[ds]);
                    if (d > out || out + len > cap) return -1;
                    for (int i = 0; i < len; i++, out++) dst[out] = dst[out - d];
                }
            }
        } else
            return -1;
        if (b.over) return -1;
        if (last) break;
    }
    if (zlib_wrapper) {
        // Adler-32 of the output against the stream's trailer (big-endian, at the next byte boundary): a damaged strip is an
        // error, never silently wrong pixels
        const long long tp = b.pos - b.cnt / 8;
        if (tp + 4 > n) return -1;
        const uint32_t want = ((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) | ((uint32_t)src[tp + 2] << 8) | (uint32_t)src[tp + 3];
        uint32_t s1 = 1, s2 = 0;
        for (long long i = 0; i < out;) {
            const long long stop = i + 5552 < out ? i + 5552 : out;      // the largest run that cannot overflow 32 bits
            for (; i < stop; i++) { s1 += dst[i]; s2 += s1; }
            s1 %= 65521u;
            s2 %= 65521u;
        }
        if (((s2 << 16) | s1) != want) return -1;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// TIFF LZW
// ---------------------------------------------------------------------------------------------
struct BitsMsb {
    const uint8_t *p;
    long long n, pos;
    uint32_t buf;
    int cnt;
    RS_HD inline int get(int k)            // -1 at the end of the input
    {
        while (cnt < k) {
            if (pos >= n) return -1;
            buf = (buf << 8) | p[pos++];
            cnt += 8;
        }
        cnt -= k;
        return (int)((buf >> cnt) & ((1u << k) - 1u));
    }
};

// table entry: prefix code (12 bits) | last byte << 12 | first byte << 20; length in a second array
RS_HD long long lzw_segment(const uint8_t *src, long long n, uint8_t *dst, long long cap, uint32_t *tab, uint16_t *tlen)
{
    BitsMsb b{src, n, 0, 0u, 0};
    for (int i = 0; i < 256; i++) { tab[i] = 0xfffu | ((uint32_t)i << 12) | ((uint32_t)i << 20); tlen[i] = 1; }
    int nbits = 9, next = 258, old = -1;
    long long out = 0;
    for (;;) {
        int code = b.get(nbits);
        if (code < 0 || code == 257) break;                // end of data / EOI
        if (code == 256) {
            nbits = 9; next = 258;
            code = b.get(9);
            if (code < 0 || code == 257) break;
            if (code > 255 || out >= cap) return -1;
            dst[out++] = (uint8_t)code;
            old = code;
            continue;
        }
        if (old < 0) {                                     // streams that do not open with a ClearCode
            if (code > 255 || out >= cap) return -1;
            dst[out++] = (uint8_t)code;
            old = code;
            continue;
        }
        uint32_t first;
        int len;
        if (code < next) {
            if (code > 257 || code < 256) { first = (tab[code] >> 20) & 255u; len = tlen[code]; }
            else return -1;
            if (out + len > cap) return -1;
            int c = code;
            for (int k = len - 1; k >= 0; k--) { dst[out + k] = (uint8_t)((tab[c] >> 12) & 255u); c = (int)(tab[c] & 0xfffu); }
        } else if (code == next) {                         // the string being defined: old + its own first byte
            first = (tab[old] >> 20) & 255u;
            len = tlen[old] + 1;
            if (out + len > cap) return -1;
            int c = old;
            for (int k = len - 2; k >= 0; k--) { dst[out + k] = (uint8_t)((tab[c] >> 12) & 255u); c = (int)(tab[c] & 0xfffu); }
            dst[out + len - 1] = (uint8_t)first;
        } else
            return -1;
        out += len;
        if (next < 4096) {
            tab[next] = (uint32_t)old | (first << 12) | (((tab[old] >> 20) & 255u) << 20);
            tlen[next] = (uint16_t)(tlen[old] + 1);
            next++;
            if (next + 1 >= (1 << nbits) && nbits < 12) nbits++;       // libtiff's early change
        }
        old = code;
    }
    return out;
}


}  // namespace codec
}  // namespace rs
