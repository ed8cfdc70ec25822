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
// DEFLATE, a WARP per segment (rs_codec.cu inflate_warp_kernel).  A Huffman stream is sequential, so a thread per segment
// (inflate_segment above) leaves the lanes of a warp in 32 different places of 32 different tables.  Here the 32 lanes run ONE
// decoder in lockstep -- all its state is lane-uniform -- and share what is not sequential:
//   * the input is held a 128-byte line at a time, one word per lane, and fed to the 64-bit bit buffer by shuffles;
//   * codes are looked up in first-level tables in the warp's shared memory (10 bits literal / length, 8 bits distance: an
//     entry is symbol << 4 | length; longer codes, rare, walk the canonical per-length counts);
//   * the tables are built by the lanes together (a lane per symbol, ranks among equal lengths by __match_any_sync);
//   * literals wait in a register of the lane their position selects and leave 32 at a time in one store; matches and stored
//     blocks are copied by all lanes; the Adler-32 of the output is a lane-strided sum.
// The same source compiles for the host with ONE lane (RS_W_LANES 1): tests/test_codec_host.py holds it to zlib and runs it
// under the sanitizers; the 32-lane form is held to zlib by the -m gpu tests.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define RS_W_LANES 32
RS_HD inline void w_sync() { __syncwarp(); }
RS_HD inline unsigned w_match(int v) { return __match_any_sync(0xffffffffu, v); }
RS_HD inline uint32_t w_shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
RS_HD inline int w_popc(unsigned m) { return __popc(m); }
RS_HD inline uint32_t w_brev(uint32_t v) { return __brev(v); }
#else
#define RS_W_LANES 1
inline void w_sync() {}
inline unsigned w_match(int) { return 1u; }
inline uint32_t w_shfl(uint32_t v, int) { return v; }
inline int w_popc(unsigned m) { return __builtin_popcount(m); }
inline uint32_t w_brev(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
#endif

enum { RS_W_LITBITS = 10, RS_W_DISTBITS = 8 };

struct WTables {                      // one per warp (shared memory on the device)
    uint16_t litlut[1 << RS_W_LITBITS];
    uint16_t distlut[1 << RS_W_DISTBITS];      // also the table of the code-length code while a dynamic header is read
    uint16_t count[2][16];            // codes per length: [0] literal / length, [1] distance
    uint16_t next[16], offs[16];      // build scratch: next code / next place in the symbol list, per length
    uint16_t symbol[2][288];          // symbols in canonical order (for the codes longer than a table index)
    uint8_t lens[320];
};

struct WBits {                        // LSB-first bit reader; every member is lane-uniform except `line`
    const uint8_t *src, *base;        // base = src rounded down to 4 bytes
    long long n;                      // bytes of the segment
    long long widx, lbase;            // next word to feed / first word of the line held in `line` (-1: none)
    uint32_t line;                    // this lane's word of the line
    unsigned long long buf;
    int cnt;                          // valid bits in buf
    long long fed;                    // bits of the segment fed into buf so far: consumed = fed - cnt

    RS_HD inline uint32_t load_word(long long w) const      // bytes outside [src, src + n) read as 0
    {
        const uint8_t *p = base + 4 * w;
        if (p >= src && p + 4 <= src + n) return *reinterpret_cast<const uint32_t *>(p);
        uint32_t v = 0;
        for (int k = 0; k < 4; k++)
            if (p + k >= src && p + k < src + n) v |= (uint32_t)p[k] << (8 * k);
        return v;
    }
    RS_HD inline uint32_t next_word(int lane)
    {
        if (lbase < 0 || widx - lbase >= RS_W_LANES) {
            lbase = widx;
            line = load_word(lbase + lane);
        }
        const uint32_t w = w_shfl(line, (int)(widx - lbase));
        widx++;
        return w;
    }
    RS_HD inline void seek(long long pos, int lane)          // continue at byte `pos` of the segment
    {
        const long long a = (long long)(src - base) + pos;
        widx = a >> 2;
        lbase = -1;
        const int sub = (int)(a & 3);
        buf = (unsigned long long)(next_word(lane) >> (8 * sub));
        cnt = 32 - 8 * sub;
        fed = 8 * pos + cnt;
    }
    RS_HD inline void refill(int lane)                       // afterwards cnt >= 32
    {
        if (cnt < 32) {
            buf |= (unsigned long long)next_word(lane) << cnt;
            cnt += 32;
            fed += 32;
        }
    }
    RS_HD inline uint32_t peek(int k) const { return (uint32_t)buf & ((1u << k) - 1u); }
    RS_HD inline void drop(int k) { buf >>= k; cnt -= k; }
    RS_HD inline uint32_t get(int k)                         // k <= 16, after a refill
    {
        const uint32_t v = peek(k);
        drop(k);
        return v;
    }
    RS_HD inline long long consumed() const { return fed - cnt; }
    RS_HD inline bool over() const { return consumed() > 8 * n; }
};

// lens[0 .. n) -> per-length counts, canonical symbol list, first-level table of `bits` index bits.  false: over-subscribed.
RS_HD inline bool w_build(WTables &t, const uint8_t *len, int n, int which, uint16_t *lut, int bits, int lane)
{
    uint16_t *count = t.count[which], *symbol = t.symbol[which];
    for (int i = lane; i < 16; i += RS_W_LANES) count[i] = 0;
    for (int i = lane; i < (1 << bits); i += RS_W_LANES) lut[i] = 0;
    w_sync();
    for (int c0 = 0; c0 < n; c0 += RS_W_LANES) {              // counts: the lowest lane of every group of equal lengths adds the group
        const int i = c0 + lane, l = i < n ? len[i] : 0;
        const unsigned m = w_match(l);
        if (l && (m & ((1u << lane) - 1u)) == 0) count[l] = (uint16_t)(count[l] + w_popc(m));
        w_sync();
    }
    int left = 1, code = 0, off = 0;
    for (int l = 1; l < 16; l++) {                            // lane-uniform: Kraft sum, first code and first place per length
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
        if (lane == 0) { t.next[l] = (uint16_t)code; t.offs[l] = (uint16_t)off; }
        code = (code + count[l]) << 1;
        off += count[l];
    }
    w_sync();
    for (int c0 = 0; c0 < n; c0 += RS_W_LANES) {
        const int i = c0 + lane, l = i < n ? len[i] : 0;
        const unsigned m = w_match(l);
        const int r = w_popc(m & ((1u << lane) - 1u));
        const int cd = l ? t.next[l] + r : 0, at = l ? t.offs[l] + r : 0;
        w_sync();
        if (l && r == 0) { t.next[l] = (uint16_t)(t.next[l] + w_popc(m)); t.offs[l] = (uint16_t)(t.offs[l] + w_popc(m)); }
        if (l) {
            symbol[at] = (uint16_t)i;
            if (l <= bits) {
                const uint32_t rev = w_brev((uint32_t)cd) >> (32 - l);
                for (uint32_t k = rev; k < (1u << bits); k += 1u << l) lut[k] = (uint16_t)((i << 4) | l);
            }
        }
        w_sync();
    }
    return true;
}

// one code: table hit, or the canonical walk for codes longer than the table index.  Needs >= 15 bits in b.buf.  -1: no such code
RS_HD inline int w_decode(WBits &b, const uint16_t *lut, int bits, const uint16_t *count, const uint16_t *symbol)
{
    const uint32_t e = lut[b.peek(bits)];
    if (e) {
        b.drop((int)(e & 15u));
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;
    uint32_t v = (uint32_t)b.buf;
    for (int l = 1; l < 16; l++) {
        code |= (int)(v & 1u);
        v >>= 1;
        const int c = count[l];
        if (code - c < first) {
            b.drop(l);
            return symbol[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// returns bytes written, or -1 (corrupt / unsupported stream, or output larger than cap); every lane returns the same value
RS_HD inline long long inflate_segment_warp(const uint8_t *src, long long n, uint8_t *dst, long long cap, bool zlib_wrapper, WTables &t,
                                            int lane)
{
    constexpr int L = RS_W_LANES;
    WBits b;
    b.src = src;
    b.base = src - ((uintptr_t)src & 3u);
    b.n = n;
    if (zlib_wrapper && n < 2) return -1;
    b.seek(0, lane);
    if (zlib_wrapper) {
        b.refill(lane);
        const uint32_t cmf = b.get(8), flg = b.get(8);
        if ((cmf & 15u) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) return -1;
    }
    long long out = 0, pend_lo = 0;             // output bytes [pend_lo, out) wait in `pend` of lane (position % L)
    uint32_t pend = 0;
    auto flush = [&]() {
        const long long p = (pend_lo & ~(long long)(L - 1)) + lane;
        if (p >= pend_lo && p < out) dst[p] = (uint8_t)pend;
        pend_lo = out;
    };
    for (;;) {
        b.refill(lane);
        const uint32_t last = b.get(1), type = b.get(2);
        if (type == 0) {
            b.drop(b.cnt & 7);                                 // to the next byte boundary
            b.refill(lane);
            const uint32_t len = b.get(16);
            b.refill(lane);
            const uint32_t nlen = b.get(16);
            const long long pos = b.consumed() >> 3;
            if ((len ^ 0xffffu) != nlen || b.over() || pos + len > n || out + len > cap) return -1;
            flush();
            for (long long i = lane; i < (long long)len; i += L) dst[out + i] = src[pos + i];
            out += len;
            pend_lo = out;
            w_sync();
            b.seek(pos + len, lane);
        } else if (type == 1 || type == 2) {
            if (type == 1) {
                for (int i = lane; i < 288; i += L) t.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                w_sync();
                w_build(t, t.lens, 288, 0, t.litlut, RS_W_LITBITS, lane);
                w_sync();
                for (int i = lane; i < 30; i += L) t.lens[i] = 5;
                w_sync();
                w_build(t, t.lens, 30, 1, t.distlut, RS_W_DISTBITS, lane);
            } else {
                b.refill(lane);
                const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
                if (nlen > 286 || ndist > 30) return -1;
                for (int i = lane; i < 19; i += L) t.lens[i] = 0;
                w_sync();
                for (int i = 0; i < ncode; i++) {
                    b.refill(lane);
                    const uint32_t v = b.get(3);
                    if (lane == 0) t.lens[CL_ORDER[i]] = (uint8_t)v;
                }
                w_sync();
                if (!w_build(t, t.lens, 19, 1, t.distlut, 7, lane)) return -1;      // the code-length code (at most 7 bits)
                w_sync();
                int idx = 0;
                while (idx < nlen + ndist) {
                    b.refill(lane);
                    const int sy = w_decode(b, t.distlut, 7, t.count[1], t.symbol[1]);
                    if (sy < 0 || b.over()) return -1;
                    if (sy < 16) {
                        if (lane == 0) t.lens[idx] = (uint8_t)sy;
                        idx++;
                    } else {
                        int rep, val = 0;
                        if (sy == 16) {
                            if (idx == 0) return -1;
                            w_sync();                                // lens[idx - 1] may have been written a moment ago
                            val = t.lens[idx - 1];
                            rep = 3 + (int)b.get(2);
                        } else if (sy == 17) rep = 3 + (int)b.get(3);
                        else rep = 11 + (int)b.get(7);
                        if (idx + rep > nlen + ndist) return -1;
                        for (int i = lane; i < rep; i += L) t.lens[idx + i] = (uint8_t)val;
                        idx += rep;
                    }
                }
                w_sync();
                if (t.lens[256] == 0) return -1;                    // no end-of-block code
                if (!w_build(t, t.lens + nlen, ndist, 1, t.distlut, RS_W_DISTBITS, lane)) return -1;
                w_sync();
                if (!w_build(t, t.lens, nlen, 0, t.litlut, RS_W_LITBITS, lane)) return -1;
            }
            w_sync();
            for (;;) {
                b.refill(lane);
                const int sy = w_decode(b, t.litlut, RS_W_LITBITS, t.count[0], t.symbol[0]);
                if (sy < 0 || b.over()) return -1;
                if (sy < 256) {
                    if (out >= cap) return -1;
                    if ((int)(out & (L - 1)) == lane) pend = (uint32_t)sy;
                    out++;
                    if ((out & (L - 1)) == 0) flush();
                } else if (sy == 256) break;
                else {
                    const int ls = sy - 257;
                    if (ls >= 29) return -1;
                    const int len = LEN_BASE[ls] + (int)b.get(LEN_EXTRA[ls]);
                    b.refill(lane);
                    const int ds = w_decode(b, t.distlut, RS_W_DISTBITS, t.count[1], t.symbol[1]);
                    if (ds < 0 || ds >= 30) return -1;
                    const long long d = DIST_BASE[ds] + (long long)b.get(DIST_EXTRA[ds]);
                    if (d > out || out + len > cap || b.over()) return -1;
                    flush();
                    w_sync();                                       // the bytes the match reads are in memory
                    for (int i = lane; i < len; i += L) dst[out + i] = dst[out - d + (d >= len ? i : i % (int)d)];
                    out += len;
                    pend_lo = out;
                    w_sync();
                }
            }
        } else
            return -1;
        if (b.over()) return -1;
        if (last) break;
    }
    flush();
    w_sync();
    if (zlib_wrapper) {
        // Adler-32 of the output against the stream's big-endian trailer at the next byte boundary:
        // s1 = 1 + sum d_i, s2 = out + sum (out - i) d_i (mod 65521), lane-strided
        const long long tp = (b.consumed() + 7) >> 3;
        if (tp + 4 > n) return -1;
        const uint32_t want = ((uint32_t)src[tp] << 24) | ((uint32_t)src[tp + 1] << 16) | ((uint32_t)src[tp + 2] << 8) | (uint32_t)src[tp + 3];
        unsigned long long s1 = 0, s2 = 0;
        int k = 0;
        for (long long i = lane; i < out; i += L) {
            const unsigned long long v = dst[i];
            s1 += v;
            s2 += (unsigned long long)((out - i) % 65521) * v;
            if (++k == 4096) { s1 %= 65521u; s2 %= 65521u; k = 0; }
        }
        s1 %= 65521u; s2 %= 65521u;
        for (int o = L >> 1; o > 0; o >>= 1) {
            s1 += w_shfl((uint32_t)s1, lane ^ o);
            s2 += w_shfl((uint32_t)s2, lane ^ o);
        }
        s1 = (s1 + 1) % 65521u;
        s2 = (s2 + (unsigned long long)(out % 65521)) % 65521u;
        if ((((uint32_t)s2 << 16) | (uint32_t)s1) != want) return -1;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// DEFLATE, a thread per segment with first-level TABLES (rs_codec.cu inflate_lut_kernel).  inflate_segment above walks a code bit
// by bit through the per-length counts (15 dependent steps for a long code, its tables in local memory); here a code is one
// look-up of the next RS_T_LITBITS (literal / length) or RS_T_DISTBITS (distance) bits in a table of symbol << 4 | length, and
// only the rare longer codes take the canonical walk.  The tables are strided like the counts (element i at p[i * stride]):
// on the device column t of the block's shared memory is thread t's, so that the 32 decoders of a warp -- each in its own
// place of its own table -- never meet in a bank.  The lanes of a warp take their segments in step (one each per round): the
// table builds at the head of a segment run together and every lane of the warp decodes -- a warp per segment
// (inflate_segment_warp) issues the same ~60 instructions per symbol for ONE decoder.
// ---------------------------------------------------------------------------------------------
#ifndef RS_T_LITBITS
#define RS_T_LITBITS 7
#endif
#ifndef RS_T_DISTBITS
#define RS_T_DISTBITS 5
#endif
#ifndef RS_T_INLINE_SUM_MAX
#define RS_T_INLINE_SUM_MAX (1ll << 28)   // longest segment whose Adler-32 sums stay unreduced in 64 bits (tests lower it)
#endif
enum { RS_T_SMEM = (1 << RS_T_LITBITS) + (1 << RS_T_DISTBITS) + 64 };    // strided uint16 elements per decoder

struct TBits {                        // LSB-first bit reader over [src, src + n), fed by aligned 32-bit words, one word ahead
    const uint8_t *src, *base;        // base = src rounded down to 4 bytes
    long long n, widx;                // widx: index of the word held in `ahead`
    uint32_t ahead;                   // loaded a refill early: its latency is off the decode chain
    unsigned long long buf;
    int cnt;                          // valid bits in buf
    long long fed;                    // bits of the segment fed into buf so far: consumed = fed - cnt
    RS_HD inline uint32_t load_word(long long w) const      // bytes outside [src, src + n) read as 0
    {
        const uint8_t *p = base + 4 * w;
        if (p >= src && p + 4 <= src + n) return *reinterpret_cast<const uint32_t *>(p);
        uint32_t v = 0;
        for (int k = 0; k < 4; k++)
            if (p + k >= src && p + k < src + n) v |= (uint32_t)p[k] << (8 * k);
        return v;
    }
    RS_HD inline void seek(long long pos)
    {
        const long long a = (long long)(src - base) + pos;
        widx = a >> 2;
        const int sub = (int)(a & 3);
        buf = (unsigned long long)(load_word(widx) >> (8 * sub));
        ahead = load_word(++widx);
        cnt = 32 - 8 * sub;
        fed = 8 * pos + cnt;
    }
    RS_HD inline void refill()                               // afterwards cnt >= 32
    {
        if (cnt < 32) {
            buf |= (unsigned long long)ahead << cnt;
            ahead = load_word(++widx);
            cnt += 32;
            fed += 32;
        }
    }
    RS_HD inline uint32_t peek(int k) const { return (uint32_t)buf & ((1u << k) - 1u); }
    RS_HD inline void drop(int k) { buf >>= k; cnt -= k; }
    RS_HD inline uint32_t get(int k) { const uint32_t v = peek(k); drop(k); return v; }
    RS_HD inline long long consumed() const { return fed - cnt; }
    RS_HD inline bool over() const { return consumed() > 8 * n; }
};

struct TCode {                        // one canonical code: strided table + counts, plain symbol list
    uint16_t *lut, *count;            // [1 << bits], [16]; strided
    uint16_t *symbol;
    int bits, stride;
};

// lens[0 .. n) -> counts, symbol list, table.  tmp: 32 strided uint16 of scratch.  false: over-subscribed
RS_HD inline bool t_build(TCode &h, const uint8_t *len, int n, uint16_t *tmp)
{
    const int st = h.stride;
    for (int i = 0; i < 16; i++) h.count[i * st] = 0;
    for (int i = 0; i < (1 << h.bits); i++) h.lut[i * st] = 0;
    for (int i = 0; i < n; i++) h.count[len[i] * st]++;
    int left = 1, code = 0, off = 0;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= h.count[l * st];
        if (left < 0) return false;
        tmp[l * st] = (uint16_t)code;                // next code of this length
        tmp[(16 + l) * st] = (uint16_t)off;          // next place in the symbol list
        code = (code + h.count[l * st]) << 1;
        off += h.count[l * st];
    }
    for (int i = 0; i < n; i++) {
        const int l = len[i];
        if (!l) continue;
        const uint32_t cd = tmp[l * st]++;
        h.symbol[tmp[(16 + l) * st]++] = (uint16_t)i;
        if (l <= h.bits) {
            uint32_t rev = 0;
            for (int k = 0; k < l; k++) rev |= ((cd >> k) & 1u) << (l - 1 - k);
            for (uint32_t k = rev; k < (1u << h.bits); k += 1u << l) h.lut[k * st] = (uint16_t)((i << 4) | l);
        }
    }
    return true;
}

RS_HD inline uint32_t t_brev15(uint32_t v)          // the low 15 bits of v, reversed
{
#ifdef __CUDACC__
    return __brev(v) >> 17;
#else
    uint32_t r = 0;
    for (int k = 0; k < 15; k++) r |= ((v >> k) & 1u) << (14 - k);
    return r;
#endif
}

RS_HD inline int t_decode(TBits &b, const TCode &h)           // needs >= 15 bits in b.buf; -1: no such code
{
    const uint32_t e = h.lut[b.peek(h.bits) * h.stride];
    if (e) {
        b.drop((int)(e & 15u));
        return (int)(e >> 4);
    }
    // a code longer than the table index (or none): the canonical walk over the per-length counts, on the next 15 bits taken
    // MSB-first; lengths up to the table index only advance the first code / first place
    const int rev = (int)t_brev15((uint32_t)b.buf);
    int first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        const int c = h.count[l * h.stride];
        if (l > h.bits) {
            const int code = rev >> (15 - l);
            if (code - first < c) {
                b.drop(l);
                return h.symbol[index + (code - first)];
            }
        }
        index += c;
        first = (first + c) << 1;
    }
    return -1;
}

// bytes [i, min(i + N, len)) of a match: N independent loads, then the stores and the Adler-32 sums in order
template <int N>
RS_HD inline void t_piece(const uint8_t *sp, uint8_t *dp, int i, int len, unsigned long long &s1, unsigned long long &s2)
{
    const int m = len - i < N ? len - i : N;
    uint8_t v[N];
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = k < m ? sp[i + k] : (uint8_t)0;
#pragma unroll
    for (int k = 0; k < N; k++)
        if (k < m) { dp[i + k] = v[k]; s1 += v[k]; s2 += s1; }
}

// The decoder as a state machine, so that the 32 decoders of a warp can be STEPPED TOGETHER (rs_codec.cu): written as one
// function with its loops inside, the lanes of a warp part at the first data-dependent branch and, with the early exits of a
// decoder, never meet again before the function returns -- one active lane per instruction, measured.  Stepped from a
// warp-uniform loop (a vote per step) they reconverge after every symbol.
//   begin -> { header -> { symbol [-> copy_match] }* }* -> trailer;   state: what the decoder needs next
struct TInflate {
    enum { HEADER = 0, SYMBOLS = 1, TRAILER = 2, DONE = 3, FAIL = 4, MATCH = 5 };     // MATCH: a decoded match waits to be copied
    TBits b;
    TCode lit, dist;
    uint16_t *tmp;
    uint8_t *lens;
    uint8_t *dst;
    long long cap, out;
    unsigned long long s1, s2;        // Adler-32 sums of the output so far, not reduced (inline: segments up to 2^28 bytes)
    int state, last;
    int mlen;                         // MATCH: length and distance of the waiting match
    long long mdist;
    bool zlib, inline_sum;

    // tab: RS_T_SMEM strided uint16; sym: RS_INFLATE_SYM uint16; lens_: RS_INFLATE_LEN uint8
    RS_HD inline void begin(const uint8_t *src, long long n, uint8_t *dst_, long long cap_, bool zlib_wrapper, uint16_t *tab, int stride,
                            uint16_t *sym, uint8_t *lens_)
    {
        b.src = src;
        b.base = src - ((uintptr_t)src & 3u);
        b.n = n;
        dst = dst_; cap = cap_; out = 0; last = 0; zlib = zlib_wrapper; lens = lens_;
        s1 = 1; s2 = 0; inline_sum = cap_ <= RS_T_INLINE_SUM_MAX;
        lit = TCode{tab, tab + ((1 << RS_T_LITBITS) + (1 << RS_T_DISTBITS)) * stride, sym, RS_T_LITBITS, stride};
        dist = TCode{tab + (1 << RS_T_LITBITS) * stride, lit.count + 16 * stride, sym + 288, RS_T_DISTBITS, stride};
        tmp = lit.count + 32 * stride;
        state = HEADER;
        if (zlib && n < 2) { state = FAIL; return; }
        b.seek(0);
        if (zlib) {
            b.refill();
            const uint32_t cmf = b.get(8), flg = b.get(8);
            if ((cmf & 15u) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) state = FAIL;
        }
    }

    // one block header; a stored block is copied here as a whole
    RS_HD inline void header()
    {
        state = FAIL;
        b.refill();
        last = (int)b.get(1);
        const uint32_t type = b.get(2);
        if (type == 0) {
            b.drop(b.cnt & 7);
            b.refill();
            const uint32_t len = b.get(16);
            b.refill();
            const uint32_t nlen = b.get(16);
            const long long pos = b.consumed() >> 3;
            if ((len ^ 0xffffu) != nlen || b.over() || pos + len > b.n || out + len > cap) return;
            for (uint32_t i = 0; i < len; i++) {
                const uint8_t v = b.src[pos + i];
                dst[out + i] = v;
                s1 += v; s2 += s1;
            }
            out += len;
            b.seek(pos + len);
            state = last ? TRAILER : HEADER;
            return;
        }
        if (type == 1) {
            for (int i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
            t_build(lit, lens, 288, tmp);
            for (int i = 0; i < 30; i++) lens[i] = 5;
            t_build(dist, lens, 30, tmp);
            state = SYMBOLS;
            return;
        }
        if (type != 2) return;
        b.refill();
        const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
        if (nlen > 286 || ndist > 30) return;
        for (int i = 0; i < 19; i++) lens[i] = 0;
        for (int i = 0; i < ncode; i++) {
            b.refill();
            lens[CL_ORDER[i]] = (uint8_t)b.get(3);
        }
        TCode cl{dist.lut, dist.count, dist.symbol, RS_T_DISTBITS < 7 ? RS_T_DISTBITS : 7, dist.stride};    // the code-length code
        if (!t_build(cl, lens, 19, tmp)) return;
        int idx = 0;
        while (idx < nlen + ndist) {
            b.refill();
            const int sy = t_decode(b, cl);
            if (sy < 0 || b.over()) return;
            if (sy < 16) lens[idx++] = (uint8_t)sy;
            else {
                int rep, val = 0;
                if (sy == 16) {
                    if (idx == 0) return;
                    val = lens[idx - 1];
                    rep = 3 + (int)b.get(2);
                } else if (sy == 17) rep = 3 + (int)b.get(3);
                else rep = 11 + (int)b.get(7);
                if (idx + rep > nlen + ndist) return;
                while (rep--) lens[idx++] = (uint8_t)val;
            }
        }
        if (lens[256] == 0) return;                             // no end-of-block code
        if (!t_build(dist, lens + nlen, ndist, tmp)) return;
        if (!t_build(lit, lens, nlen, tmp)) return;
        state = SYMBOLS;
    }

    // one literal, match (decoded and checked, not yet copied: state MATCH) or end-of-block code
    RS_HD inline void symbol()
    {
        b.refill();
        const int sy = t_decode(b, lit);
        if (sy < 0 || b.over()) { state = FAIL; return; }
        if (sy < 256) {
            if (out >= cap) { state = FAIL; return; }
            dst[out++] = (uint8_t)sy;
            s1 += (unsigned)sy; s2 += s1;
            return;
        }
        if (sy == 256) {
            state = last ? TRAILER : HEADER;
            return;
        }
        const int ls = sy - 257;
        if (ls >= 29) { state = FAIL; return; }
        const int len = LEN_BASE[ls] + (int)b.get(LEN_EXTRA[ls]);
        b.refill();
        const int ds = t_decode(b, dist);
        if (ds < 0 || ds >= 30) { state = FAIL; return; }
        const long long d = DIST_BASE[ds] + (long long)b.get(DIST_EXTRA[ds]);
        if (d > out || out + len > cap || b.over()) { state = FAIL; return; }
        mlen = len;
        mdist = d;
        state = MATCH;
    }

    // the copy of the match decoded by symbol().  Apart from the stepping kernel's reasons (the copies of a warp's decoders run
    // together, at most one trip through memory for all of them), a decoder is simply symbol(); copy_match() in turn.
    RS_HD inline void copy_match()
    {
        const int len = mlen;
        const long long d = mdist;
        state = SYMBOLS;
        uint8_t *dp = dst + out;
        const uint8_t *sp = dp - d;
        out += len;
        // Every byte a match reads was stored a moment ago and comes back through the memory system: the copy is arranged so
        // that a match costs one such round trip per piece of 4 or 8 bytes (independent loads, then the stores), not one per
        // byte.
        if (d >= 8 || d >= len) {                     // no piece reads what it writes: 4 bytes, then 8 at a time
            t_piece<4>(sp, dp, 0, len, s1, s2);       // most matches of a fast encoder are 3 or 4 bytes long
            for (int i = 4; i < len; i += 8) t_piece<8>(sp, dp, i, len, s1, s2);
        } else if (d >= 4) {
            for (int i = 0; i < len; i += 4) t_piece<4>(sp, dp, i, len, s1, s2);
        } else {                                      // a pattern of d < 4 bytes repeated: read once, kept in a register
            unsigned long long pat = 0;
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (k < d) pat |= (unsigned long long)sp[k] << (8 * k);
            int j = 0;
            for (int i = 0; i < len; i++) {
                const uint8_t v = (uint8_t)(pat >> (8 * j));
                dp[i] = v;
                s1 += v; s2 += s1;
                j = j + 1 == (int)d ? 0 : j + 1;
            }
        }
    }

    // Adler-32 of the output against the stream's trailer
    RS_HD inline void trailer()
    {
        state = FAIL;
        if (b.over()) return;
        if (zlib) {
            const long long tp = (b.consumed() + 7) >> 3;
            if (tp + 4 > b.n) return;
            const uint8_t *q = b.src + tp;
            const uint32_t want = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
            uint32_t a1 = (uint32_t)(s1 % 65521u), a2 = (uint32_t)(s2 % 65521u);
            if (!inline_sum) {                                   // a segment too long for the unreduced sums: sum it again
                a1 = 1; a2 = 0;
                for (long long i = 0; i < out;) {
                    const long long stop = i + 5552 < out ? i + 5552 : out;      // the largest run that cannot overflow 32 bits
                    for (; i < stop; i++) { a1 += dst[i]; a2 += a1; }
                    a1 %= 65521u;
                    a2 %= 65521u;
                }
            }
            if (((a2 << 16) | a1) != want) return;
        }
        state = DONE;
    }
    RS_HD inline long long result() const { return state == DONE ? out : -1; }
};

// the machine run to its end by one caller (host tests; a lone decoder)
RS_HD inline long long inflate_segment_lut(const uint8_t *src, long long n, uint8_t *dst, long long cap, bool zlib_wrapper, uint16_t *tab,
                                           int stride, uint16_t *sym, uint8_t *lens)
{
    TInflate d;
    d.begin(src, n, dst, cap, zlib_wrapper, tab, stride, sym, lens);
    while (d.state == TInflate::HEADER || d.state == TInflate::SYMBOLS || d.state == TInflate::MATCH) {
        if (d.state == TInflate::HEADER) d.header();
        else if (d.state == TInflate::MATCH) d.copy_match();
        else d.symbol();
    }
    if (d.state == TInflate::TRAILER) d.trailer();
    return d.result();
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
