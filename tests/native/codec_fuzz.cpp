// Corrupt-stream fuzz of the decoder cores (proj_roadsurf_b200/csrc/rs_codec_core.h), built with AddressSanitizer and UBSan by
// tests/test_codec_host.py: every stream of the input file is decoded intact, then many times with random bytes flipped,
// bytes cut off the end and random output capacities, into heap buffers of EXACTLY the stated sizes -- a decoder that reads
// past its input or writes past its capacity aborts the process.  The file: repeated [kind u8][raw length u32][n u32][n bytes].
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "rs_codec_core.h"

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd()
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

static long long run(int kind, const uint8_t *src, long long n, long long cap)
{
    uint8_t *in = (uint8_t *)malloc(n > 0 ? n : 1);            // exact sizes: the sanitizer sees any overrun
    uint8_t *out = (uint8_t *)malloc(cap > 0 ? cap : 1);
    if (n > 0) memcpy(in, src, n);
    long long got;
    if (kind == 0) {
        uint16_t hot[rs::codec::RS_INFLATE_HOT], sym[rs::codec::RS_INFLATE_SYM];
        uint8_t lens[rs::codec::RS_INFLATE_LEN];
        got = rs::codec::inflate_segment(in, n, out, cap, true, hot, 1, sym, lens);
        // the warp-per-segment decoder (one lane here) must agree on intact streams and stay inside its buffers on damaged ones
        uint8_t *out2 = (uint8_t *)malloc(cap > 0 ? cap : 1);
        rs::codec::WTables *t = (rs::codec::WTables *)malloc(sizeof(rs::codec::WTables));
        const long long got2 = rs::codec::inflate_segment_warp(in, n, out2, cap, true, *t, 0);
        if (got2 < -1 || got2 > cap || (got >= 0) != (got2 >= 0) || (got >= 0 && (got2 != got || memcmp(out, out2, got) != 0))) got = -2;
        free(t);
        // and the table-driven thread-per-segment decoder
        uint16_t tab[rs::codec::RS_T_SMEM];
        const long long got3 = rs::codec::inflate_segment_lut(in, n, out2, cap, true, tab, 1, sym, lens);
        if (got3 < -1 || got3 > cap || (got >= 0) != (got3 >= 0) || (got >= 0 && (got3 != got || memcmp(out, out2, got) != 0))) got = -2;
        free(out2);
    } else {
        uint32_t *tab = (uint32_t *)malloc(4096 * sizeof(uint32_t));
        uint16_t *len = (uint16_t *)malloc(4096 * sizeof(uint16_t));
        got = rs::codec::lzw_segment(in, n, out, cap, tab, len);
        free(tab);
        free(len);
    }
    free(in);
    free(out);
    return got;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    const int rounds = atoi(argv[2]);
    long long intact = 0, mutated = 0, accepted = 0;
    for (;;) {
        uint8_t kind;
        uint32_t raw_len, n;
        if (fread(&kind, 1, 1, f) != 1) break;
        if (fread(&raw_len, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1) return 3;
        std::vector<uint8_t> comp(n);
        if (n && fread(comp.data(), 1, n, f) != n) return 3;
        if (run(kind, comp.data(), n, raw_len) != (long long)raw_len) return 4;      // the intact stream decodes to its length
        intact++;
        for (int r = 0; r < rounds; r++) {
            std::vector<uint8_t> m(comp);
            const int flips = 1 + (int)(rnd() % 4);
            for (int k = 0; k < flips && !m.empty(); k++) m[rnd() % m.size()] ^= (uint8_t)(1u << (rnd() % 8));
            long long len = (long long)m.size();
            if (rnd() % 4 == 0 && len > 0) len = (long long)(rnd() % (uint32_t)len);          // truncated
            long long cap = raw_len;
            if (rnd() % 4 == 0) cap = (long long)(rnd() % (raw_len + 2));                    // a smaller (or 1 larger) segment
            const long long got = run(kind, m.data(), len, cap);
            if (got < -1 || got > cap) return 5;
            accepted += got >= 0 ? 1 : 0;
            mutated++;
        }
    }
    fclose(f);
    printf("%lld %lld %lld\n", intact, mutated, accepted);
    return 0;
}
