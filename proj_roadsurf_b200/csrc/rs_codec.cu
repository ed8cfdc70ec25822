// Tile ingest, decompression on the device: the compressed strips / internal tiles of a batch of GeoTIFFs cross
// the host link as they are in the files and are decoded here, one decoder per segment, straight into the sample buffer
// rs_assemble_tiles_* reads (predictor, byte order, band selection and the 16 -> 8 bit rescale follow in assemble_kernel).
// This is what rasterio.open(tile).read() does through libtiff inside the reference's pair loop
// (scripts/functions/fct_misc.py:76-77; the tiles of config/config_stats.yaml:39 come from a tile server over COGs).
//
//   codec 8 / 32946  zlib-wrapped DEFLATE (RFC 1950 / 1951): stored, fixed and dynamic Huffman blocks; matches copy from the
//                    output itself (the 32 KiB window is the already written part of the segment); the Adler-32 trailer is
//                    checked.  Three decoders (RS_INFLATE, rs_codec_core.h):
//                      lut   inflate_lut_kernel (default): a thread per segment, first-level code tables as shared-memory
//                            columns, the 32 decoders of a warp stepped together as state machines
//                      warp  inflate_warp_kernel: a warp per segment, one decoder's state uniform over the lanes
//                      bits  decode_kernel: a thread per segment, every code walked bit by bit through the per-length counts
//   codec 5          TIFF LZW (MSB-first codes of 9 - 12 bits, ClearCode 256, EOI 257, the "early change" of libtiff); the
//                    string table (4096 x 6 bytes per decoder) lives in a scratch buffer, decoders run grid-strided
//   codec 1          none: a copy
// A decoder is sequential by nature (every code depends on the bits before it); the parallelism is across segments -- a
// 256 x 256 tile stored in 8-row strips is 32 of them, a canton 67 M.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "rs_codec_core.h"
#include "rs_internal.h"

namespace rs {

namespace {

using namespace codec;

struct CodecArgs {
    const uint8_t *comp;
    const long long *comp_off;        // [n_seg + 1]
    uint8_t *raw;
    const long long *raw_off;         // [n_seg + 1]: where each segment goes and how many bytes it may produce
    int n_seg, codec;
    uint32_t *lzw_tab;                // [n_threads][4096]
    uint16_t *lzw_len;
    int *status;                      // latched: RS_ERR_CODEC when a segment does not decode to its expected size
    int *bad_segment;                 // index of (one of) the failing segments, -1 = none
};

constexpr int DEC_THREADS = 128;      // decoders per block

__global__ void __launch_bounds__(DEC_THREADS) decode_kernel(const CodecArgs a)
{
    // the per-length counts of the two Huffman codes are read for every bit: column t of this array is thread t's (12 KiB per
    // block); the symbol lists and code lengths are touched once per code and stay in the thread's local memory
    __shared__ uint16_t hot[RS_INFLATE_HOT * DEC_THREADS];
    uint16_t sym[RS_INFLATE_SYM];
    uint8_t lens[RS_INFLATE_LEN];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int s = tid; s < a.n_seg; s += nthr) {
        const uint8_t *src = a.comp + a.comp_off[s];
        const long long n = a.comp_off[s + 1] - a.comp_off[s];
        uint8_t *dst = a.raw + a.raw_off[s];
        const long long cap = a.raw_off[s + 1] - a.raw_off[s];
        long long got;
        if (a.codec == 1) {
            got = n < cap ? n : cap;
            for (long long i = 0; i < got; i++) dst[i] = src[i];
        } else if (a.codec == 5)
            got = lzw_segment(src, n, dst, cap, a.lzw_tab + (size_t)tid * 4096, a.lzw_len + (size_t)tid * 4096);
        else
            got = inflate_segment(src, n, dst, cap, true, hot + threadIdx.x, DEC_THREADS, sym, lens);
        // libtiff pads nothing: a strip decodes to exactly rows * row_bytes (the last strip of an image to its remaining rows)
        if (got != cap) {
            atomicMin(a.status, (int)RS_ERR_CODEC);
            atomicMax(a.bad_segment, s);
        }
    }
}

// DEFLATE, a warp per segment (rs_codec_core.h inflate_segment_warp): one decoder per warp with its first-level tables in shared
// memory, the lanes sharing the input line, the table builds, the copies and the checksum
constexpr int WDEC_WARPS = 8;

__global__ void __launch_bounds__(WDEC_WARPS * 32, 4) inflate_warp_kernel(const CodecArgs a)
{
    __shared__ WTables tables[WDEC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wid = blockIdx.x * WDEC_WARPS + warp, nw = gridDim.x * WDEC_WARPS;
    for (int s = wid; s < a.n_seg; s += nw) {
        const long long c0 = a.comp_off[s], r0 = a.raw_off[s];
        const long long n = a.comp_off[s + 1] - c0, cap = a.raw_off[s + 1] - r0;
        const long long got = inflate_segment_warp(a.comp + c0, n, a.raw + r0, cap, true, tables[warp], lane);
        if (got != cap && lane == 0) {
            atomicMin(a.status, (int)RS_ERR_CODEC);
            atomicMax(a.bad_segment, s);
        }
        __syncwarp();
    }
}

// DEFLATE, a thread per segment with first-level tables in shared memory (rs_codec_core.h inflate_segment_lut): column t of the
// block's array is thread t's tables; the threads of a warp take their segments in step
#ifndef RS_T_MATCH_BATCH
#define RS_T_MATCH_BATCH 16          // lanes with a waiting match that start a round of copies
#endif
#ifndef RS_TDEC_THREADS
#define RS_TDEC_THREADS 64
#endif
constexpr int TDEC_THREADS = RS_TDEC_THREADS;

__global__ void __launch_bounds__(TDEC_THREADS) inflate_lut_kernel(const CodecArgs a)
{
    extern __shared__ uint16_t tdec_tab[];          // [RS_T_SMEM][TDEC_THREADS]
    uint16_t sym[RS_INFLATE_SYM];
    uint8_t lens[RS_INFLATE_LEN];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x, lane = threadIdx.x & 31;
    // the lanes of a warp take one segment each per round and step their decoders together: every loop below is warp-uniform
    // (its condition is a vote), so the lanes meet again after each header, each symbol and each trailer
    for (int s0 = tid - lane; s0 < a.n_seg; s0 += nthr) {
        const int s = s0 + lane;
        const bool valid = s < a.n_seg;
        TInflate d;
        d.state = TInflate::DONE;
        long long cap = 0;
        if (valid) {
            const long long c0 = a.comp_off[s], r0 = a.raw_off[s];
            cap = a.raw_off[s + 1] - r0;
            d.begin(a.comp + c0, a.comp_off[s + 1] - c0, a.raw + r0, cap, true, tdec_tab + threadIdx.x, TDEC_THREADS, sym, lens);
        }
        for (;;) {
            const bool hd = d.state == TInflate::HEADER;
            if (__any_sync(0xffffffffu, hd)) {
                if (hd) d.header();
                __syncwarp();
            }
            bool more = false;
            for (;;) {                              // symbols, until every lane's block ends (or one reaches its next header)
                // A decoded match is not copied at once: its lane waits (state MATCH) until half of the warp's decoders hold
                // one or no lane can decode on -- then the waiting copies run together.  Every step used to pay the match
                // path (a fifth of the symbols of a fast encoder's stream are matches, so some lane always had one) at a
                // quarter of the lanes and one trip through memory.
                const unsigned am = __ballot_sync(0xffffffffu, d.state == TInflate::SYMBOLS);
                const unsigned mm = __ballot_sync(0xffffffffu, d.state == TInflate::MATCH);
                if (!am && !mm) break;
                if (__popc(mm) >= RS_T_MATCH_BATCH || !am) {
                    if (d.state == TInflate::MATCH) d.copy_match();
                    continue;
                }
                if (d.state == TInflate::SYMBOLS) d.symbol();
                if (__any_sync(0xffffffffu, d.state == TInflate::HEADER)) {
                    if (d.state == TInflate::MATCH) d.copy_match();     // (MATCH lanes are in SYMBOLS again before the headers)
                    more = true;
                    break;
                }
            }
            if (!more && !__any_sync(0xffffffffu, d.state == TInflate::HEADER)) break;
        }
        if (d.state == TInflate::TRAILER) d.trailer();
        __syncwarp();
        if (valid && d.result() != cap) {
            atomicMin(a.status, (int)RS_ERR_CODEC);
            atomicMax(a.bad_segment, s);
        }
    }
}

}  // namespace

int launch_decode_segments(rs_ctx *ctx, const uint8_t *comp, const long long *comp_off, int n_seg, int codec, uint8_t *raw,
                           const long long *raw_off, cudaStream_t st)
{
    if (n_seg <= 0) return RS_OK;
    if (codec != 1 && codec != 5 && codec != 8 && codec != 32946) return RS_ERR_UNSUPPORTED;
    CodecArgs a{comp, comp_off, raw, raw_off, n_seg, codec == 32946 ? 8 : codec, nullptr, nullptr, ctx->d_status, ctx->d_counters + 8};
    int blocks = (n_seg + DEC_THREADS - 1) / DEC_THREADS;
    if (blocks > ctx->sm_count * 48) blocks = ctx->sm_count * 48;       // grid-stride over the segments
    if (codec == 5) {                                      // string tables: bound the number of concurrent decoders
        const int max_blocks = ctx->sm_count * 2;
        if (blocks > max_blocks) blocks = max_blocks;
        int rc = ensure(ctx, ctx->lzw_scratch, (size_t)blocks * DEC_THREADS * 4096 * (sizeof(uint32_t) + sizeof(uint16_t)));
        if (rc) return rc;
        a.lzw_tab = (uint32_t *)ctx->lzw_scratch.p;
        a.lzw_len = (uint16_t *)(a.lzw_tab + (size_t)blocks * DEC_THREADS * 4096);
    }
    RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_counters + 8, 0xff, sizeof(int), st));     // bad_segment = -1
    // DEFLATE: RS_INFLATE = lut (default: a thread per segment, first-level tables in shared memory) | warp (a warp per segment)
    // | bits (a thread per segment, canonical walk bit by bit) -- three decoders held to each other and to zlib by the tests
    const char *mode = getenv("RS_INFLATE");
    const char *env = getenv("RS_INFLATE_WARP");              // older switch: 1 = warp, 0 = bits
    const bool use_warp = (mode && mode[0] == 'w') || (!mode && env && atoi(env) != 0);
    const bool use_bits = (mode && mode[0] == 'b') || (!mode && env && atoi(env) == 0);
    if (a.codec == 8 && !use_warp && !use_bits) {
        const size_t smem = sizeof(uint16_t) * RS_T_SMEM * TDEC_THREADS;
        RS_CUDA_OK(ctx, cudaFuncSetAttribute(inflate_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        int tblocks = (n_seg + TDEC_THREADS - 1) / TDEC_THREADS;
        if (tblocks > ctx->sm_count * per_sm) tblocks = ctx->sm_count * per_sm;        // one wave: threads stride over the segments
        inflate_lut_kernel<<<tblocks, TDEC_THREADS, smem, st>>>(a);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
        return RS_OK;
    }
    if (a.codec == 8 && use_warp) {
        int wblocks = (n_seg + WDEC_WARPS - 1) / WDEC_WARPS;
        if (wblocks > ctx->sm_count * 16) wblocks = ctx->sm_count * 16;                // warps stride over the segments
        inflate_warp_kernel<<<wblocks, WDEC_WARPS * 32, 0, st>>>(a);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
        return RS_OK;
    }
    decode_kernel<<<blocks, DEC_THREADS, 0, st>>>(a);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
