// K1+K2: fused scanline rasterization of road polygons and per-road zonal accumulation (sm_100a).
//
// Replaces, for a whole road-major pair list in one launch, the reference's per-(road, tile)
// Python loop  scripts/statistical_analysis/statistical_analysis.py:180-193  whose body is
// scripts/functions/fct_misc.py:57-123 get_pixel_values = rasterio.mask.mask(crop=True) (:77,
// GDAL GDALRasterizeGeometries -> GDALdllImageFilledPolygon) + np.extract (:95).
//
// Work decomposition
//   item  = up to PPI consecutive (road, tile) pairs of ONE road (prep_items_kernel builds the list);
//   team  = one warp; teams pull items from a global counter (persistent grid, whole waves of CTAs);
//   a team folds its item into a team-private shared-memory histogram; a road that fits one item is
//   written once with plain 128-bit stores, a road split over several items is pre-zeroed by
//   prep_items_kernel and merged with integer atomics (exact, order-independent).
//
// Per pair the even-odd scanline fill is evaluated on a BIT MASK instead of sorted crossing lists:
//   GDAL sorts the rounded crossings c_0 <= c_1 <= ... of a scanline and burns [c_2i, c_2i+1 - 1]; for
//   the even crossing counts closed rings produce, pixel x is burnt iff #{ j : c_j <= x } is odd.  So
//   each crossing toggles one bit (atomicXor in shared memory), and the inside mask of a row is the
//   prefix-XOR of its toggle bits (5 shift-xor steps per 32-bit word + a carry).  No sort, no pairing,
//   no capacity limit on crossings per scanline; holes and overlapping parts cancel by construction.
//   Horizontal edges lying exactly on a scanline (burnt separately by GDAL when they run towards -x)
//   are OR-ed into the inside mask in a second edge pass that only runs when such an edge exists.
// Phases per (pair, chunk of rows):
//   cull     per-road vertex chunks carry bounds; a ballot keeps the chunks that can touch the window
//   edges    lanes take edges (TMA-staged in shared memory), transform both ends with the pair's inverse
//            geotransform (IEEE binary64, no FMA contraction), derive the contiguous active row range
//            [ya, yb] = { y : y_low <= y + 0.5 < y_high }; a warp prefix sum flattens (edge, row)
//            crossings over the lanes; each lane evaluates GDAL's intersect expression, rounds
//            floor(x + 0.5) and toggles
//   prefix   lane per row: toggle words -> inside words (in place), clipped to the window columns
//   pixels   groups of 8 pixels with a non-zero mask byte are compacted into an entry list (warp scan);
//            lanes take entries: 8 interleaved pixels arrive as 64/128-bit loads, mask bits predicate the
//            shared-memory histogram atomics
//
// This translation unit is compiled with -fmad=false and uses explicit _rn intrinsics for the
// geometry: the rounding of every operation is part of the specification.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "rs_internal.h"
#include "rs_raster.cuh"

namespace rs {

constexpr int WARPS = 4;        // teams per CTA
constexpr int CTAS_PER_SM = 5;  // occupancy target (shared memory: ~10.6 KB per team)
constexpr int PPI = 8;          // pairs per work item, at most
constexpr int AREA_MAX = 6 * 65536;   // window pixels per work item (a pair above it gets an item of its own)
constexpr int VCAP = 96;        // vertices staged in shared memory per road (longer roads read L2)
constexpr int MASKW = 896;      // mask words per team
constexpr int RCMAX = 128;      // rows per mask chunk
constexpr int ENTCAP = 448;     // 8-pixel group entries queued per team
constexpr int NCHUNK = 32;      // culling chunks per road
constexpr int RINGCAP = 16;     // ring starts kept in shared memory
constexpr int ITEM_SPLIT = 1 << 30;           // item flag: the road has several items (accumulate with atomics)
// two-kernel form (emit_kernel = zonal_kernel<PxEmit> -> accum_kernel): the rasterizer writes its 8-pixel-group entries to a
// pool in HBM instead of consuming them; pool addresses are in UNITS of 16 bytes
constexpr uint32_t CHUNK_UNITS = 4096;        // units a team takes from the pool with one atomicAdd (64 KiB)
constexpr uint32_t HEAD_OVERFLOW = 0xffffffffu;   // heads[]: the pool ran out under this item, the fused kernel redoes it
#ifndef RS_EMIT_CTAS
#define RS_EMIT_CTAS 6
#endif
constexpr int EMIT_CTAS = RS_EMIT_CTAS;      // zonal_kernel<PxEmit>: no histogram, no pixel registers
constexpr int ACC_CTAS = 12;                  // accum_kernel: CTAs of 4 warps per SM (48 warps/SM, <= 40 registers)
constexpr uint32_t ROW_REWALK = 0xffffffffu;   // rowmap marker: recount this row over [0, pitch)

// ---------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------
struct F32Args {
    double nodata;
    int has_nodata;
    uint32_t *count;                      // per feature: valid in-mask pixels (count pass)
    uint32_t *cursor;                     // per feature: values written so far (write pass)
    const unsigned long long *offset;     // per feature: start of its slice of `values`
    float *values;
};

struct ExtractArgs {
    uint32_t *count;                      // per pair: in-mask pixels (count pass)
    const unsigned long long *offset;     // per pair: first row of its slice of `values` (write pass)
    uint8_t *values;                      // [total][bpp]
    int bpp;                              // bytes per pixel (channels * element size)
};

struct ZonalArgs {
    const double2 *xy;
    const int *ring_off;
    const int *road_ring_off;
    const double *road_bbox;
    const int *road_pair_off;
    const int *pair_tile;
    const int4 *items;        // (road, first pair, pairs | ITEM_SPLIT, first row or -1)
    const PairGeom *pgeom;    // per pair
    const int *n_items;       // device-resident item counts: [0] big items (front of the list), [1] small items (back)
    int items_cap;            // capacity of the item list: small item k sits at items[items_cap - 1 - k]
    const void *pixels;
    const double *gt;
    int H, W;
    int fast;                 // 1: 64/128-bit group loads are legal (W % 8 == 0, base 16-byte aligned)
    int sparse;               // 1: the pixels are in host memory (read over the host link): load only the needed pieces
    uint32_t one;             // 1, opaque to the compiler (see red_inc3)
    const int *road_slot;
    uint32_t *hist;
    uint32_t *nzero;
    uint32_t *minzero;        // optional, per slot: sum over the road's pairs of min over bands of the pair's zero-valued in-mask pixels
    uint32_t *pair_zero;      // [n_pairs][4] scratch of the row-split pairs of tall tiles (their rows live in several items)
    uint8_t *masks;
    int window_mode;
    double sk[4], so[4];
    // integer form of the 16 -> 8 bit rescale (PxU16x4Lut): out = guess +- 1, guess = (s * lut_k + lut_b) >> 32 clamped to 0..255,
    // corrected with the exact thresholds lut_lohi[band][guess] = first | last << 16 source value that maps to `guess`
    uint32_t lut_k[4];
    long long lut_b[4];
    const uint32_t *lut_lohi;
    int *work_counter;
    int *status;
    // two-kernel form
    uint32_t *pool;           // segments: 1 header unit {pixel base lo, hi, entries, next segment + 1} then the entries (4 per unit)
    uint32_t pool_units;
    uint32_t *pool_cursor;
    uint32_t *heads;          // per item-queue index: last segment + 1 of the item (0 = no entries, HEAD_OVERFLOW)
    int4 *ov_items;           // items the pool could not hold ...
    int *ov_count;            // ... and their number
    F32Args f;                // PxF32
    ExtractArgs x;            // PxExtract
};

// team-local allocation state of the emitting rasterizer (warp-uniform registers)
struct EmitState {
    uint32_t pos = 0, end = 0;      // the team's current chunk of the pool
    uint32_t prev = 0;              // last segment + 1 of the current item
    uint32_t overflow = 0;          // sticky: the pool is exhausted
};

// thread per pair: its geometry record (the road of pair p is found by bisection of the CSR offsets)
__global__ void __launch_bounds__(256) pair_geom_kernel(const int *__restrict__ road_pair_off, const int *__restrict__ pair_tile,
                                                        const double *__restrict__ road_bbox, const double *__restrict__ gt, int n_roads,
                                                        int n_pairs, int W, int H, int window_mode, int border, int tile_lo, int tile_hi,
                                                        PairGeom *__restrict__ out, int *status)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int lo = 0, hi = n_roads;                    // largest road with road_pair_off[road] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(road_pair_off + mid) <= p) lo = mid;
        else hi = mid;
    }
    PairGeom g;
    g.inv0 = g.inv1 = g.inv3 = g.inv5 = 0.0;
    g.col_off = g.row_off = g.w = g.h = g.xshift = g.yshift = g.wu = 0;
    const int t = pair_tile[p];
    if (t < tile_lo || t >= tile_hi) g.status = 0;          // streaming: this launch only sees the tiles of one chunk
    else g.status = pair_geometry(gt + 6 * (size_t)t, road_bbox + 4 * (size_t)lo, W, H, window_mode, border, g);
    if (g.status < 0) atomicMin(status, g.status);
    out[p] = g;
}

int launch_pair_geom(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode, int border,
                     int tile_lo, int tile_hi, cudaStream_t st)
{
    if (pairs->n_pairs <= 0) return RS_OK;
    int rc = ensure(ctx, ctx->pgeom, (size_t)pairs->n_pairs * sizeof(PairGeom));
    if (rc) return rc;
    pair_geom_kernel<<<(pairs->n_pairs + 255) / 256, 256, 0, st>>>(pairs->road_pair_off, pairs->pair_tile, roads->road_bbox, tiles->gt,
                                                                   roads->n_roads, pairs->n_pairs, tiles->width, tiles->height,
                                                                   window_mode, border, tile_lo, tile_hi, (PairGeom *)ctx->pgeom.p,
                                                                   ctx->d_status);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// pixel policies: a group is 8 consecutive pixels = NW 32-bit words in registers
// ---------------------------------------------------------------------------------------------
template <int OFF, int N>
__device__ __forceinline__ uint32_t byte_at(const uint32_t (&r)[N]) { return (r[OFF >> 2] >> ((OFF & 3) * 8)) & 255u; }
template <int OFF, int N>
__device__ __forceinline__ uint32_t half_at(const uint32_t (&r)[N]) { return (r[OFF >> 2] >> ((OFF & 3) * 8)) & 0xffffu; }

// three (two, one) histogram increments under one predicate: straight-line code, the mask bit only
// predicates the atomics (no divergent branch per pixel)
// The addend of the histogram increments is the kernel argument ZonalArgs::one (= 1): with a literal 1
// ptxas turns every increment into a warp-aggregated ATOMS.POPC.INC, which needs a reconvergence
// point (BSSY / BRA / BSYNC) per atomic and cannot be predicated.
__device__ __forceinline__ void red_inc3(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t v)
{
    // shared atomics cannot be predicated: v is the pixel's mask bit (0 / 1, derived from ZonalArgs::one), masked-off
    // pixels add 0
    asm volatile(
        "red.shared.add.u32 [%0], %3;\n\t"
        "red.shared.add.u32 [%1], %3;\n\t"
        "red.shared.add.u32 [%2], %3;"
        ::"r"(a0), "r"(a1), "r"(a2), "r"(v)
        : "memory");
}
__device__ __forceinline__ void red_inc1(uint32_t a0, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a0), "r"(v) : "memory");
}
// byte K of w, times 4 (a histogram bin's byte offset)
__device__ __forceinline__ uint32_t bin_off(uint32_t w, int k)      // k is a compile-time constant after unrolling
{
    return k == 0 ? (w << 2) & 0x3fcu : (w >> (8 * k - 2)) & 0x3fcu;
}

template <int C_>
struct PxBandsU8 {
    static constexpr int C = C_, HC = C_, BPP = C_, NW = 2 * C_;
    static constexpr bool MASK = false, EMIT = false, FLT = false, EXTRACT = false;
    // pixel I of the group: hist (shared-memory byte address of the team histogram) gets one increment per band
    // `on` is the pixel's mask bit (0 / 1), also the addend of its increments
    template <int I>
    __device__ static __forceinline__ void pixel(const ZonalArgs &, const uint32_t (&r)[NW], uint32_t on, uint32_t hist, uint32_t, uint32_t &nz)
    {
        // the team histogram is 1 KiB aligned: band base | bin offset is one LOP3 after the shift
        uint32_t ad[C];
#pragma unroll
        for (int c = 0; c < C; c++) ad[c] = (hist + 1024 * c) | bin_off(r[(I * C + c) >> 2], (I * C + c) & 3);
        if constexpr (C == 3) red_inc3(ad[0], ad[1], ad[2], on);
        else {
#pragma unroll
            for (int c = 0; c < C; c++) red_inc1(ad[c], on);
        }
        // all bands zero: a mask test on the raw words (the pixel's C bytes start at byte I * C)
        constexpr int B0 = I * C, K = B0 & 3, NB = K + C;          // bytes K .. NB-1 of the two words at r[B0 >> 2]
        constexpr uint32_t M0 = (NB >= 4 ? 0xffffffffu : ((1u << (8 * NB)) - 1u)) & ~((1u << (8 * K)) - 1u);
        uint32_t any = r[B0 >> 2] & M0;
        if constexpr (NB > 4) any |= r[(B0 >> 2) + 1] & ((1u << (8 * (NB - 4))) - 1u);
        nz += (any == 0) ? on : 0u;
    }
    __device__ static __forceinline__ void pixel_slow(const ZonalArgs &a, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint8_t *p = (const uint8_t *)a.pixels + pix * C;
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const uint32_t v = __ldg(p + c);
            atomicAdd(&hist[c * 256 + v], 1u);
            any |= v;
        }
        nz += (any == 0);
    }
};

struct PxClassScore {
    static constexpr int C = 2, HC = 3, BPP = 2, NW = 4;
    static constexpr bool MASK = false, EMIT = false, FLT = false, EXTRACT = false;
    __device__ static __forceinline__ void one(uint32_t cls, uint32_t score, uint32_t *hist, uint32_t &nz)
    {
        nz += ((cls | score) == 0);
        if (cls > 2u) cls = 0u;   // unknown class codes count as "no detection"
        atomicAdd(&hist[cls * 256 + score], 1u);
    }
    template <int I>
    __device__ static __forceinline__ void pixel(const ZonalArgs &, const uint32_t (&r)[NW], uint32_t on, uint32_t hist, uint32_t one, uint32_t &nz)
    {
        uint32_t cls = byte_at<2 * I>(r);
        const uint32_t score = byte_at<2 * I + 1>(r);
        nz += ((cls | score) == 0) ? on : 0u;
        if (cls > 2u) cls = 0u;   // unknown class codes count as "no detection"
        red_inc1(hist + 4u * (cls * 256u + score), on);
    }
    __device__ static __forceinline__ void pixel_slow(const ZonalArgs &a, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint8_t *p = (const uint8_t *)a.pixels + pix * 2;
        one(__ldg(p), __ldg(p + 1), hist, nz);
    }
};

template <bool F32>
struct PxU16x4Rescale {
    static constexpr int C = 4, HC = 4, BPP = 8, NW = 16;
    static constexpr bool MASK = false, EMIT = false, FLT = false, EXTRACT = false;
    __device__ static __forceinline__ uint32_t scale(const ZonalArgs &a, uint32_t s, int c)
    {
        if (F32) {
            float f = __fadd_rn(__fmul_rn((float)s, (float)a.sk[c]), (float)a.so[c]);
            f = fminf(fmaxf(f, 0.0f), 255.0f);
            return (uint32_t)(int)__fadd_rn(f, 0.5f);
        }
        double f = __dadd_rn(__dmul_rn((double)s, a.sk[c]), a.so[c]);
        f = fmin(fmax(f, 0.0), 255.0);
        return (uint32_t)(int)__dadd_rn(f, 0.5);
    }
    template <int I>
    __device__ static __forceinline__ void pixel(const ZonalArgs &a, const uint32_t (&r)[NW], uint32_t on, uint32_t hist, uint32_t one, uint32_t &nz)
    {
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t o = scale(a, (r[(I * 8 + 2 * c) >> 2] >> (((I * 8 + 2 * c) & 3) * 8)) & 0xffffu, c);
            red_inc1(hist + 4u * (c * 256u + o), on);
            any |= o;
        }
        nz += (any == 0) ? on : 0u;
    }
    __device__ static __forceinline__ void pixel_slow(const ZonalArgs &a, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint16_t *p = (const uint16_t *)a.pixels + pix * 4;
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t o = scale(a, __ldg(p + c), c);
            atomicAdd(&hist[c * 256 + o], 1u);
            any |= o;
        }
        nz += (any == 0);
    }
};

// the same rescale without floating point in the pixel loop: gdal.Translate's  byte(clamp(s k + off, 0, 255) + 0.5)  is a
// non-decreasing step function of the 16-bit source value, so it is fixed by the 255 source values at which it steps.  The
// launcher tabulates them per band with the exact arithmetic of PxU16x4Rescale (either precision), checks on ALL 65 536 inputs
// that a 32.32 fixed-point guess is never more than one step off, and only then selects this policy: bit-identical by
// construction, four integer instructions and one L1-resident table read per band byte instead of the FP64 / XU pipe.
__shared__ uint32_t g_lut_s[4 * 256];        // PxU16x4Lut: the CTA's copy of the thresholds (static: in front of the team memory)

struct PxU16x4Lut {
    static constexpr int C = 4, HC = 4, BPP = 8, NW = 16;
    static constexpr bool MASK = false, EMIT = false, FLT = false, EXTRACT = false, LUT = true;
    __device__ static __forceinline__ uint32_t scale(const ZonalArgs &a, uint32_t s, int c)
    {
        const long long t = (long long)((unsigned long long)s * a.lut_k[c]) + a.lut_b[c];
        const int g = min(max((int)(t >> 32), 0), 255);
        const uint32_t lh = g_lut_s[c * 256 + g];
        return (uint32_t)(g + (s > (lh >> 16) ? 1 : 0) - (s < (lh & 0xffffu) ? 1 : 0));
    }
    template <int I>
    __device__ static __forceinline__ void pixel(const ZonalArgs &a, const uint32_t (&r)[NW], uint32_t on, uint32_t hist, uint32_t one, uint32_t &nz)
    {
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t o = scale(a, (r[(I * 8 + 2 * c) >> 2] >> (((I * 8 + 2 * c) & 3) * 8)) & 0xffffu, c);
            red_inc1(hist + 4u * (c * 256u + o), on);
            any |= o;
        }
        nz += (any == 0) ? on : 0u;
    }
    __device__ static __forceinline__ void pixel_slow(const ZonalArgs &a, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint16_t *p = (const uint16_t *)a.pixels + pix * 4;
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t o = scale(a, __ldg(p + c), c);
            atomicAdd(&hist[c * 256 + o], 1u);
            any |= o;
        }
        nz += (any == 0);
    }
};

struct PxMask {
    static constexpr int C = 1, HC = 0, BPP = 1, NW = 2;
    static constexpr bool MASK = true, EMIT = false, FLT = false, EXTRACT = false;
};
struct PxEmit {                     // the rasterizer of the two-kernel form: entries go to the pool
    static constexpr int C = 1, HC = 0, BPP = 1, NW = 2;
    static constexpr bool MASK = true, EMIT = true, FLT = false, EXTRACT = false;
};
// float32 single-band rasters (rasterstats.zonal_stats over a DEM, fct_rasters.py:147-163): the in-mask pixels that are neither
// NaN nor nodata are counted per feature (WRITE = false) or written to the feature's slice of a compact value array
template <bool WRITE_>
struct PxF32 {
    static constexpr int C = 1, HC = 0, BPP = 4, NW = 8;
    static constexpr bool MASK = true, EMIT = false, FLT = true, EXTRACT = false, WRITE = WRITE_;
};
// ordered extraction (fct_misc.get_pixel_values' return value, fct_misc.py:87-99): the in-mask pixels of every pair, row-major,
// counted per pair (WRITE = false) or copied behind the pair's offset (WRITE = true).  A pair is ONE work item here (no row
// slices), so a team walks it in raster order and a running position is all the ordering needs -- no P x H x W mask.
template <bool WRITE_>
struct PxExtract {
    static constexpr int C = 1, HC = 0, BPP = 1, NW = 2;
    static constexpr bool MASK = true, EMIT = false, FLT = false, EXTRACT = true, WRITE = WRITE_;
};

// 8 pixels x BPP bytes from a (8*BPP)-byte aligned address into NW words
template <int BPP, int NW>
__device__ __forceinline__ void load_group(const uint8_t *p, uint32_t (&r)[NW])
{
#ifdef RS_EXP_NOLOAD      // experiment only: how much of the kernel time is pixel-load latency?
#pragma unroll
    for (int i = 0; i < NW; i++) r[i] = (uint32_t)(uintptr_t)p * 2654435761u + i;
    return;
#endif
    if constexpr ((BPP & 1) == 0) {
#pragma unroll
        for (int i = 0; i < NW / 4; i++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
            r[4 * i] = q.x; r[4 * i + 1] = q.y; r[4 * i + 2] = q.z; r[4 * i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW / 2; i++) {
            const uint2 q = __ldg(reinterpret_cast<const uint2 *>(p) + i);
            r[2 * i] = q.x; r[2 * i + 1] = q.y;
        }
    }
}

template <class PX, int I>
__device__ __forceinline__ void group_pixels(const ZonalArgs &a, const uint32_t (&r)[PX::NW], uint32_t m8, uint32_t hist, uint32_t one,
                                             uint32_t &nz)
{
    if constexpr (I < 8) {
        PX::template pixel<I>(a, r, (m8 >> I) & one, hist, one, nz);        // one == 1, opaque to the compiler
        group_pixels<PX, I + 1>(a, r, m8, hist, one, nz);
    }
}

// ---------------------------------------------------------------------------------------------
// team shared memory
// ---------------------------------------------------------------------------------------------
struct EdgeParams {                 // the 32 edges of one block, written by their lanes, read by the crossing lanes
    double dx1[32], dy1[32], a[32], b[32], rb[32];
    int ya[32];
    int n[32];
    int off[33];
};
template <int HC, bool SPARSE = false>
struct alignas(1024) TeamSmem {
    alignas(1024) uint32_t hist[HC > 0 ? HC * 256 : 4];
    alignas(16) double2 verts[VCAP];
    alignas(16) uint32_t mask[MASKW];
    alignas(16) uint32_t rowmap[RCMAX];
    union {                         // the edge pass and the pixel phase never overlap
        EdgeParams e;
        uint32_t entries[ENTCAP];
    } u;
    float cb_ymin[NCHUNK], cb_ymax[NCHUNK], cb_xmin[NCHUNK], cb_xmax[NCHUNK];
    int ring_start[RINGCAP + 1];
    alignas(8) uint64_t mbar;
    // SPARSE (tiles read in place from host memory): one round of 8-byte pieces, at most 3 per entry
    alignas(8) unsigned long long paddr[SPARSE ? 96 : 1];      // address of every needed piece, in entry order
    alignas(8) uint2 stage[SPARSE ? 96 : 1];                    // the loaded pieces, slot = 3 * lane of the entry + piece
    uint8_t pdst[SPARSE ? 96 : 1];
};

// ---------------------------------------------------------------------------------------------
// one work item
// ---------------------------------------------------------------------------------------------
template <class PX, bool FAST, bool SPARSE>
__device__ __forceinline__ void process_item(const ZonalArgs &a, TeamSmem<PX::HC, SPARSE> &s, const int4 item, const int lane,
                                             uint32_t &mbar_phase, EmitState &es, const int qidx)
{
    if constexpr (PX::EMIT) {
        if (es.overflow) {          // nothing can be emitted any more: hand the item to the fused kernel right away
            if (lane == 0) {
                a.heads[qidx] = HEAD_OVERFLOW;
                a.ov_items[atomicAdd(a.ov_count, 1)] = item;
            }
            return;
        }
        es.prev = 0;
    }
    const uint32_t hist_addr = smem_u32(s.hist), one = a.one;
    const int road = item.x, pb = item.y, pe = item.y + (item.z & ~ITEM_SPLIT), row_first = item.w;
    const bool split = (item.z & ITEM_SPLIT) != 0;
    const int g0 = a.road_ring_off[road], g1 = a.road_ring_off[road + 1];
    const int v0 = a.ring_off[g0], nv = a.ring_off[g1] - v0;
    const int nrings = g1 - g0;
    const bool staged = nv > 0 && nv <= VCAP;

    if (staged && lane == 0) {
        const uint32_t bytes = (uint32_t)nv * 16u;
        mbar_arrive_expect_tx(&s.mbar, bytes);
        tma_bulk_g2s(s.verts, a.xy + v0, bytes, &s.mbar);
    }
    if constexpr (!PX::MASK) {
        for (int i = lane; i < PX::HC * 64; i += 32) reinterpret_cast<uint4 *>(s.hist)[i] = make_uint4(0, 0, 0, 0);
    }
    if (nrings > 1 && nrings <= RINGCAP)
        for (int k = lane; k <= nrings; k += 32) s.ring_start[k] = a.ring_off[g0 + k] - v0;
    uint32_t nz = 0, mz = 0, fcnt = 0, ecnt = 0;
    unsigned long long epos = 0;
    uint32_t zprev[PX::MASK ? 1 : PX::HC];          // hist[band][0] after the previous pair (a.minzero only)
#pragma unroll
    for (int c = 0; c < (PX::MASK ? 1 : PX::HC); c++) zprev[c] = 0;
    if (staged) {
        mbar_wait(&s.mbar, mbar_phase);
        mbar_phase ^= 1u;
    }
    __syncwarp();

    auto vertex = [&](int i) -> double2 { return staged ? s.verts[i] : __ldg(&a.xy[v0 + i]); };
    // previous vertex of i along its ring (GDAL: the first index of a ring pairs with its last)
    auto prev_index = [&](int i) -> int {
        if (nrings == 1) return i == 0 ? nv - 1 : i - 1;
        if (nrings <= RINGCAP) {
            int p = i - 1;
            for (int k = 0; k < nrings; k++)
                if (s.ring_start[k] == i) p = s.ring_start[k + 1] - 1;
            return p;
        }
        int lo = g0, hi = g1;       // ring containing vertex v0 + i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.ring_off[mid] - v0 <= i) lo = mid;
            else hi = mid;
        }
        const int rs_ = a.ring_off[lo] - v0, re_ = a.ring_off[lo + 1] - v0;
        return i == rs_ ? re_ - 1 : i - 1;
    };

    // ---------------- culling chunks: bounds over the edges (i, prev(i)) of each vertex chunk ----------------
    int cshift = 5;
    while (((nv + (1 << cshift) - 1) >> cshift) > NCHUNK) cshift++;
    const int nchunks = (nv + (1 << cshift) - 1) >> cshift;
    for (int c = 0; c < nchunks; c++) {
        float ymin = INFINITY, ymax = -INFINITY, xmin = INFINITY, xmax = -INFINITY;
        const int cend = min(nv, (c + 1) << cshift);
        for (int i = (c << cshift) + lane; i < cend; i += 32) {
            const double2 q2 = vertex(i), q1 = vertex(prev_index(i));
            ymin = fminf(ymin, __double2float_rd(fmin(q1.y, q2.y)));
            ymax = fmaxf(ymax, __double2float_ru(fmax(q1.y, q2.y)));
            xmin = fminf(xmin, __double2float_rd(fmin(q1.x, q2.x)));
            xmax = fmaxf(xmax, __double2float_ru(fmax(q1.x, q2.x)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ymin = fminf(ymin, __shfl_xor_sync(FULL, ymin, o));
            ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
            xmin = fminf(xmin, __shfl_xor_sync(FULL, xmin, o));
            xmax = fmaxf(xmax, __shfl_xor_sync(FULL, xmax, o));
        }
        if (lane == 0) { s.cb_ymin[c] = ymin; s.cb_ymax[c] = ymax; s.cb_xmin[c] = xmin; s.cb_xmax[c] = xmax; }
    }
    __syncwarp();


    for (int p = pb; p < pe && nv > 0; p++) {
        const int t = a.pair_tile[p];
        PairGeom g;
        {
            const int4 *gp = reinterpret_cast<const int4 *>(a.pgeom + p);        // same address in every lane: broadcast loads
            const int4 q0 = __ldg(gp), q1 = __ldg(gp + 1), q2 = __ldg(gp + 2), q3 = __ldg(gp + 3);
            g.inv0 = __hiloint2double(q0.y, q0.x); g.inv1 = __hiloint2double(q0.w, q0.z);
            g.inv3 = __hiloint2double(q1.y, q1.x); g.inv5 = __hiloint2double(q1.w, q1.z);
            g.col_off = q2.x; g.row_off = q2.y; g.w = q2.z; g.h = q2.w;
            g.xshift = q3.x; g.yshift = q3.y; g.wu = q3.z; g.status = q3.w;
        }
        if (g.status <= 0) continue;
        if constexpr (PX::EXTRACT) {
            ecnt = 0;
            if constexpr (PX::WRITE) epos = a.x.offset[p];
        }
        const int cbcol = g.col_off & ~31;                              // absolute column of mask bit 0
        const int pitch = ((g.col_off + g.w - 1) >> 5) - (cbcol >> 5) + 1;   // mask words per row
        const int pp = pitch | 1;                                       // odd row stride: lane-per-row walks are conflict-free
        const int lo = g.col_off - cbcol;                               // mask bit of window column 0
        const int rcmax = min((int)RCMAX, (int)MASKW / pp);

        // lane c holds the window-pixel bounds of culling chunk c (float bounds rounded outwards, one pixel of slack)
        float cy_lo = 1.0f, cy_hi = 0.0f;
        if (lane < nchunks) {
            const double ya_ = __dadd_rn(g.inv3, __dmul_rn((double)s.cb_ymin[lane], g.inv5));
            const double yb_ = __dadd_rn(g.inv3, __dmul_rn((double)s.cb_ymax[lane], g.inv5));
            const double xa_ = __dadd_rn(g.inv0, __dmul_rn((double)s.cb_xmin[lane], g.inv1));
            const double xb_ = __dadd_rn(g.inv0, __dmul_rn((double)s.cb_xmax[lane], g.inv1));
            if (fmin(xa_, xb_) - 1.0 <= (double)(g.xshift + g.w)) {       // chunks right of the window toggle nothing
                cy_lo = __double2float_rd(fmin(ya_, yb_) - 1.0 - (double)g.yshift);
                cy_hi = __double2float_ru(fmax(ya_, yb_) + 1.0 - (double)g.yshift);
            }
        }

        // rows of this item: the whole window, or ROWS_ITEM rows of a tall one
        const int rbeg = row_first < 0 ? 0 : row_first, rend = row_first < 0 ? g.h : min(g.h, row_first + ROWS_ITEM);
        const int nrow = rend - rbeg, nchunk_r = (nrow + rcmax - 1) / rcmax;
        const int rcbal = nchunk_r > 0 ? (nrow + nchunk_r - 1) / nchunk_r : 1;                   // balanced row chunks
        for (int r0 = rbeg; r0 < rend; r0 += rcbal) {
            const int rc = min(rcbal, rend - r0);
            // chunks whose bounds reach a row of this row chunk
            const unsigned rel = __ballot_sync(FULL, cy_lo <= cy_hi && cy_hi >= (float)r0 && cy_lo <= (float)(r0 + rc));
            if (rel == 0) continue;
            // invariant: mask and rowmap are all zero here (zeroed at kernel start; the pixel phase clears what
            // it consumes)

            // ---------------- edge passes: 0 = crossings (toggles), 1 = horizontal-edge burns ----------------
            bool any_hb = false;
            int nact = 0;                                   // active edges waiting in s.u.e (warp-uniform)
            // evaluate the crossings of the queued edges: a warp prefix sum flattens (edge, row) over the lanes
            auto flush_edges = [&]() {
                if (nact == 0) return;
                __syncwarp();
                const int n = lane < nact ? s.u.e.n[lane] : 0;
                int incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                s.u.e.off[lane] = incl - n;
                if (lane == 31) s.u.e.off[32] = total;
                __syncwarp();
                int j = -1;
                for (int f = lane; f < total; f += 32) {
                    if (j < 0) {            // edge of crossing f: binary search once, then walk forward
                        j = 0;
#pragma unroll
                        for (int st = 16; st > 0; st >>= 1)
                            if (s.u.e.off[j + st] <= f) j += st;
                    } else {
                        while (s.u.e.off[j + 1] <= f) j++;
                    }
                    const int y = s.u.e.ya[j] + (f - s.u.e.off[j]);
                    const double dy = __dadd_rn(int2double_magic(y + g.yshift), 0.5);
                    const double dx1 = s.u.e.dx1[j];
                    // GDAL: intersect = (dy - dy1) * (dx2 - dx1) / (dy2 - dy1) + dx1, then floor(intersect + 0.5).
                    // The quotient is first taken through the edge's reciprocal; that differs from the
                    // correctly rounded division by a few ulp, which can only change the floor when
                    // intersect + 0.5 is within 1e-4 of an integer -- those crossings (and absurdly
                    // large coordinates) take the exact division.
                    const double num = __dmul_rn(__dsub_rn(dy, s.u.e.dy1[j]), s.u.e.a[j]);
                    const double qf = __dmul_rn(num, s.u.e.rb[j]);
                    double v = __dadd_rn(__dadd_rn(qf, dx1), 0.5);
                    double t;
                    int ti = rint_magic(fmin(fmax(v, -1.0e9), 1.0e9), t);
                    if (!(fabs(qf) < 1.0e9) || !(fabs(v) < 1.0e9) || fabs(__dsub_rn(v, t)) < 1.0e-4) {
                        v = __dadd_rn(__dadd_rn(__ddiv_rn(num, s.u.e.b[j]), dx1), 0.5);
                        v = fmin(fmax(v, -1.0e9), 1.0e9);          // order-preserving: far outside either way
                        ti = rint_magic(v, t);
                    }
                    const int fl = ti - (__dsub_rn(v, t) < 0.0 ? 1 : 0) - g.xshift;     // floor(intersect + 0.5), visible column
                    // crossings at or beyond the right edge toggle nothing (xor / or with 0, no branch)
                    const int bit = lo + max(fl, 0);
                    const int word = min(bit >> 5, pitch - 1);
                    const uint32_t onbit = fl < g.w ? 1u : 0u;
                    atomicXor(&s.mask[(y - r0) * pp + word], onbit << (bit & 31));
                    atomicOr(&s.rowmap[y - r0], onbit << min(word, 31));
                }
                __syncwarp();
                nact = 0;
            };
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1 && !any_hb) break;
                for (unsigned rm = rel; rm; rm &= rm - 1) {
                    const int c = __ffs(rm) - 1;
                    const int cend = min(nv, (c + 1) << cshift);
                    for (int base = c << cshift; base < cend; base += 32) {
                        const int i = base + lane;
                        int n = 0, ya = 0;
                        bool hb = false;
                        double x1 = 0, y1 = 0, x2 = 0, y2 = 0;
                        if (i < cend) {
                            const double2 q2 = vertex(i), q1 = vertex(prev_index(i));
                            x1 = __dadd_rn(g.inv0, __dmul_rn(q1.x, g.inv1));
                            y1 = __dadd_rn(g.inv3, __dmul_rn(q1.y, g.inv5));
                            x2 = __dadd_rn(g.inv0, __dmul_rn(q2.x, g.inv1));
                            y2 = __dadd_rn(g.inv3, __dmul_rn(q2.y, g.inv5));
                            if (y1 == y2) {
                                // horizontal edge: burnt separately iff it lies exactly on a scanline of this
                                // chunk and runs towards -x
                                const double fy = floor(y1);
                                hb = (x1 > x2) && (fy + 0.5 == y1) && fy >= (double)(r0 + g.yshift) && fy < (double)(r0 + g.yshift + rc);
                            } else if (fmin(x1, x2) <= (double)(g.xshift + g.w) + 1.0) {
                                ya = max(first_row_ge(fmin(y1, y2)) - g.yshift, r0);
                                const int yb = min(last_row_lt(fmax(y1, y2)) - g.yshift, r0 + rc - 1);
                                n = max(yb - ya + 1, 0);
                            }
                        }
                        if (pass == 1) {
                            if (hb) {
                                const double hx1 = floor(__dadd_rn(x2, 0.5)), hx2 = floor(__dadd_rn(x1, 0.5));
                                if (!(hx1 > (double)(g.wu - 1) || hx2 <= 0.0)) {
                                    const int xa = max((int)fmax(hx1, 0.0) - g.xshift, 0);
                                    const int xb = min((int)fmin(hx2 - 1.0, (double)(g.wu - 1)) - g.xshift, g.w - 1);
                                    const int row = (int)floor(y1) - g.yshift - r0;
                                    if (xa <= xb) {
                                        for (int k = (lo + xa) >> 5; k <= ((lo + xb) >> 5); k++) {
                                            const int b0 = max(lo + xa - 32 * k, 0), b1 = min(lo + xb - 32 * k, 31);
                                            const uint32_t bits = (b1 >= 31 ? FULL : ((1u << (b1 + 1)) - 1u)) & ~((1u << b0) - 1u);
                                            atomicOr(&s.mask[row * pp + k], bits);
                                        }
                                        s.rowmap[row] = ROW_REWALK;
                                    }
                                }
                            }
                            continue;
                        }
                        any_hb |= __any_sync(FULL, hb);
                        // active edges (n > 0) are compacted into the 32 parameter slots; the slots are flushed
                        // (crossings evaluated) when the next block does not fit
                        const unsigned act = __ballot_sync(FULL, n > 0);
                        if (!act) continue;
                        if (nact + __popc(act) > 32) flush_edges();
                        if (n > 0) {
                            const int slot = nact + __popc(act & ((1u << lane) - 1u));
                            double dx1, dy1, dx2, dy2;
                            if (y1 < y2) { dx1 = x1; dy1 = y1; dx2 = x2; dy2 = y2; }
                            else         { dx1 = x2; dy1 = y2; dx2 = x1; dy2 = y1; }
                            const double eb = __dsub_rn(dy2, dy1);
                            s.u.e.dx1[slot] = dx1; s.u.e.dy1[slot] = dy1;
                            s.u.e.a[slot] = __dsub_rn(dx2, dx1); s.u.e.b[slot] = eb;
                            s.u.e.rb[slot] = __ddiv_rn(1.0, eb);
                            s.u.e.ya[slot] = ya; s.u.e.n[slot] = n;
                        }
                        nact += __popc(act);
                    }
                }
                if (pass == 0) flush_edges();
                __syncwarp();
                if (pass == 1) break;

                // ---------------- prefix: toggle words -> inside words, lane per row ----------------
                for (int row = lane; row < rc; row += 32) {
                    const uint32_t rmw = s.rowmap[row];
                    uint32_t range = 0;
                    if (rmw) {
                        uint32_t *mrow = s.mask + row * pp;
                        const int kfirst = __ffs(rmw) - 1;
                        int klast = kfirst - 1, cnt = 0;
                        uint32_t carry = 0;
                        for (int k = kfirst; k < pitch; k++) {
                            const uint32_t tg = mrow[k];
                            uint32_t m = prefix_xor32(tg) ^ (carry ? FULL : 0u);
                            carry ^= __popc(tg) & 1u;
                            const int hi_k = lo + g.w - 32 * k;             // window columns end here
                            if (hi_k < 32) m &= (hi_k <= 0 ? 0u : ((1u << hi_k) - 1u));
                            mrow[k] = m;
                            if (m) { klast = k; cnt += __popc(nonzero_bytes(m)); }
                            if (!carry && k < 31 && (rmw >> (k + 1)) == 0) break;   // bit 31 stands for every word >= 31
                        }
                        if (klast >= kfirst) range = (uint32_t)kfirst | ((uint32_t)(klast + 1) << 8) | ((uint32_t)cnt << 16);
                    }
                    s.rowmap[row] = range;
                }
                __syncwarp();
            }

            // ---------------- pixels: rows -> queue of 8-pixel group entries -> lanes ----------------
            const size_t tile_pix = (size_t)t * a.H * a.W;
            // entries [0, n) of the queue: one lane per entry; the next round's pixels are in flight
            // while this round's histogram atomics issue
            auto consume = [&](const int n) {
                if constexpr (PX::EMIT) {
                    // one segment: header unit + entries, from the team's chunk of the pool
                    const uint32_t units = 1u + (((uint32_t)n + 3u) >> 2);
                    if (!es.overflow && es.pos + units > es.end) {
                        uint32_t b = 0;
                        if (lane == 0) b = atomicAdd(a.pool_cursor, CHUNK_UNITS);
                        b = __shfl_sync(FULL, b, 0);
                        if (b > a.pool_units || a.pool_units - b < CHUNK_UNITS) es.overflow = 1;
                        else { es.pos = b; es.end = b + CHUNK_UNITS; }
                    }
                    if (!es.overflow) {
                        uint32_t *seg = a.pool + 4 * (size_t)es.pos;
                        const size_t base = tile_pix + (size_t)(g.row_off + r0) * a.W + cbcol;
                        if (lane == 0)
                            *reinterpret_cast<uint4 *>(seg) = make_uint4((uint32_t)base, (uint32_t)(base >> 32), (uint32_t)n, es.prev);
                        for (int e = lane; e < n; e += 32) seg[4 + e] = s.u.entries[e];
                        es.prev = es.pos + 1u;
                        es.pos += units;
                    }
                } else if constexpr (PX::EXTRACT) {
                    for (int base = 0; base < n; base += 32) {
                        const int e = base + lane;
                        const uint32_t en = e < n ? s.u.entries[e] : 0u;
                        const int c = __popc(en & 255u);
                        if constexpr (!PX::WRITE) {
                            ecnt += (uint32_t)c;
                        } else {
                            int incl = c;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int u = __shfl_up_sync(FULL, incl, o);
                                if (lane >= o) incl += u;
                            }
                            const int total = __shfl_sync(FULL, incl, 31);
                            if (c) {
                                const int x8 = cbcol + 8 * (int)((en >> 8) & 0xfffu);
                                const int yabs = g.row_off + r0 + (int)(en >> 20);
                                const int bpp = a.x.bpp;
                                const uint8_t *src = (const uint8_t *)a.pixels + (tile_pix + (size_t)yabs * a.W + x8) * bpp;
                                uint8_t *dst = a.x.values + (epos + (unsigned long long)(incl - c)) * bpp;
                                for (int i = 0; i < 8; i++)
                                    if (en & (1u << i)) {
                                        for (int k = 0; k < bpp; k++) dst[k] = __ldg(src + i * bpp + k);
                                        dst += bpp;
                                    }
                            }
                            epos += (unsigned long long)total;
                        }
                    }
                } else if constexpr (PX::FLT) {
                    const float *fpx = (const float *)a.pixels;
                    for (int base = 0; base < n; base += 32) {
                        const int e = base + lane;
                        float v[8];
                        uint32_t sel = 0;
                        if (e < n) {
                            const uint32_t en = s.u.entries[e];
                            const int x8 = cbcol + 8 * (int)((en >> 8) & 0xfffu);
                            const int yabs = g.row_off + r0 + (int)(en >> 20);
                            const float *gp = fpx + tile_pix + (size_t)yabs * a.W + x8;
                            if constexpr (FAST) {
                                const float4 q0 = __ldg(reinterpret_cast<const float4 *>(gp)), q1 = __ldg(reinterpret_cast<const float4 *>(gp) + 1);
                                v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; i++) v[i] = (en & (1u << i)) ? __ldg(gp + i) : 0.0f;
                            }
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                // rasterstats: masked where array == nodata or NaN (main.py: isnodata | isnan)
                                const bool ok = (en & (1u << i)) && (v[i] == v[i]) && !(a.f.has_nodata && (double)v[i] == a.f.nodata);
                                sel |= ok ? (1u << i) : 0u;
                            }
                        }
                        const int c = __popc(sel);
                        if constexpr (!PX::WRITE) {
                            fcnt += (uint32_t)c;
                        } else {
                            int incl = c;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int u = __shfl_up_sync(FULL, incl, o);
                                if (lane >= o) incl += u;
                            }
                            uint32_t at = 0;
                            if (lane == 31 && incl > 0) at = atomicAdd(&a.f.cursor[road], (uint32_t)incl);
                            at = __shfl_sync(FULL, at, 31);
                            float *dst = a.f.values + a.f.offset[road] + at + (uint32_t)(incl - c);
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                if (sel & (1u << i)) *dst++ = v[i];
                        }
                    }
                } else if constexpr (PX::MASK) {
                    for (int e = lane; e < n; e += 32) {
                        const uint32_t en = s.u.entries[e];
                        const int x8 = cbcol + 8 * (int)((en >> 8) & 0xfffu);
                        const int yabs = g.row_off + r0 + (int)(en >> 20);
                        uint8_t *mp = a.masks + ((size_t)p * a.H + yabs) * a.W + x8;
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            if (en & (1u << i)) mp[i] = 1;
                    }
                } else {
                    auto pixel_index = [&](uint32_t en) -> size_t {
                        const int x8 = cbcol + 8 * (int)((en >> 8) & 0xfffu);
                        const int yabs = g.row_off + r0 + (int)(en >> 20);
                        return tile_pix + (size_t)yabs * a.W + x8;
                    };
                    if constexpr (FAST && SPARSE && PX::BPP == 3) {
                        // Tiles read in place from host memory.  The host link serves a fixed number of read requests per
                        // second, each up to a 128-byte line (profiles/microbench/sysmem_read.cu: ~330 M/s with 1, 2 or 4
                        // sectors), and one warp-wide load instruction makes one request per line it touches.  So the loads
                        // are issued per 8-byte PIECE, not per entry: the needed pieces of a round's entries are listed in
                        // entry order -- row-major, so the pieces of one road row sit in consecutive lanes -- and each lane
                        // loads one piece.  A row costs one request instead of one per piece instruction.  The pieces
                        // come back through shared memory to the lane that owns the entry.
                        const uint8_t *px = (const uint8_t *)a.pixels;
                        for (int base = 0; base < n; base += 32) {
                            const int e = base + lane;
                            uint32_t en = 0, need = 0;
                            const uint8_t *gp = px;
                            if (e < n) {
                                en = s.u.entries[e];
                                gp = px + pixel_index(en) * 3;
                                need = ((en & 0x07u) ? 1u : 0u) | ((en & 0x3cu) ? 2u : 0u) | ((en & 0xe0u) ? 4u : 0u);
                            }
                            const int c = __popc(need);
                            int incl = c;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int v = __shfl_up_sync(FULL, incl, o);
                                if (lane >= o) incl += v;
                            }
                            const int np = __shfl_sync(FULL, incl, 31);
                            int k = incl - c;
#pragma unroll
                            for (int pc = 0; pc < 3; pc++)
                                if ((need >> pc) & 1u) {
                                    s.paddr[k] = (unsigned long long)(uintptr_t)(gp + 8 * pc);
                                    s.pdst[k] = (uint8_t)(3 * lane + pc);
                                    k++;
                                }
                            __syncwarp();
                            for (int pi = lane; pi < np; pi += 32)
                                s.stage[s.pdst[pi]] = __ldg(reinterpret_cast<const uint2 *>((uintptr_t)s.paddr[pi]));
                            __syncwarp();
                            if (e < n) {
                                uint32_t r[PX::NW];
#pragma unroll
                                for (int pc = 0; pc < 3; pc++) {
                                    const uint2 v = s.stage[3 * lane + pc];          // pieces that were not loaded feed increments of 0
                                    r[2 * pc] = v.x; r[2 * pc + 1] = v.y;
                                }
                                group_pixels<PX, 0>(a, r, en & 255u, hist_addr, one, nz);
                            }
                            __syncwarp();
                        }
                    } else if constexpr (FAST) {
                        uint32_t rn[PX::NW];
                        uint32_t m8n = 0;
                        int e = lane;
                        if (e < n) {
                            const uint32_t en = s.u.entries[e];
                            m8n = en & 255u;
                            load_group<PX::BPP, PX::NW>((const uint8_t *)a.pixels + pixel_index(en) * PX::BPP, rn);
                        }
                        while (e < n) {
                            uint32_t r[PX::NW];
#pragma unroll
                            for (int w = 0; w < PX::NW; w++) r[w] = rn[w];
                            const uint32_t m8 = m8n;
                            e += 32;
                            if (e < n) {
                                const uint32_t en = s.u.entries[e];
                                m8n = en & 255u;
                                load_group<PX::BPP, PX::NW>((const uint8_t *)a.pixels + pixel_index(en) * PX::BPP, rn);
                            }
                            group_pixels<PX, 0>(a, r, m8, hist_addr, one, nz);
                        }
                    } else {
                        for (int e = lane; e < n; e += 32) {
                            const uint32_t en = s.u.entries[e];
                            const size_t pix = pixel_index(en);
                            for (int i = 0; i < 8; i++)
                                if (en & (1u << i)) PX::pixel_slow(a, pix + i, s.hist, nz);
                        }
                    }
                }
            };
            int nq = 0;                                      // queued entries (warp-uniform)
            for (int b0 = 0; b0 < rc;) {
                const int row = b0 + lane;
                int cnt = 0, klo = 0, khi = 0;
                if (row < rc) {
                    const uint32_t rr = s.rowmap[row];
                    if (rr == ROW_REWALK) {
                        khi = pitch;
                        const uint32_t *mrow = s.mask + row * pp;
                        for (int k = 0; k < pitch; k++) cnt += __popc(nonzero_bytes(mrow[k]));
                    } else {
                        klo = rr & 255u; khi = (rr >> 8) & 255u; cnt = rr >> 16;
                    }
                }
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                // rows whose entries fit the queue form a prefix of the lanes (at least one once the queue holds
                // less than a round: a row has at most W / 8 <= ENTCAP - 32 groups)
                const int nrows = __popc(__ballot_sync(FULL, nq + incl <= ENTCAP));
                if (nrows < min(32, rc - b0) && nq >= 32) {
                    // the queue is full: drain the full rounds, keep the partial round, retry these rows
                    const int nfull = nq & ~31;
                    consume(nfull);
                    __syncwarp();
                    const int rest = nq - nfull;
                    uint32_t keep = 0;
                    if (lane < rest) keep = s.u.entries[nfull + lane];
                    __syncwarp();
                    if (lane < rest) s.u.entries[lane] = keep;
                    __syncwarp();
                    nq = rest;
                    continue;
                }
                const int added = __shfl_sync(FULL, incl, nrows - 1);
                if (lane < nrows && row < rc) s.rowmap[row] = 0;                 // consumed: restore the all-zero invariant
                if (lane < nrows && cnt) {
                    int off = nq + incl - cnt;
                    uint32_t *mrow = s.mask + row * pp;
                    for (int k = klo; k < khi; k++) {
                        const uint32_t m = mrow[k];
                        mrow[k] = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t m8 = (m >> (8 * j)) & 255u;
                            if (m8) s.u.entries[off++] = ((uint32_t)row << 20) | ((uint32_t)(k * 4 + j) << 8) | m8;
                        }
                    }
                }
                __syncwarp();
                nq += added;
                b0 += nrows;
            }
            if (nq) consume(nq);
            __syncwarp();
        }
        if constexpr (PX::EXTRACT) {
            if constexpr (!PX::WRITE) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ecnt += __shfl_xor_sync(FULL, ecnt, o);
                if (lane == 0) a.x.count[p] = ecnt;
            }
        }
        if constexpr (!PX::MASK) {
            // get_pixel_values with tile nodata == 0 pads every band of ONE (road, tile) call up to that call's longest band
            // (fct_misc.py:95-111): the per-road total of the padding needs min over bands of the pair's zero count
            if (a.minzero) {
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int c = 0; c < PX::HC; c++) {
                    const uint32_t z = s.hist[c * 256], d = z - zprev[c];
                    zprev[c] = z;
                    m = min(m, d);
                    if (row_first >= 0 && d && lane == 0) atomicAdd(&a.pair_zero[4 * (size_t)p + c], d);
                }
                if (row_first < 0) mz += m;
            }
        }
    }

    if constexpr (PX::FLT) {
        if constexpr (!PX::WRITE) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fcnt += __shfl_xor_sync(FULL, fcnt, o);
            if (lane == 0 && fcnt) atomicAdd(&a.f.count[road], fcnt);
        }
    }
    if constexpr (PX::EMIT) {
        if (lane == 0) {
            a.heads[qidx] = es.overflow ? HEAD_OVERFLOW : es.prev;
            if (es.overflow) a.ov_items[atomicAdd(a.ov_count, 1)] = item;
        }
    }
    // ---------------- write the road's accumulators ----------------
    if constexpr (!PX::MASK) {
        __syncwarp();
        const int slot = a.road_slot ? a.road_slot[road] : road;
        uint32_t *dst = a.hist + (size_t)slot * PX::HC * 256;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(FULL, nz, o);
        if (!split) {
            for (int i = lane; i < PX::HC * 64; i += 32)
                reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(s.hist)[i];
            if (lane == 0) a.nzero[slot] = nz;
            if (lane == 0 && a.minzero) a.minzero[slot] = mz;
        } else {
            for (int i = lane; i < PX::HC * 256; i += 32) {
                const uint32_t v = s.hist[i];
                if (v) atomicAdd(&dst[i], v);
            }
            if (lane == 0 && nz) atomicAdd(&a.nzero[slot], nz);
            if (lane == 0 && mz) atomicAdd(&a.minzero[slot], mz);          // mz != 0 only when a.minzero is set
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// the kernel: persistent teams pulling items
// ---------------------------------------------------------------------------------------------
template <class PX, bool FAST, bool SPARSE = false>
__global__ void __launch_bounds__(WARPS * 32, PX::EMIT ? EMIT_CTAS : CTAS_PER_SM) zonal_kernel(const ZonalArgs a)
{
    using S = TeamSmem<PX::HC, SPARSE>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((smem_u32(smem_raw) & 1023u) != 0u) {       // team histograms must be 1 KiB aligned (band base | bin offset)
        if (threadIdx.x == 0) atomicMin(a.status, (int)RS_ERR_CUDA);
        return;
    }
    S &s = reinterpret_cast<S *>(smem_raw)[warp];
    if constexpr (std::is_same<PX, PxU16x4Lut>::value) {
        for (int i = threadIdx.x; i < 4 * 256; i += WARPS * 32) g_lut_s[i] = __ldg(a.lut_lohi + i);
        __syncthreads();
    }
    if (lane == 0) mbar_init(&s.mbar, 1);
    for (int i = lane; i < MASKW / 4; i += 32) reinterpret_cast<uint4 *>(s.mask)[i] = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < RCMAX / 4; i += 32) reinterpret_cast<uint4 *>(s.rowmap)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    uint32_t mbar_phase = 0;
    EmitState es;
    // longest items first: the tail of the dynamic queue (teams finishing their last item while the others idle) is then made
    // of the short items
    const int n_big = a.n_items[0], n_items = n_big + a.n_items[1];
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= n_items) break;
        const int at = idx < n_big ? idx : a.items_cap - 1 - (idx - n_big);
        process_item<PX, FAST, SPARSE>(a, s, __ldg(a.items + at), lane, mbar_phase, es, idx);
    }
}


// ---------------------------------------------------------------------------------------------
// two-kernel form, second kernel: the entries of an item (segments chained from heads[]) -> the road's histograms.
// A team needs nothing but its histogram (HC KiB of shared memory) and <= 40 registers: 48 warps per SM hide the
// dependent entry -> pixel -> atomic chain that the fused kernel (20 warps per SM) stalls on.
// ---------------------------------------------------------------------------------------------
template <class PX>
__global__ void __launch_bounds__(WARPS * 32, ACC_CTAS) accum_kernel(const ZonalArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((smem_u32(smem_raw) & 1023u) != 0u) {
        if (threadIdx.x == 0) atomicMin(a.status, (int)RS_ERR_CUDA);
        return;
    }
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw) + warp * PX::HC * 256;
    const uint32_t hist_addr = smem_u32(hist), one = a.one;
    const int n_big = a.n_items[0], n_items = n_big + a.n_items[1];
    const uint8_t *px = (const uint8_t *)a.pixels;
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter + 6, 1);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= n_items) break;
        const uint32_t head = __ldg(a.heads + idx);
        if (head == HEAD_OVERFLOW) continue;                         // the fused kernel redoes this item
        const int at = idx < n_big ? idx : a.items_cap - 1 - (idx - n_big);
        const int4 item = __ldg(a.items + at);
        for (int i = lane; i < PX::HC * 64; i += 32) reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        uint32_t nz = 0;
        for (uint32_t sg = head; sg != 0u;) {
            const uint32_t *seg = a.pool + 4 * (size_t)(sg - 1u);
            const uint4 h = __ldg(reinterpret_cast<const uint4 *>(seg));
            const size_t base = (size_t)h.x | ((size_t)h.y << 32);
            const int n = (int)h.z;
            sg = h.w;
            auto address = [&](uint32_t en) -> const uint8_t * {
                return px + (base + (size_t)(en >> 20) * a.W + 8u * ((en >> 8) & 0xfffu)) * PX::BPP;
            };
            // one lane per entry, software-pipelined two deep: while round k's atomics issue, the pixels of round k + 1 are in
            // flight (their entry word arrived a round ago) and the entry word of round k + 2 is being fetched
            uint32_t rn[PX::NW];
            uint32_t m8n = 0, en2 = 0;
            int e = lane;
            if (e < n) {
                const uint32_t en = __ldg(seg + 4 + e);
                m8n = en & 255u;
                load_group<PX::BPP, PX::NW>(address(en), rn);
            }
            if (e + 32 < n) en2 = __ldg(seg + 4 + e + 32);
            while (e < n) {
                uint32_t r[PX::NW];
#pragma unroll
                for (int w = 0; w < PX::NW; w++) r[w] = rn[w];
                const uint32_t m8 = m8n;
                e += 32;
                if (e < n) {
                    const uint32_t en = en2;
                    m8n = en & 255u;
                    load_group<PX::BPP, PX::NW>(address(en), rn);
                    if (e + 32 < n) en2 = __ldg(seg + 4 + e + 32);
                }
                group_pixels<PX, 0>(a, r, m8, hist_addr, one, nz);
            }
        }
        __syncwarp();
        const int road = item.x;
        const bool split = (item.z & ITEM_SPLIT) != 0;
        const int slot = a.road_slot ? a.road_slot[road] : road;
        uint32_t *dst = a.hist + (size_t)slot * PX::HC * 256;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(FULL, nz, o);
        if (!split) {
            for (int i = lane; i < PX::HC * 64; i += 32) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(hist)[i];
            if (lane == 0) a.nzero[slot] = nz;
        } else {
            for (int i = lane; i < PX::HC * 256; i += 32) {
                const uint32_t v = hist[i];
                if (v) atomicAdd(&dst[i], v);
            }
            if (lane == 0 && nz) atomicAdd(&a.nzero[slot], nz);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// item list: thread per road.  An item is up to PPI consecutive pairs of the road whose windows sum to at most AREA_MAX
// pixels; a window taller than ROWS_ITEM rows becomes ceil(h / ROWS_ITEM) single-pair items.  Rows of roads with no item
// or several items are zeroed here (their teams accumulate with atomics).
// ---------------------------------------------------------------------------------------------
constexpr int BIG_PAIRS = 4;    // an item of at least this many pairs (or a row slice of a tall window) is queued in the first wave

// Items of one road.  COUNT pass (WRITE = false): returns the number of items, n_big = how many of them are big.
// WRITE pass: big items go to items[big_base++], small ones to items[cap - 1 - small_base++].
template <bool WRITE>
__device__ __forceinline__ int road_items(const PairGeom *__restrict__ pgeom, int road, int p0, int p1, int4 *items, int big_base,
                                          int small_base, int cap, int flag, bool tall, int ppi, int &n_big)
{
    int ni = 0;
    n_big = 0;
    auto emit = [&](int first, int count, int row, bool big) {
        if (WRITE) items[big ? big_base + n_big : cap - 1 - (small_base + (ni - n_big))] = make_int4(road, first, count | flag, row);
        ni++;
        n_big += big ? 1 : 0;
    };
    if (!tall) {            // no window can exceed ROWS_ITEM rows (tile height <= ROWS_ITEM): groups of PPI pairs, no geometry reads
        for (int p = p0; p < p1; p += ppi) {
            const int cnt = min(ppi, p1 - p);
            emit(p, cnt, -1, cnt >= BIG_PAIRS);
        }
        return ni;
    }
    int gp = p0, gn = 0, garea = 0;
    auto close = [&]() {
        if (gn > 0) emit(gp, gn, -1, gn >= BIG_PAIRS || garea >= AREA_MAX / 2);
        gn = 0; garea = 0;
    };
    for (int p = p0; p < p1; p++) {
        const int st = pgeom[p].status, h = st > 0 ? pgeom[p].h : 0, area = st > 0 ? pgeom[p].w * h : 0;
        if (h > ROWS_ITEM) {
            close();
            for (int r = 0; r < h; r += ROWS_ITEM) emit(p, 1, r, true);
            gp = p + 1;
            continue;
        }
        if (gn == ppi || (gn > 0 && garea + area > AREA_MAX)) close();
        if (gn == 0) gp = p;
        gn++; garea += area;
    }
    close();
    return ni;
}

__global__ void __launch_bounds__(256) prep_items_kernel(const int *__restrict__ road_pair_off, const PairGeom *__restrict__ pgeom,
                                                         int n_roads, const int *__restrict__ road_slot, uint32_t *hist, uint32_t *nzero,
                                                         uint32_t *minzero, int hc, int4 *items, int *n_items, int items_cap, int tall,
                                                         int accumulate, int ppi)
{
    const int road = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    int p0 = 0, p1 = 0, ni = 0, nb = 0;
    if (road < n_roads) {
        p0 = road_pair_off[road]; p1 = road_pair_off[road + 1];
        ni = road_items<false>(pgeom, road, p0, p1, nullptr, 0, 0, items_cap, 0, tall != 0, ppi, nb);
    }
    // two warp scans (big, small), one atomicAdd per warp and list
    int ib = nb, is = ni - nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int vb = __shfl_up_sync(FULL, ib, o), vs = __shfl_up_sync(FULL, is, o);
        if (lane >= o) { ib += vb; is += vs; }
    }
    const int tb = __shfl_sync(FULL, ib, 31), ts = __shfl_sync(FULL, is, 31);
    int bb = 0, bs = 0;
    if (lane == 31 && tb > 0) bb = atomicAdd(n_items, tb);
    if (lane == 31 && ts > 0) bs = atomicAdd(n_items + 1, ts);
    bb = __shfl_sync(FULL, bb, 31) + ib - nb;
    bs = __shfl_sync(FULL, bs, 31) + is - (ni - nb);
    if (ni > 0) {
        int dummy;
        road_items<true>(pgeom, road, p0, p1, items, bb, bs, items_cap, (ni > 1 || accumulate) ? ITEM_SPLIT : 0, tall != 0, ppi, dummy);
    }
    if (hist && !accumulate) {
        unsigned m = __ballot_sync(FULL, road < n_roads && ni != 1);
        for (; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const int r = __shfl_sync(FULL, road, src);
            const int slot = road_slot ? road_slot[r] : r;
            uint4 *dst = reinterpret_cast<uint4 *>(hist + (size_t)slot * hc * 256);
            for (int i = lane; i < hc * 64; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
            if (lane == 0) nzero[slot] = 0;
            if (lane == 0 && minzero) minzero[slot] = 0;
        }
    }
}

// rows of a pair taller than ROWS_ITEM live in several items: their per-band zero counts met in pair_zero; thread per pair
__global__ void __launch_bounds__(256) pair_minzero_kernel(const int *__restrict__ road_pair_off, const PairGeom *__restrict__ pgeom,
                                                           const uint32_t *__restrict__ pair_zero, int n_roads, int n_pairs, int hc,
                                                           const int *__restrict__ road_slot, uint32_t *minzero)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    if (pgeom[p].status <= 0 || pgeom[p].h <= ROWS_ITEM) return;
    uint32_t m = 0xffffffffu;
    for (int c = 0; c < hc; c++) m = min(m, pair_zero[4 * (size_t)p + c]);
    if (m == 0) return;
    int lo = 0, hi = n_roads;                    // largest road with road_pair_off[road] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(road_pair_off + mid) <= p) lo = mid;
        else hi = mid;
    }
    atomicAdd(&minzero[road_slot ? road_slot[lo] : lo], m);
}

// ---------------------------------------------------------------------------------------------
// road bbox kernel (one warp per road)
// ---------------------------------------------------------------------------------------------
__global__ void road_bbox_kernel(const double2 *__restrict__ xy, const int *__restrict__ ring_off,
                                 const int *__restrict__ road_ring_off, int n_roads, double *__restrict__ out)
{
    const int road = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (road >= n_roads) return;
    const int v0 = ring_off[road_ring_off[road]], v1 = ring_off[road_ring_off[road + 1]];
    double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
    for (int i = v0 + lane; i < v1; i += 32) {
        const double2 q = xy[i];
        xmin = fmin(xmin, q.x); xmax = fmax(xmax, q.x);
        ymin = fmin(ymin, q.y); ymax = fmax(ymax, q.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if (lane == 0) {
        out[4 * (size_t)road + 0] = xmin; out[4 * (size_t)road + 1] = ymin;
        out[4 * (size_t)road + 2] = xmax; out[4 * (size_t)road + 3] = ymax;
    }
}

int launch_road_bbox(rs_ctx *ctx, const rs_roads *roads, double *out, cudaStream_t st)
{
    if (roads->n_roads == 0) return RS_OK;
    const int threads = 256, blocks = (int)(((size_t)roads->n_roads * 32 + threads - 1) / threads);
    road_bbox_kernel<<<blocks, threads, 0, st>>>((const double2 *)roads->xy, roads->ring_off, roads->road_ring_off,
                                                 roads->n_roads, out);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
template <class PX, bool FAST, bool SPARSE = false>
static int launch_fast(rs_ctx *ctx, const ZonalArgs &args, cudaStream_t st)
{
    using S = TeamSmem<PX::HC, SPARSE>;
    const size_t smem = sizeof(S) * WARPS;
    auto kern = zonal_kernel<PX, FAST, SPARSE>;
    RS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    RS_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    const unsigned grid = (unsigned)(ctx->sm_count * per_sm);     // persistent: a whole number of resident waves
    kern<<<grid, WARPS * 32, smem, st>>>(args);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}


// the two-kernel form: rasterize -> entry pool (zonal_kernel<PxEmit>), entries -> histograms (accum_kernel<PX>), then the fused
// kernel over the items the pool could not hold (normally none: every CTA of that launch exits on its first queue read)
template <class PX>
static int launch_split(rs_ctx *ctx, ZonalArgs a, cudaStream_t st)
{
    {
        using S = TeamSmem<0, false>;
        const size_t smem = sizeof(S) * WARPS;
        auto kern = zonal_kernel<PxEmit, false, false>;
        RS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        RS_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        kern<<<(unsigned)(ctx->sm_count * per_sm), WARPS * 32, smem, st>>>(a);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
    }
    {
        const size_t smem = (size_t)WARPS * PX::HC * 1024;
        auto kern = accum_kernel<PX>;
        RS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        RS_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        kern<<<(unsigned)(ctx->sm_count * per_sm), WARPS * 32, smem, st>>>(a);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
    }
    // overflow items: the fused kernel, its queue = ov_items[0 .. *ov_count)
    ZonalArgs f = a;
    f.items = a.ov_items;
    f.n_items = a.ov_count;              // {count, 0}: counters [4], [5]
    f.work_counter = a.work_counter + 7;
    return launch_fast<PX, true>(ctx, f, st);
}

template <class PX>
static int launch_one(rs_ctx *ctx, const ZonalArgs &args, cudaStream_t st)
{
    if constexpr (PX::MASK) return launch_fast<PX, false>(ctx, args, st);
    else {
        if constexpr (PX::BPP == 3)          // tiles read in place from host memory: fewer sectors over the host link
            if (args.fast && args.sparse) return launch_fast<PX, true, true>(ctx, args, st);
        if constexpr (PX::BPP != 8)          // (uint16 x 4 keeps the fused kernel: its groups do not fit accum_kernel's 40 registers)
            if (args.pool && args.fast && !args.sparse) return launch_split<PX>(ctx, args, st);
        return args.fast ? launch_fast<PX, true>(ctx, args, st) : launch_fast<PX, false>(ctx, args, st);
    }
}

int launch_zonal(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                 const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks, int window_mode,
                 cudaStream_t st)
{
    return launch_zonal_chunk(ctx, roads, tiles, pairs, prm, hist, n_allzero, masks, window_mode, 0, 0x7fffffff, 0, st);
}

// tile_lo / tile_hi: only pairs whose tile index lies in [tile_lo, tile_hi) are processed (tiles->pixels is then indexed with
// the global tile index, so a caller that holds one chunk passes chunk_base - tile_lo * tile_bytes);
// accumulate: add into hist / n_allzero (zeroed by the caller) instead of writing every row once

// ---------------------------------------------------------------------------------------------
// 16 -> 8 bit rescale tables (PxU16x4Lut): built and verified on the host for one (k, off, precision) set, cached in the context
// ---------------------------------------------------------------------------------------------
static inline int rescale_exact(unsigned s, double k, double off, bool f32)
{
    if (f32) {
        volatile float m = (float)s * (float)k;          // separate multiply and add, like __fmul_rn / __fadd_rn
        float f = m + (float)off;
        f = f < 0.0f ? 0.0f : f;
        f = f > 255.0f ? 255.0f : f;
        return (int)(f + 0.5f);
    }
    volatile double m = (double)s * k;
    double f = m + off;
    f = f < 0.0 ? 0.0 : f;
    f = f > 255.0 ? 255.0 : f;
    return (int)(f + 0.5);
}

// returns 1 when the integer form is exact for all 65 536 inputs of all 4 bands (tables uploaded, args filled), 0 otherwise
static int prepare_rescale_lut(rs_ctx *ctx, const rs_zonal_params *prm, ZonalArgs &a, cudaStream_t st)
{
    const bool f32 = prm->rescale == 2;
    bool same = ctx->lut_valid && ctx->lut_f32 == (int)f32;
    for (int c = 0; c < 4 && same; c++) same = ctx->lut_key[c] == prm->scale_k[c] && ctx->lut_key[4 + c] == prm->scale_off[c];
    if (!same) {
        ctx->lut_valid = false;
        ctx->lut_ok = false;
        // binary64 semantics: does the float32 evaluation give the same byte for EVERY 16-bit input of every band?  (True for
        // ranges whose steps stay clear of the .5 ties by more than float32's rounding, e.g. 0 .. 65535.)  Then the float32
        // policy -- the one form of this kernel that runs at the memory roofline -- computes the binary64 result.
        bool fsame = !f32;
        for (int c = 0; c < 4 && fsame; c++)
            for (unsigned sv = 0; sv < 65536u && fsame; sv++)
                fsame = rescale_exact(sv, prm->scale_k[c], prm->scale_off[c], true) == rescale_exact(sv, prm->scale_k[c], prm->scale_off[c], false);
        ctx->lut_f32_same = fsame;
        static thread_local uint32_t lohi[4 * 256];
        bool ok = true;
        for (int c = 0; c < 4 && ok; c++) {
            const double k = prm->scale_k[c], off = prm->scale_off[c];
            if (!(k > 0.0) || !(k < 1.0) || !(off == off) || fabs(off) > 1.0e9) { ok = false; break; }
            const double kf = f32 ? (double)(float)k : k, of = f32 ? (double)(float)off : off;
            ctx->lut_k[c] = (uint32_t)llrint(kf * 4294967296.0);
            ctx->lut_b[c] = (long long)llrint((of + 0.5) * 4294967296.0);
            int lo[256], hi[256];
            for (int v = 0; v < 256; v++) { lo[v] = 65536; hi[v] = -1; }
            int prev = 0;
            for (unsigned sv = 0; sv < 65536u && ok; sv++) {
                const int o = rescale_exact(sv, k, off, f32);
                if (o < prev || o > 255) ok = false;                     // not a non-decreasing step function
                prev = o;
                if ((int)sv < lo[o]) lo[o] = (int)sv;
                hi[o] = (int)sv;
            }
            // values no input maps to: an empty interval placed where the neighbours meet, so that the correction moves on
            int next_lo = 65536;
            for (int v = 255; v >= 0; v--) {
                if (hi[v] < 0) { lo[v] = next_lo > 65535 ? 65535 : next_lo; hi[v] = lo[v] - 1; }
                else next_lo = lo[v];
            }
            for (int v = 0; v < 256; v++) lohi[c * 256 + v] = (uint32_t)lo[v] | ((uint32_t)(hi[v] < 0 ? 0 : hi[v]) << 16);
            // the device expression on every input
            for (unsigned sv = 0; sv < 65536u && ok; sv++) {
                const long long t = (long long)((unsigned long long)sv * ctx->lut_k[c]) + ctx->lut_b[c];
                long long gq = t >> 32;
                const int g = gq < 0 ? 0 : (gq > 255 ? 255 : (int)gq);
                const uint32_t lh = lohi[c * 256 + g];
                const int o = g + (sv > (lh >> 16) ? 1 : 0) - (sv < (lh & 0xffffu) ? 1 : 0);
                if (o != rescale_exact(sv, k, off, f32)) ok = false;
            }
        }
        if (ok) {
            int rc = ensure(ctx, ctx->lut_dev, sizeof(lohi));
            if (rc) return 0;
            if (cudaMemcpyAsync(ctx->lut_dev.p, lohi, sizeof(lohi), cudaMemcpyHostToDevice, st) != cudaSuccess) return 0;
            if (cudaStreamSynchronize(st) != cudaSuccess) return 0;     // lohi is reused by the next call
        }
        for (int c = 0; c < 4; c++) { ctx->lut_key[c] = prm->scale_k[c]; ctx->lut_key[4 + c] = prm->scale_off[c]; }
        ctx->lut_f32 = (int)f32;
        ctx->lut_ok = ok;
        ctx->lut_valid = true;
    }
    if (!ctx->lut_ok) return 0;
    for (int c = 0; c < 4; c++) { a.lut_k[c] = ctx->lut_k[c]; a.lut_b[c] = ctx->lut_b[c]; }
    a.lut_lohi = (const uint32_t *)ctx->lut_dev.p;
    return 1;
}

static int launch_impl(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                       const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks, int window_mode,
                       int tile_lo, int tile_hi, int accumulate, const F32Args *f32, int f32_write, cudaStream_t st,
                       const ExtractArgs *ex = nullptr);

int launch_zonal_chunk(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                       const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks, int window_mode,
                       int tile_lo, int tile_hi, int accumulate, cudaStream_t st)
{
    return launch_impl(ctx, roads, tiles, pairs, prm, hist, n_allzero, masks, window_mode, tile_lo, tile_hi, accumulate, nullptr, 0, st);
}

// float32 single-band raster: count (write == 0) or write (write == 1) the valid in-mask pixels of every feature
int launch_zonal_f32(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode, double nodata,
                     int has_nodata, uint32_t *count, uint32_t *cursor, const unsigned long long *offset, float *values, int write,
                     cudaStream_t st)
{
    F32Args f{nodata, has_nodata, count, cursor, offset, values};
    return launch_impl(ctx, roads, tiles, pairs, nullptr, nullptr, nullptr, nullptr, window_mode, 0, 0x7fffffff, 0, &f, write, st);
}

// ordered extraction: count (write == 0: count[p] for every pair) or write (values behind offset[p]) the in-mask pixels
int launch_zonal_extract(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode,
                         uint32_t *count, const unsigned long long *offset, uint8_t *values, int bpp, int write, cudaStream_t st)
{
    ExtractArgs x{count, offset, values, bpp};
    return launch_impl(ctx, roads, tiles, pairs, nullptr, nullptr, nullptr, nullptr, window_mode, 0, 0x7fffffff, 0, nullptr, write, st, &x);
}

static int launch_impl(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                       const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks, int window_mode,
                       int tile_lo, int tile_hi, int accumulate, const F32Args *f32, int f32_write, cudaStream_t st,
                       const ExtractArgs *ex)
{
    if (!roads || !tiles || !pairs) return RS_ERR_INVALID_ARG;
    if (roads->n_roads < 0 || roads->n_verts < 0 || tiles->n_tiles < 0 || pairs->n_pairs < 0) return RS_ERR_INVALID_ARG;
    if (roads->n_roads == 0) return RS_OK;
    if (!roads->xy || !roads->ring_off || !roads->road_ring_off || !roads->road_bbox || !pairs->road_pair_off)
        return RS_ERR_INVALID_ARG;
    if (pairs->n_pairs > 0 && (!pairs->pair_tile || !tiles->gt)) return RS_ERR_INVALID_ARG;
    if (((uintptr_t)roads->xy & 15u) != 0) return RS_ERR_INVALID_ARG;      // TMA bulk source alignment
    if (tiles->width < 1 || tiles->height < 1) return RS_ERR_UNSUPPORTED;      // (windows wider than 2048 px: pair_geometry)
    if (window_mode != RS_WINDOW_CROP && window_mode != RS_WINDOW_FULL && window_mode != RS_WINDOW_BOUNDLESS) return RS_ERR_INVALID_ARG;
    if (prm && (prm->border_px < 0 || prm->border_px > 4096)) return RS_ERR_INVALID_ARG;

    // item list scratch: every pair can close one group and a tall window adds ceil(H / ROWS_ITEM) row items
    const size_t cap = (size_t)pairs->n_pairs * (1 + (size_t)(tiles->height + ROWS_ITEM - 1) / ROWS_ITEM) + (size_t)roads->n_roads + 1;
    if (cap > 0x7fffffffu) return RS_ERR_UNSUPPORTED;
    int rc = ensure(ctx, ctx->items, cap * sizeof(int4));
    if (rc) return rc;
    if ((rc = ensure(ctx, ctx->pgeom, (size_t)pairs->n_pairs * sizeof(PairGeom)))) return rc;

    ZonalArgs a{};
    a.xy = (const double2 *)roads->xy;
    a.ring_off = roads->ring_off;
    a.road_ring_off = roads->road_ring_off;
    a.road_bbox = roads->road_bbox;
    a.road_pair_off = pairs->road_pair_off;
    a.pair_tile = pairs->pair_tile;
    a.items = (const int4 *)ctx->items.p;
    a.pgeom = (const PairGeom *)ctx->pgeom.p;
    a.work_counter = ctx->d_counters;
    a.n_items = ctx->d_counters + 1;
    a.items_cap = (int)cap;
    a.pixels = tiles->pixels;
    a.gt = tiles->gt;
    a.H = tiles->height;
    a.W = tiles->width;
    a.road_slot = prm ? prm->road_slot : nullptr;
    a.hist = hist;
    a.nzero = n_allzero;
    a.minzero = (prm && !masks) ? prm->min_zero : nullptr;
    a.masks = masks;
    a.window_mode = window_mode;
    a.status = ctx->d_status;
    if (prm)
        for (int c = 0; c < 4; c++) { a.sk[c] = prm->scale_k[c]; a.so[c] = prm->scale_off[c]; }
    a.one = 1u;
    a.fast = (tiles->width % 8 == 0) && (((uintptr_t)tiles->pixels & 15u) == 0);
    if (f32) a.f = *f32;
    if (ex) a.x = *ex;
    if (tiles->pixels && !f32) {
        // where do the tiles live?  (the first tile this launch reads: a streamed chunk is addressed through a shifted base)
        const size_t tile_bytes = (size_t)tiles->height * tiles->width * tiles->channels * (tiles->dtype == RS_U16 ? 2 : 1);
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, (const uint8_t *)tiles->pixels + (size_t)tile_lo * tile_bytes) == cudaSuccess)
            a.sparse = attr.type == cudaMemoryTypeHost;
        else
            cudaGetLastError();
    }

    int HC = 0;
    if (!masks && !f32 && !ex) {
        if (!prm || !hist || !n_allzero || (pairs->n_pairs > 0 && !tiles->pixels)) return RS_ERR_INVALID_ARG;
        if (prm->hist_mode == RS_HIST_CLASS_SCORE) {
            if (tiles->channels != 2 || tiles->dtype != RS_U8) return RS_ERR_UNSUPPORTED;
            HC = 3;
        } else if (prm->hist_mode == RS_HIST_BANDS) {
            if (tiles->channels < 1 || tiles->channels > 4) return RS_ERR_UNSUPPORTED;
            if (tiles->dtype == RS_U16 && tiles->channels != 4) return RS_ERR_UNSUPPORTED;
            if (tiles->dtype == RS_U16 && prm->rescale != 1 && prm->rescale != 2) return RS_ERR_INVALID_ARG;
            if (tiles->dtype == RS_U16 && ((uintptr_t)tiles->pixels & 1u)) return RS_ERR_INVALID_ARG;
            if (tiles->dtype != RS_U8 && tiles->dtype != RS_U16) return RS_ERR_INVALID_ARG;
            HC = tiles->channels;
        } else
            return RS_ERR_INVALID_ARG;
    }

    if (a.minzero && prm->hist_mode != RS_HIST_BANDS) return RS_ERR_INVALID_ARG;
    if (!masks && !f32 && !ex && wide_eligible(tiles, prm, !a.sparse))
        return launch_zonal_wide(ctx, roads, tiles, pairs, prm, hist, n_allzero, window_mode, tile_lo, tile_hi, accumulate, st);
    // one context = one set of scratch buffers: a launch on another stream than the previous one waits for it
    if (ctx->scratch_used && ctx->scratch_stream != st) RS_CUDA_OK(ctx, cudaStreamWaitEvent(st, ctx->ev_scratch, 0));
    if (a.minzero && tiles->height > ROWS_ITEM && pairs->n_pairs > 0) {
        if ((rc = ensure(ctx, ctx->pair_zero, (size_t)pairs->n_pairs * 4 * sizeof(uint32_t)))) return rc;
        RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->pair_zero.p, 0, (size_t)pairs->n_pairs * 4 * sizeof(uint32_t), st));
        a.pair_zero = (uint32_t *)ctx->pair_zero.p;
    }
    // two-kernel form (rasterize -> entry pool -> accumulate), RS_ZONAL_SPLIT=1: resident tiles with 64/128-bit group loads, no
    // per-pair by-products.  Measured on the benchmark shard (profiles/README.md): the rasterizer alone takes 5.85 ms and the
    // accumulate kernel 4.45 ms (48 warps/SM, bound by shared-memory atomic wavefronts) against 9.2 ms for the fused kernel,
    // whose warps overlap the two -- so the fused kernel stays the default and this form is kept as its parity twin.
    {
        const char *env = getenv("RS_ZONAL_SPLIT");
        const bool want = env ? atoi(env) != 0 : false;
        if (want && !masks && !f32 && !ex && a.fast && !a.sparse && !a.minzero && !accumulate && pairs->n_pairs > 0) {
            size_t free_b = 0, total_b = 0;
            RS_CUDA_OK(ctx, cudaMemGetInfo(&free_b, &total_b));
            size_t units = (size_t)pairs->n_pairs * 96 + (size_t)ctx->sm_count * 32 * CHUNK_UNITS;       // ~1.5 KiB per pair
            const size_t have = ctx->pool.cap / 16;
            const char *forced = getenv("RS_ZONAL_POOL_UNITS");        // tests: a pool too small for the work (overflow path)
            bool ok = true;
            if (forced) units = (size_t)atoll(forced);
            else {
                if (units > have) {
                    const size_t afford = (free_b / 4 + ctx->pool.cap) / 16;
                    if (units > afford) units = afford;
                }
                if (units < have) units = have;
                ok = units >= (size_t)ctx->sm_count * 8 * CHUNK_UNITS;
            }
            if (units > 0xfffffff0u) units = 0xfffffff0u;
            if (ok) {
                if (units > have && (rc = ensure(ctx, ctx->pool, units * 16))) return rc;
                if ((rc = ensure(ctx, ctx->heads, cap * sizeof(uint32_t)))) return rc;
                if ((rc = ensure(ctx, ctx->ov_items, cap * sizeof(int4)))) return rc;
                a.pool = (uint32_t *)ctx->pool.p;
                a.pool_units = (uint32_t)units;
                a.pool_cursor = (uint32_t *)(ctx->d_counters + 3);
                a.heads = (uint32_t *)ctx->heads.p;
                a.ov_items = (int4 *)ctx->ov_items.p;
                a.ov_count = ctx->d_counters + 4;
            }
        }
    }
    RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(int), st));      // work counter, big items, small items, pool cursor, ...
    if ((rc = launch_pair_geom(ctx, roads, tiles, pairs, window_mode, prm ? prm->border_px : 0, tile_lo, tile_hi, st))) return rc;
    // pairs per item: PPI when the launch has plenty of items per team; a small launch (a shard of a strongly scaled job) gets
    // shorter items, so that the tail of the dynamic queue -- teams idle while the last items finish -- stays a few per cent of it.
    // About ten items per team: on an eighth of the benchmark shard (125 k pairs) 8 / 4 / 2 / 1 pairs per item take
    // 1.52 / 1.33 / 1.39 / 1.47 ms (shorter items pay the per-item set-up and the atomic row merge more often).
    int ppi = PPI;
    {
        const size_t teams = (size_t)ctx->sm_count * CTAS_PER_SM * WARPS;
        const size_t want = (size_t)pairs->n_pairs / (teams * 10);
        if (want < (size_t)PPI) ppi = want < 1 ? 1 : (int)want;
        const char *env = getenv("RS_ZONAL_PPI");
        if (env && atoi(env) >= 1 && atoi(env) <= PPI) ppi = atoi(env);
    }
    prep_items_kernel<<<(roads->n_roads + 255) / 256, 256, 0, st>>>(pairs->road_pair_off, (const PairGeom *)ctx->pgeom.p, roads->n_roads,
                                                                    a.road_slot, (masks || f32 || ex) ? nullptr : hist, n_allzero, a.minzero, HC,
                                                                    (int4 *)ctx->items.p, ctx->d_counters + 1, (int)cap,
                                                                    !ex && tiles->height > ROWS_ITEM,
                                                                    accumulate, ppi);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());

    if (ex)
        rc = f32_write ? launch_fast<PxExtract<true>, false>(ctx, a, st) : launch_fast<PxExtract<false>, false>(ctx, a, st);
    else if (f32)
        rc = f32_write ? (a.fast ? launch_fast<PxF32<true>, true>(ctx, a, st) : launch_fast<PxF32<true>, false>(ctx, a, st))
                       : (a.fast ? launch_fast<PxF32<false>, true>(ctx, a, st) : launch_fast<PxF32<false>, false>(ctx, a, st));
    else if (masks) rc = launch_one<PxMask>(ctx, a, st);
    else if (prm->hist_mode == RS_HIST_CLASS_SCORE) rc = launch_one<PxClassScore>(ctx, a, st);
    else if (tiles->dtype == RS_U16) {
        // the float32 semantics go through the float32 policy itself (at the roofline); the binary64 semantics through the same
        // policy when the launcher verified on all 65 536 inputs per band that both precisions give the same byte, else through
        // the integer thresholds (PxU16x4Lut) when those verify, else through FP64.  RS_ZONAL_LUT=0 / 1 forces plain / verified
        // forms, RS_ZONAL_F32EQ=0 skips the float32 evaluation.
        const char *env = getenv("RS_ZONAL_LUT");
        const int mode = env ? atoi(env) : (prm->rescale == 1 ? 1 : 0);
        const int lut = mode == 1 ? prepare_rescale_lut(ctx, prm, a, st) : 0;
        const char *feq = getenv("RS_ZONAL_F32EQ");           // 0: never take the verified float32 evaluation for binary64 semantics
        if (mode == 1 && prm->rescale == 1 && ctx->lut_f32_same && !(feq && atoi(feq) == 0))
            rc = launch_one<PxU16x4Rescale<true>>(ctx, a, st);
        else if (lut) rc = launch_one<PxU16x4Lut>(ctx, a, st);
        else rc = prm->rescale == 1 ? launch_one<PxU16x4Rescale<false>>(ctx, a, st) : launch_one<PxU16x4Rescale<true>>(ctx, a, st);
    }
    else
        switch (tiles->channels) {
            case 1: rc = launch_one<PxBandsU8<1>>(ctx, a, st); break;
            case 2: rc = launch_one<PxBandsU8<2>>(ctx, a, st); break;
            case 3: rc = launch_one<PxBandsU8<3>>(ctx, a, st); break;
            default: rc = launch_one<PxBandsU8<4>>(ctx, a, st); break;
        }
    if (rc) return rc;
    if (a.pair_zero) {
        pair_minzero_kernel<<<(pairs->n_pairs + 255) / 256, 256, 0, st>>>(pairs->road_pair_off, (const PairGeom *)ctx->pgeom.p, a.pair_zero,
                                                                          roads->n_roads, pairs->n_pairs, HC, a.road_slot, a.minzero);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
    }
    // the scratch (items, pair records, counters) is busy until this point of `st`: a call on another stream waits for it
    RS_CUDA_OK(ctx, cudaEventRecord(ctx->ev_scratch, st));
    ctx->scratch_stream = st;
    ctx->scratch_used = true;
    return RS_OK;
}

}  // namespace rs
