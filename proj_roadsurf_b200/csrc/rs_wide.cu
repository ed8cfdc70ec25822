// Wide-window form of the fused rasterize + zonal accumulation (sm_100a): BASELINE configs[4], 1024 px tiles under wide
// multi-lane polygons with holes and 1 k - 10 k vertices per ring.  Same arithmetic as zonal_kernel (rs_zonal.cu) -- the GDAL
// crossing expression, the even-odd bit mask, rasterio's windows -- organised for windows of up to 2048 x 2048 pixels:
//
//   * one CTA of 16 warps per SM works on one (road, tile) pair at a time; its warps pull BANDS of the window (32 rows for a
//     1024 px window) from a shared counter and do everything for their band on their own -- cull, edges, prefix, pixels -- in a
//     private 4 KiB bit mask, so the only block barriers are the two around a pair (no phase is waited for block-wide);
//   * the road's vertices are bucketed once per launch into 64-vertex chunks with bounds (wide_chunk_kernel); a band culls
//     the chunks against its rows and streams the surviving ones through the warp's two-stage TMA pipeline (cp.async.bulk +
//     mbarrier) -- no vertex is read through the generic path except ring-closing ones;
//   * crossings toggle bits (atomicXor), a lane per row turns toggles into the inside mask;
//   * pixels: the warp compacts the 16-pixel groups of its band into a queue and consumes them one group per lane, three
//     128-bit loads per group (two rounds in flight);
//   * histogram: LANE-PRIVATE copies, hist[band][bin][lane] (96 KiB for 3 bands): bank == lane, so a warp-wide shared-memory
//     atomic is always one wavefront (the team histograms of zonal_kernel take ~3.3 on random values), and the address is
//     base | (byte << 7): shift + LOP3 per band byte.  Copies are folded and added to the road's row once per pair.
//
// Replaces the same reference calls as zonal_kernel: fct_misc.py:57-123 get_pixel_values (rasterio.mask.mask :77 + np.extract
// :95) under the loop statistical_analysis.py:180-193, for the polygons prepare_data_obj_detec.py:186-191 leaves after the
// forest difference (holes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cub/device/device_scan.cuh>

#include "rs_internal.h"
#include "rs_raster.cuh"

namespace rs {

namespace {

constexpr int WT = 512;              // threads per CTA
constexpr int WWARPS = WT / 32;
constexpr int WMASKW = 1056;         // mask words of a warp's band (32 rows x 33 words; fewer rows for wider windows)
constexpr int WBAND = 32;            // rows per band, at most (a lane per row in the prefix)
constexpr int WQCAP = 160;           // entries of a warp's queue
constexpr int WCHUNK = 64;           // vertices per chunk: bounds granularity = TMA transfer
constexpr int WRELCAP = 64;          // chunks a warp culls per sweep
constexpr int WRINGCAP = 32;         // ring starts kept in shared memory
constexpr int WMAXW = 2048;          // tile width limit (65 mask words per row)
constexpr int WPG = 8;               // consecutive pairs a CTA takes at once: those of one road share the histogram and the band queue

struct WideArgs {
    const double2 *xy;
    const int *ring_off;
    const int *road_ring_off;
    const int *pair_tile;
    const int *pair_road;
    const PairGeom *pgeom;
    const int *chunk_off;            // [n_roads + 1] first chunk of every road
    const float4 *chunk_bounds;      // per chunk: ymin, ymax, xmin, xmax over its edges (i, prev(i)), rounded outwards
    const uint8_t *chunk_start;      // per chunk: 1 when one of its vertices opens a ring (its predecessor is the ring's last vertex)
    const uint8_t *pixels;
    int H, W, n_pairs;
    uint32_t one, smem_bytes;
    const int *road_slot;
    uint32_t *hist;
    uint32_t *nzero;
    int *work_counter;
    int *status;
};

// ---------------------------------------------------------------------------------------------
// per-launch precomputation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wide_count_kernel(const int *__restrict__ ring_off, const int *__restrict__ road_ring_off, int n_roads,
                                                         int *__restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_roads) return;
    const int nv = ring_off[road_ring_off[r + 1]] - ring_off[road_ring_off[r]];
    cnt[r] = (nv + WCHUNK - 1) / WCHUNK;
}

__global__ void wide_total_kernel(const int *__restrict__ cnt, int *__restrict__ off, int n_roads)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n_roads] = n_roads ? off[n_roads - 1] + cnt[n_roads - 1] : 0;
}

// warp per chunk slot: bounds over the edges (prev(i), i) of the chunk's vertices
__global__ void __launch_bounds__(256) wide_chunk_kernel(const double2 *__restrict__ xy, const int *__restrict__ ring_off,
                                                         const int *__restrict__ road_ring_off, const int *__restrict__ chunk_off,
                                                         int n_roads, int n_slots, float4 *__restrict__ bounds,
                                                         uint8_t *__restrict__ has_start)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_slots || w >= chunk_off[n_roads]) return;
    int lo = 0, hi = n_roads;                    // largest road with chunk_off[road] <= w
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(chunk_off + mid) <= w) lo = mid;
        else hi = mid;
    }
    const int road = lo, c = w - chunk_off[road];
    const int g0 = road_ring_off[road], g1 = road_ring_off[road + 1];
    const int v0 = ring_off[g0], nv = ring_off[g1] - v0;
    float ymin = INFINITY, ymax = -INFINITY, xmin = INFINITY, xmax = -INFINITY;
    bool starts = false;
    const int cend = min(nv, (c + 1) * WCHUNK);
    for (int i = c * WCHUNK + lane; i < cend; i += 32) {
        int a = g0, b = g1;                      // ring of vertex v0 + i
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (ring_off[mid] - v0 <= i) a = mid;
            else b = mid;
        }
        const int rs_ = ring_off[a] - v0, re_ = ring_off[a + 1] - v0;
        starts |= i == rs_;
        const double2 q2 = xy[v0 + i], q1 = xy[v0 + (i == rs_ ? re_ - 1 : i - 1)];
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
    starts = __any_sync(FULL, starts);
    if (lane == 0) {
        bounds[w] = make_float4(ymin, ymax, xmin, xmax);
        has_start[w] = starts ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) wide_pair_road_kernel(const int *__restrict__ road_pair_off, int n_roads, int n_pairs,
                                                             int *__restrict__ pair_road)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int lo = 0, hi = n_roads;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(road_pair_off + mid) <= p) lo = mid;
        else hi = mid;
    }
    pair_road[p] = lo;
}

// rows of the roads of this launch: zero (the pairs add into them)
__global__ void __launch_bounds__(256) wide_zero_rows_kernel(const int *__restrict__ road_slot, int n_roads, int hc, uint32_t *hist,
                                                             uint32_t *nzero)
{
    const int road = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (road >= n_roads) return;
    const int slot = road_slot ? road_slot[road] : road;
    uint4 *dst = reinterpret_cast<uint4 *>(hist + (size_t)slot * hc * 256);
    for (int i = lane; i < hc * 64; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
    if (lane == 0) nzero[slot] = 0;
}

// ---------------------------------------------------------------------------------------------
// shared memory of the CTA
// ---------------------------------------------------------------------------------------------
struct WarpSmem {                                   // private to a warp
    alignas(16) uint32_t mask[WMASKW];               // all zero between bands: the pixel phase clears what it consumes
    uint32_t rowinfo[WBAND];                         // edge phase: words toggled in the row (bit min(word, 31)); after the prefix:
                                                     // first word | (last word + 1) << 8 of the row's inside mask; 0 = empty row
    uint32_t queue[WQCAP];
    alignas(16) double2 verts[2][WCHUNK + 1];        // [stage]: a chunk, [0] = the vertex before it
    int rel[WRELCAP];
    alignas(8) uint64_t mbar[2];
};
struct WideSmem {
    WarpSmem w[WWARPS];
    int ring_start[WRINGCAP + 1];
    int band_off[WPG + 1];                           // first band item of every pair of the current group
    int s_pair, s_band;
    uint32_t s_nz;
};

template <int OFF, int N>
__device__ __forceinline__ uint32_t bin_part(const uint32_t (&r)[N])      // (byte OFF of the group) << 7
{
    constexpr int s = (OFF & 3) * 8;
    const uint32_t w = r[OFF >> 2];
    return s >= 7 ? (w >> (s - 7)) & 0x7f80u : (w << (7 - s)) & 0x7f80u;
}

// pixel I of a 16-pixel group of C bands held in r; base[c] = address of hist[c][0][lane] (32 KiB aligned histogram)
template <int C, int I>
__device__ __forceinline__ void wide_pixel(const uint32_t (&r)[4 * C], uint32_t on, const uint32_t (&base)[C], uint32_t &nz)
{
    const uint32_t ad0 = base[0] | bin_part<I * C>(r);
    uint32_t ad1 = 0, ad2 = 0;
    if constexpr (C > 1) ad1 = base[1] | bin_part<I * C + (C > 1 ? 1 : 0)>(r);
    if constexpr (C > 2) ad2 = base[2] | bin_part<I * C + (C > 2 ? 2 : 0)>(r);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ad0), "r"(on) : "memory");
    if constexpr (C > 1) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ad1), "r"(on) : "memory");
    if constexpr (C > 2) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(ad2), "r"(on) : "memory");
    constexpr int B0 = I * C, K = B0 & 3, NB = K + C;          // the pixel's bytes: K .. NB-1 of the words at r[B0 >> 2]
    constexpr uint32_t M0 = (NB >= 4 ? 0xffffffffu : ((1u << (8 * NB)) - 1u)) & ~((1u << (8 * K)) - 1u);
    uint32_t any = r[B0 >> 2] & M0;
    if constexpr (NB > 4) any |= r[(B0 >> 2) + 1] & ((1u << (8 * (NB - 4))) - 1u);
    nz += (any == 0) ? on : 0u;
}
template <int C, int I>
__device__ __forceinline__ void wide_group(const uint32_t (&r)[4 * C], uint32_t m16, uint32_t one, const uint32_t (&base)[C], uint32_t &nz)
{
    if constexpr (I < 16) {
        wide_pixel<C, I>(r, (m16 >> I) & one, base, nz);
        wide_group<C, I + 1>(r, m16, one, base, nz);
    }
}
template <int C>
__device__ __forceinline__ void wide_load(const uint8_t *p, uint32_t (&r)[4 * C])
{
#pragma unroll
    for (int i = 0; i < C; i++) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        r[4 * i] = q.x; r[4 * i + 1] = q.y; r[4 * i + 2] = q.z; r[4 * i + 3] = q.w;
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel: persistent CTAs pulling pairs
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(WT, 1) zonal_wide_kernel(const WideArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the lane-private histogram sits on a 32 KiB boundary (bin address = base | (byte << 7)); everything else lives in front of it
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t hist_addr = (raw + (uint32_t)sizeof(WideSmem) + 32767u) & ~32767u;
    if (hist_addr - raw + (uint32_t)C * 32768u > a.smem_bytes) {          // the launch assumed a lower base address than it got
        if (tid == 0) atomicMin(a.status, (int)RS_ERR_CUDA);
        return;
    }
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw + (hist_addr - raw));
    WideSmem &s = *reinterpret_cast<WideSmem *>(smem_raw);
    WarpSmem &ws = s.w[warp];
    uint32_t base[C];
#pragma unroll
    for (int c = 0; c < C; c++) base[c] = hist_addr + (uint32_t)c * 32768u + (uint32_t)lane * 4u;
    const uint32_t one = a.one;

    for (int i = tid; i < C * 8192; i += WT) hist[i] = 0;
    for (int i = lane; i < WMASKW; i += 32) ws.mask[i] = 0;
    ws.rowinfo[lane] = 0;
    if (lane == 0) {
        mbar_init(&ws.mbar[0], 1);
        mbar_init(&ws.mbar[1], 1);
    }
    if (tid == 0) { s.s_nz = 0; s.s_band = 0; }
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;                       // phases of this warp's two mbarriers (warp-uniform)

    for (;;) {
        if (tid == 0) s.s_pair = atomicAdd(a.work_counter, WPG);
        __syncthreads();
        const int pb = s.s_pair;
        if (pb >= a.n_pairs) break;
        const int pe = min(pb + WPG, a.n_pairs);
        for (int p0 = pb; p0 < pe;) {                        // a group: the consecutive pairs of one road
            const int road = a.pair_road[p0];
            int p1 = p0 + 1;
            while (p1 < pe && a.pair_road[p1] == road) p1++;
            const int g0 = a.road_ring_off[road], g1 = a.road_ring_off[road + 1];
            const int v0 = a.ring_off[g0], nv = a.ring_off[g1] - v0, nrings = g1 - g0;
            const int c_first = a.chunk_off[road], nch = a.chunk_off[road + 1] - c_first;
            __syncthreads();                                 // the previous group is folded, s_pair read by everybody
            if (nrings > 1 && nrings <= WRINGCAP)
                for (int k = tid; k <= nrings; k += WT) s.ring_start[k] = a.ring_off[g0 + k] - v0;
            if (tid == 0) {                                  // band items of the group: pair k owns [band_off[k], band_off[k + 1])
                int acc = 0;
                for (int k = 0; k < p1 - p0; k++) {
                    s.band_off[k] = acc;
                    const PairGeom &pg = a.pgeom[p0 + k];
                    if (pg.status > 0 && nv > 0) {
                        const int cb = pg.col_off & ~31, pt = ((pg.col_off + pg.w - 1) >> 5) - (cb >> 5) + 1;
                        const int rbk = min((int)WBAND, (int)WMASKW / (pt | 1));
                        acc += (pg.h + rbk - 1) / rbk;
                    }
                }
                s.band_off[p1 - p0] = acc;
                s.s_band = 0;
            }
            __syncthreads();
            const int n_items = s.band_off[p1 - p0];
            uint32_t nz = 0;
            int cur = -1;                                    // pair of the group whose geometry this warp holds
            PairGeom g;
            int cbcol = 0, pitch = 1, pp = 1, lo = 0, rb = 1;
            size_t tile_pix = 0;

            // previous vertex of i along its ring, when i starts a ring (GDAL pairs the first index of a ring with its last)
            auto ring_prev = [&](int i, bool &is_start) -> int {
                is_start = false;
                if (nrings == 1) { is_start = i == 0; return nv - 1; }
                if (nrings <= WRINGCAP) {
                    int pr = i - 1;
                    for (int k = 0; k < nrings; k++)
                        if (s.ring_start[k] == i) { pr = s.ring_start[k + 1] - 1; is_start = true; }
                    return pr;
                }
                int l = g0, h = g1;
                while (h - l > 1) {
                    const int mid = (l + h) >> 1;
                    if (a.ring_off[mid] - v0 <= i) l = mid;
                    else h = mid;
                }
                const int rs_ = a.ring_off[l] - v0;
                is_start = i == rs_;
                return a.ring_off[l + 1] - v0 - 1;
            };

            for (;;) {
                int item = 0;
                if (lane == 0) item = atomicAdd(&s.s_band, 1);
                item = __shfl_sync(FULL, item, 0);
                if (item >= n_items) break;
                int kp = 0;
                while (item >= s.band_off[kp + 1]) kp++;
                if (kp != cur) {                             // another pair of the group: its window geometry
                    cur = kp;
                    const int4 *gp = reinterpret_cast<const int4 *>(a.pgeom + p0 + kp);
                    const int4 q0 = __ldg(gp), q1 = __ldg(gp + 1), q2 = __ldg(gp + 2), q3 = __ldg(gp + 3);
                    g.inv0 = __hiloint2double(q0.y, q0.x); g.inv1 = __hiloint2double(q0.w, q0.z);
                    g.inv3 = __hiloint2double(q1.y, q1.x); g.inv5 = __hiloint2double(q1.w, q1.z);
                    g.col_off = q2.x; g.row_off = q2.y; g.w = q2.z; g.h = q2.w;
                    g.xshift = q3.x; g.yshift = q3.y; g.wu = q3.z; g.status = q3.w;
                    cbcol = g.col_off & ~31;                                        // absolute column of mask bit 0
                    pitch = ((g.col_off + g.w - 1) >> 5) - (cbcol >> 5) + 1;        // mask words per row (<= 65)
                    pp = pitch | 1;                                                 // odd row stride: lane-per-row walks are conflict-free
                    lo = g.col_off - cbcol;                                         // mask bit of window column 0
                    rb = min((int)WBAND, (int)WMASKW / pp);                         // rows per band
                    tile_pix = (size_t)a.pair_tile[p0 + kp] * a.H * a.W;
                }
                const int band = item - s.band_off[kp];
                const int r0 = band * rb, rc = min(rb, g.h - r0);
                bool any_hb = false, touched_any = false;

                // one edge of the road against this band: crossings toggle bits (pass 0); horizontal edges lying exactly on a
                // scanline and running towards -x are burnt after the prefix (pass 1)
                auto edge = [&](const double2 q1, const double2 q2, const int pass) {
                    const double x1 = __dadd_rn(g.inv0, __dmul_rn(q1.x, g.inv1));
                    const double y1 = __dadd_rn(g.inv3, __dmul_rn(q1.y, g.inv5));
                    const double x2 = __dadd_rn(g.inv0, __dmul_rn(q2.x, g.inv1));
                    const double y2 = __dadd_rn(g.inv3, __dmul_rn(q2.y, g.inv5));
                    if (y1 == y2) {
                        const double fy = floor(y1);
                        const bool hb = (x1 > x2) && (fy + 0.5 == y1) && fy >= (double)(r0 + g.yshift) && fy < (double)(r0 + g.yshift + rc);
                        if (hb && pass == 0) any_hb = true;
                        if (hb && pass == 1) {
                            const double hx1 = floor(__dadd_rn(x2, 0.5)), hx2 = floor(__dadd_rn(x1, 0.5));
                            if (!(hx1 > (double)(g.wu - 1) || hx2 <= 0.0)) {
                                const int xa = max((int)fmax(hx1, 0.0) - g.xshift, 0);
                                const int xb = min((int)fmin(hx2 - 1.0, (double)(g.wu - 1)) - g.xshift, g.w - 1);
                                const int row = (int)floor(y1) - g.yshift - r0;
                                if (xa <= xb) {
                                    for (int kk = (lo + xa) >> 5; kk <= ((lo + xb) >> 5); kk++) {
                                        const int b0 = max(lo + xa - 32 * kk, 0), b1 = min(lo + xb - 32 * kk, 31);
                                        const uint32_t bits = (b1 >= 31 ? FULL : ((1u << (b1 + 1)) - 1u)) & ~((1u << b0) - 1u);
                                        atomicOr(&ws.mask[row * pp + kk], bits);
                                    }
                                    ws.rowinfo[row] = (uint32_t)pitch << 8;          // the whole row is rescanned
                                }
                            }
                        }
                    } else if (pass == 0 && fmin(x1, x2) <= (double)(g.xshift + g.w) + 1.0) {
                        const int ya = max(first_row_ge(fmin(y1, y2)) - g.yshift, r0);
                        const int yb = min(last_row_lt(fmax(y1, y2)) - g.yshift, r0 + rc - 1);
                        if (ya <= yb) {
                            double dx1, dy1, dx2, dy2;
                            if (y1 < y2) { dx1 = x1; dy1 = y1; dx2 = x2; dy2 = y2; }
                            else         { dx1 = x2; dy1 = y2; dx2 = x1; dy2 = y1; }
                            const double ea = __dsub_rn(dx2, dx1), eb = __dsub_rn(dy2, dy1), erb = __ddiv_rn(1.0, eb);
                            for (int y = ya; y <= yb; y++) {
                                // GDAL: intersect = (dy - dy1) * (dx2 - dx1) / (dy2 - dy1) + dx1, then floor(intersect + 0.5);
                                // reciprocal first, the correctly rounded division when the floor could differ
                                const double dy = __dadd_rn(int2double_magic(y + g.yshift), 0.5);
                                const double num = __dmul_rn(__dsub_rn(dy, dy1), ea);
                                const double qf = __dmul_rn(num, erb);
                                double v = __dadd_rn(__dadd_rn(qf, dx1), 0.5);
                                double tt;
                                int ti = rint_magic(fmin(fmax(v, -1.0e9), 1.0e9), tt);
                                if (!(fabs(qf) < 1.0e9) || !(fabs(v) < 1.0e9) || fabs(__dsub_rn(v, tt)) < 1.0e-4) {
                                    v = __dadd_rn(__dadd_rn(__ddiv_rn(num, eb), dx1), 0.5);
                                    v = fmin(fmax(v, -1.0e9), 1.0e9);
                                    ti = rint_magic(v, tt);
                                }
                                const int fl = ti - (__dsub_rn(v, tt) < 0.0 ? 1 : 0) - g.xshift;
                                if (fl < g.w) {
                                    const int bit = lo + max(fl, 0), word = bit >> 5;
                                    atomicXor(&ws.mask[(y - r0) * pp + word], 1u << (bit & 31));
                                    atomicOr(&ws.rowinfo[y - r0], 1u << min(word, 31));
                                }
                            }
                        }
                    }
                };

                // the edges of the chunks in ws.rel through the warp's two-stage TMA pipeline
                auto edge_sweep = [&](const int nrel, const int pass) {
                    auto issue = [&](int k, int stg) {                  // lane 0
                        const int cs = ws.rel[k] * WCHUNK, ce = min(nv, cs + WCHUNK), from = cs > 0 ? cs - 1 : 0;
                        const uint32_t bytes = (uint32_t)(ce - from) * 16u;
                        mbar_arrive_expect_tx(&ws.mbar[stg], bytes);
                        tma_bulk_g2s(&ws.verts[stg][from - cs + 1], a.xy + v0 + from, bytes, &ws.mbar[stg]);
                    };
                    if (nrel > 0 && lane == 0) issue(0, 0);
                    for (int k = 0, stg = 0; k < nrel; k++, stg ^= 1) {
                        if (k + 1 < nrel && lane == 0) issue(k + 1, stg ^ 1);
                        mbar_wait(&ws.mbar[stg], stg ? ph1 : ph0);
                        if (stg) ph1 ^= 1u; else ph0 ^= 1u;
                        const int cs = ws.rel[k] * WCHUNK, ce = min(nv, cs + WCHUNK);
                        if (__ldg(a.chunk_start + c_first + ws.rel[k])) {       // a ring opens inside this chunk: look the predecessor up
                            for (int j = lane; cs + j < ce; j += 32) {
                                bool is_start;
                                const int pr = ring_prev(cs + j, is_start);
                                edge(is_start ? __ldg(&a.xy[v0 + pr]) : ws.verts[stg][j], ws.verts[stg][j + 1], pass);
                            }
                        } else {
                            for (int j = lane; cs + j < ce; j += 32) edge(ws.verts[stg][j], ws.verts[stg][j + 1], pass);
                        }
                        __syncwarp();                                   // the stage is free for the transfer after next
                    }
                };
                // cull: chunks whose bounds reach a row of this band (and, for the crossings, are not right of the window)
                auto sweeps = [&](const int pass) {
                    int nrel = 0;
                    for (int cb = 0; cb < nch; cb += 32) {
                        const int c = cb + lane;
                        bool rel = false;
                        if (c < nch) {
                            const float4 b = __ldg(a.chunk_bounds + c_first + c);
                            const double ya_ = __dadd_rn(g.inv3, __dmul_rn((double)b.x, g.inv5));
                            const double yb_ = __dadd_rn(g.inv3, __dmul_rn((double)b.y, g.inv5));
                            const double xa_ = __dadd_rn(g.inv0, __dmul_rn((double)b.z, g.inv1));
                            const double xb_ = __dadd_rn(g.inv0, __dmul_rn((double)b.w, g.inv1));
                            const double cy_lo = fmin(ya_, yb_) - 1.0 - (double)g.yshift, cy_hi = fmax(ya_, yb_) + 1.0 - (double)g.yshift;
                            rel = cy_hi >= (double)r0 && cy_lo <= (double)(r0 + rc) &&
                                  (pass == 1 || fmin(xa_, xb_) - 1.0 <= (double)(g.xshift + g.w));
                        }
                        const unsigned m = __ballot_sync(FULL, rel);
                        if (!m) continue;
                        if (nrel + __popc(m) > WRELCAP) {               // the list is full: work it off first
                            __syncwarp();
                            edge_sweep(nrel, pass);
                            touched_any = true;
                            nrel = 0;
                        }
                        if (rel) ws.rel[nrel + __popc(m & ((1u << lane) - 1u))] = c;
                        nrel += __popc(m);
                    }
                    __syncwarp();
                    if (nrel) { edge_sweep(nrel, pass); touched_any = true; }
                };

                sweeps(0);
                if (!touched_any) continue;
                __syncwarp();
                // ---------------- prefix: a lane per row of the band ----------------
                if (lane < rc) {
                    const uint32_t touched = ws.rowinfo[lane];
                    uint32_t range = 0;
                    if (touched) {
                        uint32_t *mrow = ws.mask + lane * pp;
                        const int kfirst = __ffs(touched) - 1;
                        int klast = kfirst - 1;
                        uint32_t carry = 0;
                        for (int k = kfirst; k < pitch; k++) {
                            const uint32_t tg = mrow[k];
                            uint32_t m = prefix_xor32(tg) ^ (carry ? FULL : 0u);
                            carry ^= __popc(tg) & 1u;
                            const int hi_k = lo + g.w - 32 * k;             // window columns end here
                            if (hi_k < 32) m &= (hi_k <= 0 ? 0u : ((1u << hi_k) - 1u));
                            mrow[k] = m;
                            if (m) klast = k;
                            if (!carry && k < 31 && (touched >> (k + 1)) == 0) break;     // bit 31 stands for every word >= 31
                        }
                        if (klast >= kfirst) range = (uint32_t)kfirst | ((uint32_t)(klast + 1) << 8);
                    }
                    ws.rowinfo[lane] = range;
                }
                __syncwarp();
                if (__any_sync(FULL, any_hb)) {                         // horizontal-edge burns: the same sweeps once more
                    sweeps(1);
                    __syncwarp();
                }

                // ---------------- pixels: the warp queues the 16-pixel groups of its rows and consumes them, a group per lane ----------------
                uint32_t *q = ws.queue;
                const uint8_t *band_px = a.pixels + (tile_pix + (size_t)(g.row_off + r0) * a.W + cbcol) * C;
                auto consume = [&](const int n) {
                    auto address = [&](uint32_t en) -> const uint8_t * {
                        return band_px + ((size_t)(en >> 24) * a.W + 16u * ((en >> 16) & 255u)) * C;
                    };
                    uint32_t rn[4 * C];
                    uint32_t mn = 0;
                    int e = lane;
                    if (e < n) {
                        const uint32_t en = q[e];
                        mn = en & 0xffffu;
                        wide_load<C>(address(en), rn);
                    }
                    while (e < n) {
                        uint32_t r[4 * C];
#pragma unroll
                        for (int w = 0; w < 4 * C; w++) r[w] = rn[w];
                        const uint32_t m16 = mn;
                        e += 32;
                        if (e < n) {
                            const uint32_t en = q[e];
                            mn = en & 0xffffu;
                            wide_load<C>(address(en), rn);
                        }
                        wide_group<C, 0>(r, m16, one, base, nz);
                    }
                };
                int nq = 0;
                for (int row = 0; row < rc; row++) {
                    const uint32_t ri = ws.rowinfo[row];
                    if (!ri) continue;
                    const int h0 = 2 * (int)(ri & 255u), h1 = 2 * (int)(ri >> 8);
                    uint32_t *mrow = ws.mask + row * pp;
                    for (int hb = h0; hb < h1; hb += 32) {
                        const int h = hb + lane;
                        const uint32_t m16 = h < h1 ? (mrow[h >> 1] >> ((h & 1) * 16)) & 0xffffu : 0u;
                        __syncwarp();
                        if (h < h1 && (h & 1)) mrow[h >> 1] = 0;            // consumed: the mask goes back to all zero
                        const unsigned bal = __ballot_sync(FULL, m16 != 0u);
                        const int cnt = __popc(bal);
                        if (nq + cnt > WQCAP) {                             // drain the full rounds, keep the rest
                            __syncwarp();
                            const int nfull = nq & ~31;
                            consume(nfull);
                            __syncwarp();
                            const int rest = nq - nfull;
                            uint32_t keep = 0;
                            if (lane < rest) keep = q[nfull + lane];
                            __syncwarp();
                            if (lane < rest) q[lane] = keep;
                            nq = rest;
                        }
                        if (m16) q[nq + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)row << 24) | ((uint32_t)h << 16) | m16;
                        nq += cnt;
                    }
                }
                __syncwarp();
                ws.rowinfo[lane] = 0;
                if (nq) consume(nq);
                __syncwarp();
            }

            // ---------------- fold the lane-private copies and add them to the road's row ----------------
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(FULL, nz, o);
            if (lane == 0 && nz) atomicAdd(&s.s_nz, nz);
            __syncthreads();                                 // every band of the group is in the histogram
            if (n_items > 0) {
                const int slot = a.road_slot ? a.road_slot[road] : road;
                uint32_t *dst = a.hist + (size_t)slot * C * 256;
                for (int bin = tid; bin < C * 256; bin += WT) {
                    uint32_t *hb = hist + bin * 32;
                    uint32_t sum = 0;
#pragma unroll 8
                    for (int jj = 0; jj < 32; jj++) {
                        const int j = (jj + tid) & 31;                 // rotated: conflict-free across the warp
                        sum += hb[j];
                        hb[j] = 0;
                    }
                    if (sum) atomicAdd(&dst[bin], sum);
                }
                if (tid == 0) {
                    if (s.s_nz) atomicAdd(&a.nzero[slot], s.s_nz);
                    s.s_nz = 0;
                }
            }
            p0 = p1;
            // (the barrier at the top of the next group orders the fold against its bands)
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
bool wide_eligible(const rs_tiles *tiles, const rs_zonal_params *prm, bool resident)
{
    const char *env = getenv("RS_ZONAL_WIDE");
    if (env && atoi(env) == 0) return false;
    if (!prm || prm->hist_mode != RS_HIST_BANDS || prm->min_zero) return false;
    if (tiles->dtype != RS_U8 || tiles->channels < 1 || tiles->channels > 3) return false;
    if (tiles->width < 512 || tiles->width > WMAXW || tiles->width % 16 != 0) return false;
    if (((uintptr_t)tiles->pixels & 15u) != 0 || !resident) return false;
    return true;
}

template <int C>
static int launch_wide_c(rs_ctx *ctx, WideArgs a, cudaStream_t st)
{
    // the per-warp state sits in front of the histogram, which starts on the next 32 KiB boundary of the shared address space;
    // dynamic shared memory begins ~1 KiB into that space (the block's reserved area), so the histogram lands at
    // round_up(sizeof(WideSmem) + ~1 KiB, 32 KiB) -- the kernel checks it
    const size_t front = ((sizeof(WideSmem) + 2048 + 32767) / 32768) * 32768;
    const size_t smem = front + (size_t)C * 32768;
    a.smem_bytes = (uint32_t)smem;
    auto kern = zonal_wide_kernel<C>;
    RS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)ctx->sm_count, WT, smem, st>>>(a);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

int launch_zonal_wide(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, const rs_zonal_params *prm,
                      uint32_t *hist, uint32_t *n_allzero, int window_mode, int tile_lo, int tile_hi, int accumulate, cudaStream_t st)
{
    const int R = roads->n_roads, P = pairs->n_pairs, C = tiles->channels;
    int rc;
    if (ctx->scratch_used && ctx->scratch_stream != st) RS_CUDA_OK(ctx, cudaStreamWaitEvent(st, ctx->ev_scratch, 0));
    const size_t n_slots = (size_t)roads->n_verts / WCHUNK + (size_t)R + 1;           // upper bound of the chunk count
    if ((rc = ensure(ctx, ctx->wide_cnt, sizeof(int) * ((size_t)R + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->wide_off, sizeof(int) * ((size_t)R + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->wide_bounds, sizeof(float4) * n_slots))) return rc;
    if ((rc = ensure(ctx, ctx->wide_flags, n_slots))) return rc;
    if ((rc = ensure(ctx, ctx->wide_pair_road, sizeof(int) * ((size_t)P + 1)))) return rc;
    size_t tmp = 0;
    RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, (int *)ctx->wide_cnt.p, (int *)ctx->wide_off.p, R, st));
    if ((rc = ensure(ctx, ctx->wide_tmp, tmp))) return rc;
    RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(int), st));
    if (!accumulate) {
        wide_zero_rows_kernel<<<(unsigned)(((size_t)R * 32 + 255) / 256), 256, 0, st>>>(prm->road_slot, R, C, hist, n_allzero);
        ctx->launches++;
    }
    if (P > 0) {
        if ((rc = launch_pair_geom(ctx, roads, tiles, pairs, window_mode, prm->border_px, tile_lo, tile_hi, st))) return rc;
        wide_count_kernel<<<(R + 255) / 256, 256, 0, st>>>(roads->ring_off, roads->road_ring_off, R, (int *)ctx->wide_cnt.p);
        RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->wide_tmp.p, tmp, (int *)ctx->wide_cnt.p, (int *)ctx->wide_off.p, R, st));
        wide_total_kernel<<<1, 32, 0, st>>>((const int *)ctx->wide_cnt.p, (int *)ctx->wide_off.p, R);
        wide_chunk_kernel<<<(unsigned)((n_slots * 32 + 255) / 256), 256, 0, st>>>((const double2 *)roads->xy, roads->ring_off,
                                                                                   roads->road_ring_off, (const int *)ctx->wide_off.p, R,
                                                                                   (int)n_slots, (float4 *)ctx->wide_bounds.p,
                                                                                   (uint8_t *)ctx->wide_flags.p);
        wide_pair_road_kernel<<<(P + 255) / 256, 256, 0, st>>>(pairs->road_pair_off, R, P, (int *)ctx->wide_pair_road.p);
        ctx->launches += 4;
        RS_CUDA_OK(ctx, cudaGetLastError());
        WideArgs a{};
        a.xy = (const double2 *)roads->xy;
        a.ring_off = roads->ring_off;
        a.road_ring_off = roads->road_ring_off;
        a.pair_tile = pairs->pair_tile;
        a.pair_road = (const int *)ctx->wide_pair_road.p;
        a.pgeom = (const PairGeom *)ctx->pgeom.p;
        a.chunk_off = (const int *)ctx->wide_off.p;
        a.chunk_bounds = (const float4 *)ctx->wide_bounds.p;
        a.chunk_start = (const uint8_t *)ctx->wide_flags.p;
        a.pixels = (const uint8_t *)tiles->pixels;
        a.H = tiles->height;
        a.W = tiles->width;
        a.n_pairs = P;
        a.one = 1u;
        a.road_slot = prm->road_slot;
        a.hist = hist;
        a.nzero = n_allzero;
        a.work_counter = ctx->d_counters;
        a.status = ctx->d_status;
        rc = C == 1 ? launch_wide_c<1>(ctx, a, st) : C == 2 ? launch_wide_c<2>(ctx, a, st) : launch_wide_c<3>(ctx, a, st);
        if (rc) return rc;
    }
    RS_CUDA_OK(ctx, cudaEventRecord(ctx->ev_scratch, st));
    ctx->scratch_stream = st;
    ctx->scratch_used = true;
    return RS_OK;
}

}  // namespace rs
