// K1+K2: fused scanline rasterization of road polygons and per-road zonal accumulation (sm_100a).
//
// Replaces, for a whole road-major pair list in one launch, the reference's per-(road, tile)
// Python loop  scripts/statistical_analysis/statistical_analysis.py:180-193  whose body is
// scripts/functions/fct_misc.py:57-123 get_pixel_values = rasterio.mask.mask(crop=True) (:77,
// GDAL GDALRasterizeGeometries -> GDALdllImageFilledPolygon) + np.extract (:95).
//
// Work decomposition: one TEAM (a warp, or a whole CTA for long edge lists) owns one road and
// folds all of that road's (tile) pairs into a TEAM-private shared-memory histogram, so every
// output row is written exactly once with plain 128-bit stores: no global atomics, no output
// clearing, results independent of scheduling.
//
// Per pair and per chunk of <= RC scanlines the algorithm is work-optimal, O(E + X + P) for E
// edges, X scanline crossings and P covered pixels (GDAL's loop is O(rows * E)):
//   pass 1   each lane takes edges; an edge is active on the contiguous row range
//            [ya, yb] = { y : y_low <= y + 0.5 < y_high } (GDAL's half-open rule), so per-row
//            crossing counts are a difference array (two shared atomics per edge) + prefix sum;
//            per-edge crossing counts are prefix-summed too (load balancing of pass 2)
//   pass 2   lanes take flattened (edge, row) crossings (binary search in the edge prefix),
//            evaluate GDAL's intersect expression in IEEE binary64 without FMA contraction,
//            round floor(x + 0.5) BEFORE sorting, and scatter into the row's slot range
//   sort     one lane per row, insertion sort of that row's (few) int16 crossings
//   spans    crossing pairs (2m, 2m+1) are the burn spans; sub-warp groups of G lanes take
//            spans, lanes take pixels: uint8 interleaved bands -> shared histogram atomics
//   hburn    horizontal edges lying exactly on a scanline (burnt separately by GDAL, only when
//            running towards -x) are kept in a side list and folded in without double counting
// Edge lists are staged in shared memory with TMA bulk copies (cp.async.bulk + mbarrier),
// in chunks of ECAP vertices when a road is longer than that.
//
// This translation unit is compiled with -fmad=false and uses explicit _rn intrinsics for the
// geometry: the rounding of every operation is part of the specification.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rs_internal.h"

namespace rs {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA 1-D bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of dst
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------
struct ZonalArgs {
    const double2 *xy;
    const int *ring_off;
    const int *road_ring_off;
    const double *road_bbox;
    const int *road_pair_off;
    const int *pair_tile;
    const int *road_list;     // optional: indices of the roads this launch handles
    const int *n_list_dev;    // optional: device-resident length of road_list
    int n_roads;              // number of work items when n_list_dev == nullptr
    const void *pixels;
    const double *gt;
    int H, W;
    const int *road_slot;
    uint32_t *hist;
    uint32_t *nzero;
    uint8_t *masks;
    int window_mode;
    double sk[4], so[4];
    int *work_counter;
    int *status;
};

// per-pair geometry: integer window inside the tile + world->window-pixel transform
struct PairGeom {
    int col_off, row_off, w, h;
    double inv0, inv1, inv3, inv5;
};

// rasterio geometry_window + window_transform + GDALInvGeoTransform, from the road bbox
// (px/py are monotone in x/y for north-up transforms, so the vertex-wise bounds rasterio takes
// are attained at the bbox corners).  Returns 0 (shapes do not overlap raster), 1, or <0.
__device__ __forceinline__ int pair_geometry(const double *__restrict__ gt, const double *__restrict__ bb, int W, int H,
                                             int window_mode, PairGeom &g)
{
    const double sa = gt[0], sb = gt[1], sc = gt[2], sd = gt[3], se = gt[4], sf = gt[5];
    if (sb != 0.0 || sd != 0.0 || sa == 0.0 || se == 0.0) return RS_ERR_ROTATED;
    if (window_mode == RS_WINDOW_FULL) {
        g.col_off = 0; g.row_off = 0; g.w = W; g.h = H;
        g.inv0 = __ddiv_rn(-sc, sa); g.inv1 = __ddiv_rn(1.0, sa);
        g.inv3 = __ddiv_rn(-sf, se); g.inv5 = __ddiv_rn(1.0, se);
        return 1;
    }
    // Affine.__invert__
    const double det = __dsub_rn(__dmul_rn(sa, se), __dmul_rn(sb, sd));
    const double idet = __ddiv_rn(1.0, det);
    const double ra = __dmul_rn(se, idet), rb = __dmul_rn(-sb, idet);
    const double rd = __dmul_rn(-sd, idet), re = __dmul_rn(sa, idet);
    const double rc = __dsub_rn(__dmul_rn(-sc, ra), __dmul_rn(sf, rb));
    const double rf = __dsub_rn(__dmul_rn(-sc, rd), __dmul_rn(sf, re));
    const double xmin = bb[0], ymin = bb[1], xmax = bb[2], ymax = bb[3];
    // (vx*ra + vy*rb) + rc ; (vx*rd + vy*re) + rf
    const double pxa = __dadd_rn(__dadd_rn(__dmul_rn(xmin, ra), __dmul_rn(ymin, rb)), rc);
    const double pxb = __dadd_rn(__dadd_rn(__dmul_rn(xmax, ra), __dmul_rn(ymin, rb)), rc);
    const double pya = __dadd_rn(__dadd_rn(__dmul_rn(xmin, rd), __dmul_rn(ymin, re)), rf);
    const double pyb = __dadd_rn(__dadd_rn(__dmul_rn(xmin, rd), __dmul_rn(ymax, re)), rf);
    const double left = fmin(pxa, pxb), right = fmax(pxa, pxb);
    const double top = fmin(pya, pyb), bottom = fmax(pya, pyb);
    if (!(left == left) || !(right == right) || !(top == top) || !(bottom == bottom)) return 0;
    const double r0 = floor(top), c0 = floor(left);
    const double hh = fmax(ceil(bottom) - r0, 0.0), ww = fmax(ceil(right) - c0, 0.0);
    const double r1 = r0 + hh, c1 = c0 + ww;
    if (r0 >= (double)H || r1 <= 0.0 || c0 >= (double)W || c1 <= 0.0) return 0;   // rasterio WindowError
    const int ir0 = (int)fmax(r0, 0.0), ir1 = (int)fmin(r1, (double)H);
    const int ic0 = (int)fmax(c0, 0.0), ic1 = (int)fmin(c1, (double)W);
    g.col_off = ic0; g.row_off = ir0; g.w = ic1 - ic0; g.h = ir1 - ir0;
    if (g.w <= 0 || g.h <= 0) return 0;
    // transform * Affine.translation(col_off, row_off), then GDALInvGeoTransform (north-up branch)
    const double xo = (double)ic0, yo = (double)ir0;
    const double wa = __dadd_rn(__dmul_rn(sa, 1.0), __dmul_rn(sb, 0.0));
    const double wc = __dadd_rn(__dadd_rn(__dmul_rn(sa, xo), __dmul_rn(sb, yo)), sc);
    const double we = __dadd_rn(__dmul_rn(sd, 0.0), __dmul_rn(se, 1.0));
    const double wf = __dadd_rn(__dadd_rn(__dmul_rn(sd, xo), __dmul_rn(se, yo)), sf);
    g.inv0 = __ddiv_rn(-wc, wa); g.inv1 = __ddiv_rn(1.0, wa);
    g.inv3 = __ddiv_rn(-wf, we); g.inv5 = __ddiv_rn(1.0, we);
    return 1;
}

// smallest integer y with y + 0.5 >= v / largest integer y with y + 0.5 < v (exact comparisons)
__device__ __forceinline__ int first_row_ge(double v)
{
    double t = fmin(fmax(v - 0.5, -4.0), 1.0e6);
    int y = (int)ceil(t);
    if ((double)(y - 1) + 0.5 >= v) y--;
    if ((double)y + 0.5 < v) y++;
    return y;
}
__device__ __forceinline__ int last_row_lt(double v)
{
    double t = fmin(fmax(v - 0.5, -4.0), 1.0e6);
    int y = (int)ceil(t) - 1;
    if ((double)(y + 1) + 0.5 < v) y++;
    if ((double)y + 0.5 >= v) y--;
    return y;
}

// ---------------------------------------------------------------------------------------------
// pixel policies
// ---------------------------------------------------------------------------------------------
template <int C_>
struct PxBandsU8 {
    static constexpr int C = C_, HC = C_, ELEM = 1;
    static constexpr bool MASK = false;
    __device__ static __forceinline__ void pixel(const ZonalArgs &a, size_t tile_idx, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint8_t *p = (const uint8_t *)a.pixels + (tile_idx * (size_t)a.H * a.W + pix) * C;
        uint32_t v[C];
        if (C == 4) {
            const uint32_t q = __ldg((const uint32_t *)p);
            v[0] = q & 255u; v[1] = (q >> 8) & 255u; v[2 % C] = (q >> 16) & 255u; v[3 % C] = q >> 24;
        } else if (C == 2) {
            const uint32_t q = __ldg((const uint16_t *)p);
            v[0] = q & 255u; v[1 % C] = q >> 8;
        } else {
#pragma unroll
            for (int c = 0; c < C; c++) v[c] = __ldg(p + c);
        }
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            atomicAdd(&hist[c * 256 + v[c]], 1u);
            any |= v[c];
        }
        nz += (any == 0);
    }
};

struct PxClassScore {
    static constexpr int C = 2, HC = 3, ELEM = 1;
    static constexpr bool MASK = false;
    __device__ static __forceinline__ void pixel(const ZonalArgs &a, size_t tile_idx, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint8_t *p = (const uint8_t *)a.pixels + (tile_idx * (size_t)a.H * a.W + pix) * 2;
        const uint32_t q = __ldg((const uint16_t *)p);
        uint32_t cls = q & 255u;
        const uint32_t score = q >> 8;
        if (cls > 2u) cls = 0u;   // unknown class codes count as "no detection"
        atomicAdd(&hist[cls * 256 + score], 1u);
        nz += (q == 0);
    }
};

template <bool F32>
struct PxU16x4Rescale {
    static constexpr int C = 4, HC = 4, ELEM = 2;
    static constexpr bool MASK = false;
    __device__ static __forceinline__ void pixel(const ZonalArgs &a, size_t tile_idx, size_t pix, uint32_t *hist, uint32_t &nz)
    {
        const uint16_t *p = (const uint16_t *)a.pixels + (tile_idx * (size_t)a.H * a.W + pix) * 4;
        const uint2 q = __ldg((const uint2 *)p);
        const uint32_t s[4] = {q.x & 0xffffu, q.x >> 16, q.y & 0xffffu, q.y >> 16};
        uint32_t any = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t o;
            if (F32) {
                float f = __fadd_rn(__fmul_rn((float)s[c], (float)a.sk[c]), (float)a.so[c]);
                f = fminf(fmaxf(f, 0.0f), 255.0f);
                o = (uint32_t)(int)__fadd_rn(f, 0.5f);
            } else {
                double f = __dadd_rn(__dmul_rn((double)s[c], a.sk[c]), a.so[c]);
                f = fmin(fmax(f, 0.0), 255.0);
                o = (uint32_t)(int)__dadd_rn(f, 0.5);
            }
            atomicAdd(&hist[c * 256 + o], 1u);
            any |= o;
        }
        nz += (any == 0);
    }
};

struct PxMask {
    static constexpr int C = 1, HC = 0, ELEM = 1;
    static constexpr bool MASK = true;
};

// ---------------------------------------------------------------------------------------------
// team-wide helpers (one team per CTA: blockDim.x == TEAM)
// ---------------------------------------------------------------------------------------------
template <int TEAM>
__device__ __forceinline__ void team_sync()
{
    if (TEAM == 32) __syncwarp();
    else __syncthreads();
}

// In-place exclusive prefix sum of a[0..n) in shared memory; returns the total to every thread.
// If `inclusive`, a[i] becomes the inclusive sum instead.
template <int TEAM>
__device__ int team_scan(int *a, int n, bool inclusive, int *wsum)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int ipt = (n + TEAM - 1) / TEAM;
    const int b = min(tid * ipt, n), e = min(b + ipt, n);
    int s = 0;
    for (int i = b; i < e; i++) s += a[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int base = incl - s, total;
    if (TEAM > 32) {
        const int warp = tid >> 5;
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        total = 0;
#pragma unroll
        for (int w = 0; w < TEAM / 32; w++) {
            const int v = wsum[w];
            if (w < warp) wbase += v;
            total += v;
        }
        base += wbase;
        __syncthreads();
    } else {
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
    int run = base;
    for (int i = b; i < e; i++) {
        const int v = a[i];
        run += v;
        a[i] = inclusive ? run : run - v;
    }
    team_sync<TEAM>();
    return total;
}

// smallest i in [0, n) with a[i] > q (a non-decreasing); n if none
__device__ __forceinline__ int upper_bound_smem(const int *a, int n, int q)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] > q) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// shared-memory layout
// ---------------------------------------------------------------------------------------------
template <int TEAM>
struct Cfg;
template <>
struct Cfg<32> {
    static constexpr int ECAP = 128;    // vertices staged per chunk
    static constexpr int RC = 256;      // scanlines per chunk
    static constexpr int CAP = 1024;    // crossings per scanline chunk
    static constexpr int HB = 16;       // horizontal-edge burns per scanline chunk
    static constexpr int G = 8;         // lanes per span
};
template <>
struct Cfg<256> {
    static constexpr int ECAP = 2048;
    static constexpr int RC = 1024;
    static constexpr int CAP = 16384;
    static constexpr int HB = 128;
    static constexpr int G = 32;
};

template <int TEAM, int HC>
struct Smem {
    using K = Cfg<TEAM>;
    alignas(16) double2 verts[K::ECAP];
    alignas(16) uint32_t hist[HC > 0 ? HC * 256 : 4];
    int eoff[K::ECAP + 1];
    int rowpos[K::RC + 2];
    int hb[K::HB][4];               // y (chunk-relative), x first, x last, unused
    int16_t pool[K::CAP];
    alignas(8) uint64_t mbar;
    int wsum[TEAM / 32 + 1];
    int work;
    int hb_count;
};

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int TEAM, class PX>
__global__ void __launch_bounds__(TEAM) zonal_kernel(const ZonalArgs a)
{
    using K = Cfg<TEAM>;
    using S = Smem<TEAM, PX::HC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x;

    if (tid == 0) mbar_init(&s.mbar, 1);
    team_sync<TEAM>();
    uint32_t mbar_phase = 0;

    const int n_work = a.n_list_dev ? *a.n_list_dev : a.n_roads;

    for (;;) {
        if (tid == 0) s.work = atomicAdd(a.work_counter, 1);
        team_sync<TEAM>();
        const int wi = s.work;
        team_sync<TEAM>();
        if (wi >= n_work) break;
        const int road = a.road_list ? a.road_list[wi] : wi;

        if (!PX::MASK) {
            for (int i = tid; i < PX::HC * 64; i += TEAM) reinterpret_cast<uint4 *>(s.hist)[i] = make_uint4(0, 0, 0, 0);
        }
        uint32_t nz = 0;

        const int g0 = a.road_ring_off[road], g1 = a.road_ring_off[road + 1];
        const int v0 = a.ring_off[g0], v1 = a.ring_off[g1];
        const int nv = v1 - v0;
        const int nchunks = (nv + K::ECAP - 1) / K::ECAP;
        const bool one_ring = (g1 - g0) == 1;
        int cbase = 0, ccount = 0;      // vertex chunk currently staged in shared memory
        int staged = -1;

        auto stage_chunk = [&](int ch) {
            if (staged == ch) return;
            team_sync<TEAM>();          // everybody is done reading the previous chunk
            cbase = ch * K::ECAP;
            ccount = min(K::ECAP, nv - cbase);
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)ccount * 16u;
                mbar_arrive_expect_tx(&s.mbar, bytes);
                tma_bulk_g2s(s.verts, a.xy + v0 + cbase, bytes, &s.mbar);
            }
            mbar_wait(&s.mbar, mbar_phase);
            mbar_phase ^= 1u;
            staged = ch;
        };
        auto vertex = [&](int i) -> double2 {
            if (i >= cbase && i < cbase + ccount) return s.verts[i - cbase];
            return a.xy[v0 + i];
        };
        // previous vertex of i along its ring (GDAL: the first index of a ring pairs with its last)
        auto prev_index = [&](int i) -> int {
            if (one_ring) return i == 0 ? nv - 1 : i - 1;
            int lo = g0, hi = g1;       // ring containing vertex v0 + i
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (a.ring_off[mid] - v0 <= i) lo = mid;
                else hi = mid;
            }
            const int rs_ = a.ring_off[lo] - v0, re_ = a.ring_off[lo + 1] - v0;
            return i == rs_ ? re_ - 1 : i - 1;
        };

        const double *bb = a.road_bbox + 4 * (size_t)road;
        const int p_begin = a.road_pair_off[road], p_end = a.road_pair_off[road + 1];

        if (nv > 0 && p_end > p_begin) stage_chunk(0);

        for (int p = p_begin; p < p_end && nv > 0; p++) {
            const int t = a.pair_tile[p];
            PairGeom g;
            const int gr = pair_geometry(a.gt + 6 * (size_t)t, bb, a.W, a.H, a.window_mode, g);
            if (gr < 0) {
                if (tid == 0) atomicMin(a.status, gr);
                continue;
            }
            if (gr == 0) continue;
            const int maxx = g.w - 1;

            // pixel-space edge (ind1 = previous vertex, ind2 = vertex i) of the staged chunk
            auto edge = [&](int i, double &x1, double &y1, double &x2, double &y2) {
                const double2 q2 = vertex(i), q1 = vertex(prev_index(i));
                x1 = __dadd_rn(g.inv0, __dmul_rn(q1.x, g.inv1));
                y1 = __dadd_rn(g.inv3, __dmul_rn(q1.y, g.inv5));
                x2 = __dadd_rn(g.inv0, __dmul_rn(q2.x, g.inv1));
                y2 = __dadd_rn(g.inv3, __dmul_rn(q2.y, g.inv5));
            };

            for (int r0 = 0; r0 < g.h;) {
                int rc = min(K::RC, g.h - r0);
                int total = 0;
                // ---------------- pass 1: per-row and per-edge crossing counts ----------------
                for (;;) {
                    for (int i = tid; i <= rc; i += TEAM) s.rowpos[i] = 0;
                    if (tid == 0) s.hb_count = 0;
                    team_sync<TEAM>();
                    for (int ch = 0; ch < nchunks; ch++) {
                        stage_chunk(ch);
                        for (int j = tid; j < ccount; j += TEAM) {
                            double x1, y1, x2, y2;
                            edge(cbase + j, x1, y1, x2, y2);
                            int n = 0;
                            if (y1 == y2) {
                                // horizontal edge: burnt separately iff it lies exactly on a scanline
                                // of this chunk and runs towards -x
                                if (x1 > x2) {
                                    const double fy = floor(y1);
                                    if (fy + 0.5 == y1 && fy >= (double)r0 && fy < (double)(r0 + rc)) {
                                        const double hx1 = floor(__dadd_rn(x2, 0.5)), hx2 = floor(__dadd_rn(x1, 0.5));
                                        if (!(hx1 > (double)maxx || hx2 <= 0.0)) {
                                            const int k = atomicAdd(&s.hb_count, 1);
                                            if (k < K::HB) {
                                                s.hb[k][0] = (int)fy - r0;
                                                s.hb[k][1] = (int)fmax(hx1, 0.0);
                                                s.hb[k][2] = (int)fmin(hx2 - 1.0, (double)maxx);
                                            }
                                        }
                                    }
                                }
                            } else {
                                const int ya = max(first_row_ge(fmin(y1, y2)), r0);
                                const int yb = min(last_row_lt(fmax(y1, y2)), r0 + rc - 1);
                                n = max(yb - ya + 1, 0);
                                if (n > 0) {
                                    atomicAdd(&s.rowpos[ya - r0], 1);
                                    atomicAdd(&s.rowpos[yb + 1 - r0], -1);
                                }
                            }
                            if (nchunks == 1) s.eoff[j] = n;
                        }
                    }
                    team_sync<TEAM>();
                    team_scan<TEAM>(s.rowpos, rc + 1, true, s.wsum);             // difference -> counts
                    total = team_scan<TEAM>(s.rowpos, rc + 1, false, s.wsum);    // counts -> row offsets
                    if (total <= K::CAP && s.hb_count <= K::HB) break;
                    if (rc == 1) {
                        if (tid == 0) atomicMin(a.status, (int)RS_ERR_CAPACITY);
                        total = -1;
                        break;
                    }
                    rc = (rc + 1) >> 1;
                    team_sync<TEAM>();
                }
                if (total < 0) { r0 += rc; team_sync<TEAM>(); continue; }

                // ---------------- pass 2: evaluate and scatter the crossings ----------------
                for (int ch = 0; ch < nchunks && total > 0; ch++) {
                    stage_chunk(ch);
                    if (nchunks > 1) {
                        for (int j = tid; j < ccount; j += TEAM) {
                            double x1, y1, x2, y2;
                            edge(cbase + j, x1, y1, x2, y2);
                            int n = 0;
                            if (y1 != y2) {
                                const int ya = max(first_row_ge(fmin(y1, y2)), r0);
                                const int yb = min(last_row_lt(fmax(y1, y2)), r0 + rc - 1);
                                n = max(yb - ya + 1, 0);
                            }
                            s.eoff[j] = n;
                        }
                    }
                    if (tid == 0) s.eoff[ccount] = 0;
                    team_sync<TEAM>();
                    const int nx = team_scan<TEAM>(s.eoff, ccount + 1, false, s.wsum);
                    for (int f = tid; f < nx; f += TEAM) {
                        const int j = upper_bound_smem(s.eoff, ccount + 1, f) - 1;
                        const int k = f - s.eoff[j];
                        double x1, y1, x2, y2;
                        edge(cbase + j, x1, y1, x2, y2);
                        double dx1, dy1, dx2, dy2;
                        if (y1 < y2) { dx1 = x1; dy1 = y1; dx2 = x2; dy2 = y2; }
                        else         { dx1 = x2; dy1 = y2; dx2 = x1; dy2 = y1; }
                        const int y = max(first_row_ge(dy1), r0) + k;
                        const double dy = (double)y + 0.5;
                        const double isect =
                            __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(dy, dy1), __dsub_rn(dx2, dx1)), __dsub_rn(dy2, dy1)), dx1);
                        double r = floor(__dadd_rn(isect, 0.5));
                        r = fmin(fmax(r, -1.0), (double)g.w);      // order-preserving clamp, see DESIGN.md
                        const int slot = atomicAdd(&s.rowpos[y - r0], 1);
                        s.pool[slot] = (int16_t)(int)r;
                    }
                    team_sync<TEAM>();
                }

                // ---------------- sort each row's crossings (rowpos[y] is now the row's end) ----------
                for (int y = tid; y < rc; y += TEAM) {
                    const int b = y ? s.rowpos[y - 1] : 0, e = s.rowpos[y];
                    for (int i = b + 1; i < e; i++) {
                        const int16_t v = s.pool[i];
                        int j = i - 1;
                        while (j >= b && s.pool[j] > v) { s.pool[j + 1] = s.pool[j]; j--; }
                        s.pool[j + 1] = v;
                    }
                }
                team_sync<TEAM>();

                // ---------------- spans -> pixels ----------------
                const size_t mask_base = PX::MASK ? (size_t)p * a.H * a.W : 0;
                {
                    constexpr int G = K::G, NG = TEAM / G;
                    const int gid = tid / G, gl = tid % G;
                    const int nspans = total >> 1;
                    for (int m = gid; m < nspans; m += NG) {
                        const int q = 2 * m;
                        const int y = upper_bound_smem(s.rowpos, rc, q);
                        int xs = s.pool[q], xe = s.pool[q + 1];
                        if (!(xs <= maxx && xe > 0)) continue;
                        xs = max(xs, 0);
                        xe = min(xe - 1, maxx);
                        const size_t rowbase = (size_t)(g.row_off + r0 + y) * a.W + g.col_off;
                        for (int x = xs + gl; x <= xe; x += G) {
                            if constexpr (PX::MASK) a.masks[mask_base + rowbase + x] = 1;
                            else PX::pixel(a, (size_t)t, rowbase + x, s.hist, nz);
                        }
                    }
                }
                // ---------------- horizontal-edge burns, minus what is already covered ----------------
                const int nhb = min(s.hb_count, (int)K::HB);
                for (int k = 0; k < nhb; k++) {
                    const int y = s.hb[k][0], xa = s.hb[k][1], xb = s.hb[k][2];
                    const int b = y ? s.rowpos[y - 1] : 0, e = s.rowpos[y];
                    const size_t rowbase = (size_t)(g.row_off + r0 + y) * a.W + g.col_off;
                    for (int x = xa + tid; x <= xb; x += TEAM) {
                        bool covered = false;
                        for (int i = b; i + 1 < e; i += 2) covered |= (s.pool[i] <= x && x < s.pool[i + 1]);
                        for (int k2 = 0; k2 < k; k2++) covered |= (s.hb[k2][0] == y && s.hb[k2][1] <= x && x <= s.hb[k2][2]);
                        if (covered) continue;
                        if constexpr (PX::MASK) a.masks[mask_base + rowbase + x] = 1;
                        else PX::pixel(a, (size_t)t, rowbase + x, s.hist, nz);
                    }
                }
                team_sync<TEAM>();
                r0 += rc;
            }
        }

        // ---------------- write the road's accumulators (exactly once) ----------------
        if constexpr (!PX::MASK) {
            team_sync<TEAM>();
            const int slot = a.road_slot ? a.road_slot[road] : road;
            uint4 *dst = reinterpret_cast<uint4 *>(a.hist + (size_t)slot * PX::HC * 256);
            for (int i = tid; i < PX::HC * 64; i += TEAM) dst[i] = reinterpret_cast<const uint4 *>(s.hist)[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
            if (TEAM > 32) {
                if ((tid & 31) == 0) s.wsum[tid >> 5] = (int)nz;
                __syncthreads();
                nz = 0;
                for (int w = 0; w < TEAM / 32; w++) nz += (uint32_t)s.wsum[w];
            }
            if (tid == 0) a.nzero[slot] = nz;
        }
        team_sync<TEAM>();
    }
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
template <int TEAM, class PX>
static int launch_one(rs_ctx *ctx, const ZonalArgs &args, int n_items, cudaStream_t st)
{
    using S = Smem<TEAM, PX::HC>;
    const size_t smem = sizeof(S);
    auto kern = zonal_kernel<TEAM, PX>;
    RS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    RS_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TEAM, smem));
    if (per_sm < 1) per_sm = 1;
    long grid = (long)ctx->sm_count * per_sm;     // persistent: a whole number of resident waves
    if (n_items >= 0 && grid > n_items) grid = n_items > 0 ? n_items : 1;
    RS_CUDA_OK(ctx, cudaMemsetAsync(args.work_counter, 0, sizeof(int), st));
    kern<<<(unsigned)grid, TEAM, smem, st>>>(args);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

int launch_zonal(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                 const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks, int window_mode,
                 cudaStream_t st)
{
    if (!roads || !tiles || !pairs) return RS_ERR_INVALID_ARG;
    if (roads->n_roads < 0 || roads->n_verts < 0 || tiles->n_tiles < 0 || pairs->n_pairs < 0) return RS_ERR_INVALID_ARG;
    if (roads->n_roads == 0) return RS_OK;
    if (!roads->xy || !roads->ring_off || !roads->road_ring_off || !roads->road_bbox || !pairs->road_pair_off)
        return RS_ERR_INVALID_ARG;
    if (pairs->n_pairs > 0 && (!pairs->pair_tile || !tiles->gt)) return RS_ERR_INVALID_ARG;
    if (((uintptr_t)roads->xy & 15u) != 0) return RS_ERR_INVALID_ARG;      // TMA bulk source alignment
    if (tiles->width > 32766 || tiles->width < 1 || tiles->height < 1) return RS_ERR_UNSUPPORTED;
    if (window_mode != RS_WINDOW_CROP && window_mode != RS_WINDOW_FULL) return RS_ERR_INVALID_ARG;

    ZonalArgs a{};
    a.xy = (const double2 *)roads->xy;
    a.ring_off = roads->ring_off;
    a.road_ring_off = roads->road_ring_off;
    a.road_bbox = roads->road_bbox;
    a.road_pair_off = pairs->road_pair_off;
    a.pair_tile = pairs->pair_tile;
    a.road_list = nullptr;
    a.n_list_dev = nullptr;
    a.n_roads = roads->n_roads;
    a.pixels = tiles->pixels;
    a.gt = tiles->gt;
    a.H = tiles->height;
    a.W = tiles->width;
    a.road_slot = prm ? prm->road_slot : nullptr;
    a.hist = hist;
    a.nzero = n_allzero;
    a.masks = masks;
    a.window_mode = window_mode;
    a.work_counter = ctx->d_counters;
    a.status = ctx->d_status;
    if (prm)
        for (int c = 0; c < 4; c++) { a.sk[c] = prm->scale_k[c]; a.so[c] = prm->scale_off[c]; }

    if (masks) return launch_one<32, PxMask>(ctx, a, roads->n_roads, st);

    if (!prm || !hist || !n_allzero || (pairs->n_pairs > 0 && !tiles->pixels)) return RS_ERR_INVALID_ARG;
    const int C = tiles->channels;
    if (prm->hist_mode == RS_HIST_CLASS_SCORE) {
        if (C != 2 || tiles->dtype != RS_U8) return RS_ERR_UNSUPPORTED;
        return launch_one<32, PxClassScore>(ctx, a, roads->n_roads, st);
    }
    if (prm->hist_mode != RS_HIST_BANDS) return RS_ERR_INVALID_ARG;
    if (tiles->dtype == RS_U16) {
        if (C != 4) return RS_ERR_UNSUPPORTED;
        if (((uintptr_t)tiles->pixels & 7u) != 0) return RS_ERR_INVALID_ARG;
        if (prm->rescale == 1) return launch_one<32, PxU16x4Rescale<false>>(ctx, a, roads->n_roads, st);
        if (prm->rescale == 2) return launch_one<32, PxU16x4Rescale<true>>(ctx, a, roads->n_roads, st);
        return RS_ERR_INVALID_ARG;
    }
    if (tiles->dtype != RS_U8) return RS_ERR_INVALID_ARG;
    switch (C) {
        case 1: return launch_one<32, PxBandsU8<1>>(ctx, a, roads->n_roads, st);
        case 2:
            if (((uintptr_t)tiles->pixels & 1u) != 0) return RS_ERR_INVALID_ARG;
            return launch_one<32, PxBandsU8<2>>(ctx, a, roads->n_roads, st);
        case 3: return launch_one<32, PxBandsU8<3>>(ctx, a, roads->n_roads, st);
        case 4:
            if (((uintptr_t)tiles->pixels & 3u) != 0) return RS_ERR_INVALID_ARG;
            return launch_one<32, PxBandsU8<4>>(ctx, a, roads->n_roads, st);
        default: return RS_ERR_UNSUPPORTED;
    }
}

}  // namespace rs
