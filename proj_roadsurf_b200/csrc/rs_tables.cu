// K3 statistics from merged histograms, K4 vote + confusion + F1, synthetic tile generator (sm_100a).
//
// K3 replaces pandas groupby.agg(['min','max','median','mean','count','std']) + margin
//    (scripts/functions/fct_statistics.py:55-63, ddof 1) and the rasterstats reductions (ddof 0),
//    including the two nodata conventions of fct_misc.get_pixel_values (fct_misc.py:95-119).
// K4 replaces determine_class.determine_detected_class (scripts/road_segmentation/determine_class.py:122-190),
//    final_metrics.get_tag / get_metrics (scripts/road_segmentation/final_metrics.py:91-105, :22-89)
//    for every cut-off of the sweep (:277-316) in one launch, on raster accumulators.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "rs_internal.h"

namespace rs {

typedef unsigned long long u64;

__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp-cooperative staging for the thread-per-row kernels below: the 128-byte block `blk` of the 32 table rows
// [row0, row0 + 32) (row_u4 uint4 per row) goes to tile[row][0..7] with coalesced 128-byte reads (eight lanes per row); the
// rows are 144 bytes apart in shared memory, so that every lane then reads ITS row with conflict-free 128-bit loads.
typedef uint4 RowTile[32][9];
__device__ __forceinline__ void stage_block(const uint4 *__restrict__ base, long long row0, long long n_rows, int row_u4, int blk,
                                            RowTile &tile, int lane)
{
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int idx = 32 * j + lane, r = idx >> 3, u = idx & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row0 + r < n_rows) v = __ldg(base + (size_t)(row0 + r) * row_u4 + 8 * blk + u);
        tile[r][u] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K3: one THREAD per (road, band).  The row's 256 bins are walked once for count / sum / sum of squares / min / max (every
// bin value is a compile-time constant of the unrolled walk) with the running count kept per block of 32 bins; an order
// statistic is then located in those eight block counts and resolved by re-reading one 128-byte block (an L1 / L2 hit).
// The rows of a warp's 32 threads are 1 KiB apart: the warp stages them block by block through shared memory (stage_block).
// A warp-per-row form with lane-distributed bins (round 1) needed ~8x the warp instructions: two 64-bit warp scans /
// reductions per statistic for 1 KiB of input.
// ---------------------------------------------------------------------------------------------
struct FinalizeArgs {
    const uint32_t *hist;
    const uint32_t *nzero;
    int n_roads, C, mode, ddof, n_pct;
    double pct[16];
    double *stats;
};

// value of the k-th (0-based) smallest element of the multiset described by the row's bins; blk[i] = elements in bins
// [0, 32 (i + 1)), bin0 = the (nodata-adjusted) count of bin 0, k < blk[7]
__device__ __forceinline__ int order_stat(const uint4 *__restrict__ hr, const u64 (&blk)[8], uint32_t bin0, u64 k)
{
    int B = 0;
    u64 run = 0;
#pragma unroll
    for (int i = 0; i < 7; i++)
        if (blk[i] <= k) { B = i + 1; run = blk[i]; }
    const uint4 *p = hr + 8 * B;
    int cnt = 0;                        // bins of the block whose inclusive running count is <= k
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint4 q = __ldg(p + i);
        if (i == 0 && B == 0) q.x = bin0;
        run += q.x; cnt += run <= k;
        run += q.y; cnt += run <= k;
        run += q.z; cnt += run <= k;
        run += q.w; cnt += run <= k;
    }
    return 32 * B + min(cnt, 31);
}

__global__ void __launch_bounds__(128) finalize_kernel(const FinalizeArgs a)
{
    __shared__ RowTile tiles[4];
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int C = a.C, NS = RS_NSTAT + a.n_pct, lane = threadIdx.x & 31;
    const long long n_rows = (long long)a.n_roads * C;
    const bool valid = row < n_rows;
    const int road = (int)((valid ? row : n_rows - 1) / C);
    const uint4 *hr = reinterpret_cast<const uint4 *>(a.hist) + (size_t)(valid ? row : n_rows - 1) * 64;
    const u64 allzero = a.nzero ? (u64)a.nzero[road] : 0ull;
    RowTile &tile = tiles[threadIdx.x >> 5];

    u64 n = 0, s1 = 0, s2 = 0, blk[8];
    uint32_t occ[8], bin0 = 0;          // occ[B]: which bins of block B are non-empty
#pragma unroll
    for (int B = 0; B < 8; B++) {
        uint4 q[8];
        __syncwarp();
        stage_block(reinterpret_cast<const uint4 *>(a.hist), row - lane, n_rows, 64, B, tile, lane);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = tile[lane][i];
        if (B == 0) {
            // NONE: rows with every band 0 are dropped (allzero = their count).  ZERO: per (road, tile) call the zeros of a
            // band are dropped and the band is padded with zeros up to the call's longest band; over the road's calls
            // the band keeps zeros(band) - sum over calls of min over bands of zeros(call, band)  (allzero = that sum,
            // rs_zonal_params::min_zero)
            if (a.mode == RS_NODATA_NONE || a.mode == RS_NODATA_ZERO) q[0].x = q[0].x >= allzero ? (uint32_t)(q[0].x - allzero) : 0u;
            else if (a.mode != RS_NODATA_RAW) q[0].x = 0u;          // RS_NODATA_ZERO_MASKED
            bin0 = q[0].x;
        }
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t c[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u64 v = (u64)(32 * B + 4 * i + j);
                n += c[j];
                s1 += v * c[j];
                s2 += v * v * c[j];
                o |= c[j] ? (1u << (4 * i + j)) : 0u;
            }
        }
        blk[B] = n;
        occ[B] = o;
    }
    if (!valid) return;
    double *out = a.stats + (size_t)row * NS;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (n == 0) {
        out[RS_STAT_COUNT] = 0.0;
        for (int i = 1; i < NS; i++) out[i] = nan;
        return;
    }
    int vmin = 256, vmax = -1;
#pragma unroll
    for (int B = 7; B >= 0; B--)
        if (occ[B]) vmin = 32 * B + __ffs(occ[B]) - 1;
#pragma unroll
    for (int B = 0; B < 8; B++)
        if (occ[B]) vmax = 32 * B + 31 - __clz(occ[B]);
    const int med_lo = order_stat(hr, blk, bin0, (n - 1) >> 1);
    const int med_hi = (n & 1ull) ? med_lo : order_stat(hr, blk, bin0, n >> 1);
    for (int i = 0; i < a.n_pct; i++) {
        // numpy.percentile, method 'linear': virtual index (n-1)*q/100, _lerp between neighbours
        const double vi = __dmul_rn((double)(n - 1), __ddiv_rn(a.pct[i], 100.0));
        double fl = floor(vi);
        if (fl < 0.0) fl = 0.0;
        u64 k0 = (u64)fl;
        if (k0 > n - 1) k0 = n - 1;
        const u64 k1 = k0 + 1 > n - 1 ? n - 1 : k0 + 1;
        const double t = __dsub_rn(vi, fl);
        const double va = (double)order_stat(hr, blk, bin0, k0), vb = (double)order_stat(hr, blk, bin0, k1);
        const double d = __dsub_rn(vb, va);
        out[RS_NSTAT + i] = t >= 0.5 ? __dsub_rn(vb, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(va, __dmul_rn(d, t));
    }
    const double dn = (double)n;
    const double mean = (double)s1 / dn;
    double sd = nan;
    if (n > (u64)a.ddof) {
        const unsigned __int128 num = (unsigned __int128)n * s2 - (unsigned __int128)s1 * s1;
        const double var = (double)num / (dn * (double)(n - (u64)a.ddof));
        sd = sqrt(var);
    }
    out[RS_STAT_COUNT] = dn;
    out[RS_STAT_MIN] = (double)vmin;
    out[RS_STAT_MAX] = (double)vmax;
    out[RS_STAT_SUM] = (double)s1;
    out[RS_STAT_SUMSQ] = (double)s2;
    out[RS_STAT_MEAN] = mean;
    out[RS_STAT_STD] = sd;
    out[RS_STAT_MEDIAN] = ((double)med_lo + (double)med_hi) * 0.5;
    out[RS_STAT_MARGIN] = 2.0 * sd / sqrt(dn);        // Z = 2, fct_statistics.py:58-59
}

int launch_finalize(rs_ctx *ctx, const uint32_t *hist, const uint32_t *n_allzero, int n_roads, int channels,
                    int nodata_mode, int ddof, const double *pct_host, int n_pct, double *stats, cudaStream_t st)
{
    if (n_roads < 0 || channels < 1 || channels > 4 || n_pct < 0 || n_pct > 16 || ddof < 0) return RS_ERR_INVALID_ARG;
    if (nodata_mode < RS_NODATA_RAW || nodata_mode > RS_NODATA_ZERO_MASKED) return RS_ERR_INVALID_ARG;
    if (n_roads == 0) return RS_OK;
    if (!hist || !stats || (n_pct > 0 && !pct_host)) return RS_ERR_INVALID_ARG;
    if ((nodata_mode == RS_NODATA_NONE || nodata_mode == RS_NODATA_ZERO) && !n_allzero) return RS_ERR_INVALID_ARG;
    FinalizeArgs a{};
    a.hist = hist; a.nzero = n_allzero; a.n_roads = n_roads; a.C = channels; a.mode = nodata_mode; a.ddof = ddof;
    a.n_pct = n_pct;
    for (int i = 0; i < n_pct; i++) a.pct[i] = pct_host[i];
    a.stats = stats;
    const int threads = 128;
    const unsigned blocks = (unsigned)(((size_t)n_roads * channels + threads - 1) / threads);
    finalize_kernel<<<blocks, threads, 0, st>>>(a);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// K4: vote (one warp per road, all cut-offs) + confusion counts, then metrics (one thread per cut-off)
// ---------------------------------------------------------------------------------------------
enum { RS_MAX_THR = 32 };

struct VoteArgs {
    const uint32_t *jh;
    const int8_t *gt;
    int n_roads, n_thr, rule;
    double min_area_frac;
    int cut[RS_MAX_THR];        // cut-offs sorted in DESCENDING order ...
    int idx[RS_MAX_THR];        // ... and the position of each in the caller's list
    int8_t *cover;
    double *scores;
    u64 *confusion;     // [n_thr][2][4]
};

// One THREAD per road.  The score bins of the two detected classes are walked once from 255 down, keeping the suffix sums
// (pixels and score-weighted pixels with score >= bin); the cut-offs are visited in descending order as the walk reaches
// them, so a cut-off costs its decision and nothing else.  (Round 1: a warp per road and four 64-bit warp reductions per
// cut-off, ~10x the warp instructions.)
__global__ void __launch_bounds__(128) vote_kernel(const VoteArgs a)
{
    __shared__ unsigned int conf[RS_MAX_THR * 8];
    for (int i = threadIdx.x; i < RS_MAX_THR * 8; i += blockDim.x) conf[i] = 0;
    __syncthreads();
    __shared__ RowTile tiles[4][2];
    const int road = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool valid = road < a.n_roads;
    const uint4 *jh4 = reinterpret_cast<const uint4 *>(a.jh);
    RowTile &t1 = tiles[threadIdx.x >> 5][0], &t2 = tiles[threadIdx.x >> 5][1];
    {
        u64 ninside = 0;
        if (a.min_area_frac > 0.0 && valid) {        // pixels of the road in all three classes
            const uint4 *h0 = jh4 + (size_t)road * 192;
#pragma unroll 4
            for (int i = 0; i < 192; i++) {
                const uint4 q = __ldg(h0 + i);
                ninside += (u64)q.x + q.y + q.z + q.w;
            }
        }
        const int g = (a.gt && valid) ? (int)a.gt[road] : -1;
        u64 n1 = 0, n2 = 0, s1 = 0, s2 = 0;
        int nx = 0;
        auto decide = [&]() {                        // cut-off nx with the current suffix sums
            u64 na = n1, nn = n2, sa = s1, sn = s2;
            if (a.min_area_frac > 0.0) {
                const double den = (double)(ninside ? ninside : 1ull);
                const double fa = rint(__dmul_rn(__ddiv_rn((double)na, den), 100.0)) / 100.0;   // np.round(x, 2)
                const double fn = rint(__dmul_rn(__ddiv_rn((double)nn, den), 100.0)) / 100.0;
                if (fa <= a.min_area_frac) { na = 0; sa = 0; }
                if (fn <= a.min_area_frac) { nn = 0; sn = 0; }
            }
            const double ia = (na > 0 && sa > 0) ? (double)sa / (255.0 * (double)na) : 0.0;
            const double in_ = (nn > 0 && sn > 0) ? (double)sn / (255.0 * (double)nn) : 0.0;
            unsigned __int128 left, right;
            if (a.rule == RS_VOTE_COUNT) { left = na; right = nn; }
            else {
                left = na > 0 ? (unsigned __int128)sa * (nn ? nn : 1ull) : 0;
                right = nn > 0 ? (unsigned __int128)sn * (na ? na : 1ull) : 0;
            }
            int cov = RS_COVER_UNDETERMINED;
            if (left > right) cov = RS_COVER_ARTIFICIAL;
            else if (left < right) cov = RS_COVER_NATURAL;
            if (na + nn == 0) cov = RS_COVER_UNDETECTED;
            const int t = a.idx[nx];
            const size_t o = (size_t)t * a.n_roads + road;
            if (a.cover && valid) a.cover[o] = (int8_t)cov;
            if (a.scores && valid) {
                a.scores[3 * o + 0] = ia;
                a.scores[3 * o + 1] = in_;
                a.scores[3 * o + 2] = fabs(ia - in_);
            }
            if (g == 0 || g == 1) atomicAdd(&conf[t * 8 + g * 4 + cov], 1u);
            nx++;
        };
        while (nx < a.n_thr && a.cut[nx] > 255) decide();           // nothing reaches these cut-offs
#pragma unroll 1
        for (int blk = 7; blk >= 0; blk--) {
            __syncwarp();
            stage_block(jh4, road - lane, a.n_roads, 192, 8 + blk, t1, lane);
            stage_block(jh4, road - lane, a.n_roads, 192, 16 + blk, t2, lane);
            __syncwarp();
#pragma unroll 1
            for (int u = 7; u >= 0; u--) {
                const uint4 q1 = t1[lane][u], q2 = t2[lane][u];
                const uint32_t c1[4] = {q1.x, q1.y, q1.z, q1.w}, c2[4] = {q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                for (int j = 3; j >= 0; j--) {
                    const int sc = 32 * blk + 4 * u + j;
                    n1 += c1[j]; s1 += (u64)sc * c1[j];
                    n2 += c2[j]; s2 += (u64)sc * c2[j];
                    while (nx < a.n_thr && a.cut[nx] >= sc) decide();
                }
            }
        }
        while (nx < a.n_thr) decide();                              // cut-offs below 0: everything counts
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_thr * 8; i += blockDim.x)
        if (conf[i]) atomicAdd(&a.confusion[i], (u64)conf[i]);
}

__global__ void metrics_kernel(const u64 *confusion, int n_thr, double *metrics)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_thr) return;
    const u64 *m = confusion + (size_t)i * 8;      // m[g*4 + cov]
    double P[2], R[2], F[2], cnt[2];
    for (int k = 0; k < 2; k++) {
        const double tp = (double)m[k * 4 + k];
        const double fp = (double)m[(1 - k) * 4 + k];
        const double fn = (double)(m[k * 4 + 2] + m[k * 4 + 3] + m[k * 4 + (1 - k)]);
        if (tp == 0.0) { P[k] = 0.0; R[k] = 0.0; F[k] = 0.0; }
        else {
            P[k] = tp / (tp + fp);
            R[k] = tp / (tp + fn);
            F[k] = 2.0 * P[k] * R[k] / (P[k] + R[k]);
        }
        cnt[k] = (double)(m[k * 4] + m[k * 4 + 1] + m[k * 4 + 2] + m[k * 4 + 3]);
    }
    const double total = cnt[0] + cnt[1];
    const double pw = (P[0] * cnt[0] + P[1] * cnt[1]) / total;
    const double rw = (R[0] * cnt[0] + R[1] * cnt[1]) / total;
    const double f1w = (pw == 0.0 && rw == 0.0) ? 0.0 : 2.0 * pw * rw / (pw + rw);
    const double pb = (P[0] + P[1]) / 2.0, rb = (R[0] + R[1]) / 2.0;      // literal 2, final_metrics.py:78-79
    const double f1b = (pb == 0.0 && rb == 0.0) ? 0.0 : 2.0 * pb * rb / (pb + rb);
    double *o = metrics + (size_t)i * RS_NMETRIC;
    o[RS_MET_P0] = P[0]; o[RS_MET_R0] = R[0]; o[RS_MET_F0] = F[0];
    o[RS_MET_P1] = P[1]; o[RS_MET_R1] = R[1]; o[RS_MET_F1] = F[1];
    o[RS_MET_PW] = pw; o[RS_MET_RW] = rw; o[RS_MET_F1W] = f1w;
    o[RS_MET_PB] = pb; o[RS_MET_RB] = rb; o[RS_MET_F1B] = f1b;
}

int launch_metrics(rs_ctx *ctx, const int64_t *confusion, int n_thr, double *metrics, cudaStream_t st)
{
    if (n_thr < 1 || n_thr > RS_MAX_THR || !confusion || !metrics) return RS_ERR_INVALID_ARG;
    metrics_kernel<<<1, RS_MAX_THR, 0, st>>>((const u64 *)confusion, n_thr, metrics);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

int launch_vote(rs_ctx *ctx, const uint32_t *joint_hist, const int8_t *gt_class, int n_roads,
                const int32_t *cutoffs_host, int n_thr, int rule, double min_area_frac, int8_t *cover, double *scores,
                int64_t *confusion, double *metrics, cudaStream_t st)
{
    if (n_roads < 0 || n_thr < 1 || n_thr > RS_MAX_THR || !cutoffs_host) return RS_ERR_INVALID_ARG;
    if (rule != RS_VOTE_COUNT && rule != RS_VOTE_SCORE) return RS_ERR_INVALID_ARG;
    if (!confusion || (n_roads > 0 && !joint_hist)) return RS_ERR_INVALID_ARG;
    RS_CUDA_OK(ctx, cudaMemsetAsync(confusion, 0, sizeof(int64_t) * 8 * (size_t)n_thr, st));
    if (n_roads > 0) {
        VoteArgs a{};
        a.jh = joint_hist; a.gt = gt_class; a.n_roads = n_roads; a.n_thr = n_thr; a.rule = rule;
        a.min_area_frac = min_area_frac;
        for (int i = 0; i < n_thr; i++) {            // insertion sort, descending, stable
            int k = i;
            while (k > 0 && a.cut[k - 1] < cutoffs_host[i]) { a.cut[k] = a.cut[k - 1]; a.idx[k] = a.idx[k - 1]; k--; }
            a.cut[k] = cutoffs_host[i]; a.idx[k] = i;
        }
        a.cover = cover; a.scores = scores; a.confusion = (u64 *)confusion;
        const int threads = 128;
        const unsigned blocks = (unsigned)(((size_t)n_roads + threads - 1) / threads);
        vote_kernel<<<blocks, threads, 0, st>>>(a);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
    }
    if (metrics) return launch_metrics(ctx, confusion, n_thr, metrics, st);
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// two-sample Kolmogorov-Smirnov D of each road's pixel values against a reference distribution, from
// histograms (scripts/statistical_analysis/statistical_analysis.py:441-451 kstest(road_values, general_values)):
// D = max_v | F_road(v) - F_ref(v) | over the 256 values, exact for integer data.  One warp per road.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ks_kernel(const uint32_t *__restrict__ hist, const int *__restrict__ ref_of_road,
                                                 const unsigned long long *__restrict__ ref_hist, int n_roads, int stride,
                                                 double *__restrict__ d_out, double *__restrict__ n_out)
{
    const int road = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (road >= n_roads) return;
    const uint32_t *h = hist + (size_t)road * stride;
    const int ri = ref_of_road ? ref_of_road[road] : 0;
    const u64 *g = ref_hist + (size_t)(ri < 0 ? 0 : ri) * 256;
    u64 a[8], b[8], sa = 0, sb = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = h[8 * lane + j]; b[j] = g[8 * lane + j]; sa += a[j]; sb += b[j]; }
    u64 ia = sa, ib = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ta; ib += tb; }
    }
    const u64 na = __shfl_sync(0xffffffffu, ia, 31), nb = __shfl_sync(0xffffffffu, ib, 31);
    double d = 0.0;
    if (na > 0 && nb > 0 && ri >= 0) {
        u64 ca = ia - sa, cb = ib - sb;
        const double da = (double)na, db = (double)nb;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            ca += a[j]; cb += b[j];
            d = fmax(d, fabs((double)ca / da - (double)cb / db));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    if (lane == 0) {
        d_out[road] = (na > 0 && nb > 0 && ri >= 0) ? d : __longlong_as_double(0x7ff8000000000000ll);
        if (n_out) n_out[road] = (double)na;
    }
}

int launch_ks(rs_ctx *ctx, const uint32_t *hist, const int *ref_of_road, const unsigned long long *ref_hist, int n_roads, int stride,
              double *d_out, double *n_out, cudaStream_t st)
{
    if (n_roads <= 0) return RS_OK;
    const unsigned blocks = (unsigned)(((size_t)n_roads * 32 + 255) / 256);
    ks_kernel<<<blocks, 256, 0, st>>>(hist, ref_of_road, ref_hist, n_roads, stride, d_out, n_out);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// 16 -> 8 bit rescale as a materialising pass: gdal.Translate(outputType=GDT_Byte, scaleParams=...)
// (scripts/preprocessing/tif2cog.py:260-270), with the band selection of the tile URL
// (config/config_stats.yaml:39, bidx=2&bidx=3&bidx=4&bidx=1).  HBM-bound: 2*C_in bytes read + C_out written per pixel.
// ---------------------------------------------------------------------------------------------
struct RescaleArgs {
    const uint16_t *src;
    uint8_t *dst;
    long long n_px;
    int bidx[4];
    double k[4], off[4];
};

// dst = (int)(clamp(src * k + off, 0, 255) + 0.5), every operation correctly rounded in the working precision.
// The uint16 -> float conversion goes through the 2^23 / 2^52 magic constants (one add on the FP pipe) so that only
// the final float -> int conversion uses the quarter-rate conversion pipe (measured: both conversions there make the
// float32 kernel XU-bound at 75 % of the HBM roofline; replacing the float -> int one as well costs more FP64 adds
// than it saves).
template <bool F32>
__device__ __forceinline__ uint32_t rescale_one(uint32_t v, double k, double off)
{
    if (F32) {
        const float x = __fsub_rn(__int_as_float(0x4b000000 | (int)v), 8388608.0f);        // exact (float)v, v < 2^16
        float f = __fadd_rn(__fmul_rn(x, (float)k), (float)off);
        f = fminf(fmaxf(f, 0.0f), 255.0f);
        return (uint32_t)(int)__fadd_rn(f, 0.5f);
    }
    const double x = __dsub_rn(__hiloint2double(0x43300000, (int)v), 4503599627370496.0); // exact (double)v
    double f = __dadd_rn(__dmul_rn(x, k), off);
    f = fmin(fmax(f, 0.0), 255.0);
    return (uint32_t)(int)__dadd_rn(f, 0.5);
}

// C_IN == C_OUT == 4: a thread converts 4 pixels (2 x 128-bit loads, 1 x 128-bit store)
template <bool F32>
__global__ void __launch_bounds__(256) rescale4_kernel(const RescaleArgs a)
{
    const long long n4 = a.n_px >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(a.src) + 2 * i);
        const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(a.src) + 2 * i + 1);
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t o[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const uint32_t s4[4] = {w[2 * p] & 0xffffu, w[2 * p] >> 16, w[2 * p + 1] & 0xffffu, w[2 * p + 1] >> 16};
            uint32_t r = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int b = a.bidx[c];
                const uint32_t v = b == 0 ? s4[0] : (b == 1 ? s4[1] : (b == 2 ? s4[2] : s4[3]));
                r |= rescale_one<F32>(v, a.k[c], a.off[c]) << (8 * c);
            }
            o[p] = r;
        }
        reinterpret_cast<uint4 *>(a.dst)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // tail pixels
    for (long long px = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; px < a.n_px; px += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < 4; c++) a.dst[px * 4 + c] = (uint8_t)rescale_one<F32>(a.src[px * 4 + a.bidx[c]], a.k[c], a.off[c]);
}

// any 1 <= C_IN, C_OUT <= 4: a thread per pixel
template <bool F32>
__global__ void __launch_bounds__(256) rescale_generic_kernel(const RescaleArgs a, int c_in, int c_out)
{
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < a.n_px; px += (long long)gridDim.x * blockDim.x)
        for (int c = 0; c < c_out; c++) a.dst[px * c_out + c] = (uint8_t)rescale_one<F32>(a.src[px * c_in + a.bidx[c]], a.k[c], a.off[c]);
}

int launch_rescale(rs_ctx *ctx, const uint16_t *src, long long n_px, int c_in, int c_out, const int32_t *bidx_host, const double *k_host,
                   const double *off_host, int f32, uint8_t *dst, cudaStream_t st)
{
    if (n_px < 0 || c_in < 1 || c_in > 4 || c_out < 1 || c_out > 4 || !k_host || !off_host) return RS_ERR_INVALID_ARG;
    if (n_px == 0) return RS_OK;
    if (!src || !dst) return RS_ERR_INVALID_ARG;
    RescaleArgs a{};
    a.src = src; a.dst = dst; a.n_px = n_px;
    for (int c = 0; c < 4; c++) {
        a.bidx[c] = c < c_out ? (bidx_host ? bidx_host[c] : c) : 0;
        if (a.bidx[c] < 0 || a.bidx[c] >= c_in) return RS_ERR_INVALID_ARG;
        a.k[c] = c < c_out ? k_host[c] : 1.0;
        a.off[c] = c < c_out ? off_host[c] : 0.0;
    }
    const int grid = ctx->sm_count * 8;
    const bool vec = c_in == 4 && c_out == 4 && (((uintptr_t)src | (uintptr_t)dst) & 15u) == 0;
    if (vec) {
        if (f32) rescale4_kernel<true><<<grid, 256, 0, st>>>(a);
        else rescale4_kernel<false><<<grid, 256, 0, st>>>(a);
    } else {
        if (f32) rescale_generic_kernel<true><<<grid, 256, 0, st>>>(a, c_in, c_out);
        else rescale_generic_kernel<false><<<grid, 256, 0, st>>>(a, c_in, c_out);
    }
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// tile ingest: decompressed TIFF segments -> the pixel-interleaved batches the overlay kernels read.
// One warp per (tile, row): undoes TIFF predictor 2 (horizontal differencing, per sample, modulo the sample width) with a
// lane-chunk sum + warp scan, swaps the bytes of big-endian 16-bit samples, turns band-sequential planes into interleaved
// pixels, selects / reorders bands (bidx) and optionally applies the 16 -> 8 bit rescale of tif2cog.py:260-270.
// ---------------------------------------------------------------------------------------------
struct AssembleArgs {
    const uint8_t *raw;
    void *out;
    int n_tiles, H, W, c_in, c_out;
    int planar;         // 1: [H][W][c_in] samples, 2: [c_in][H][W]
    int predictor;      // 1 none, 2 horizontal differencing
    int bytes;          // 1 or 2 per sample
    int big_endian;
    int rescale;        // 0 none (output keeps the sample width), 1 float64, 2 float32 (uint8 output)
    int bidx[4];
    double k[4], off[4];
};

__device__ __forceinline__ uint32_t raw_sample(const AssembleArgs &a, const uint8_t *tile, int y, int x, int c)
{
    const size_t idx = a.planar == 1 ? ((size_t)y * a.W + x) * a.c_in + c : ((size_t)c * a.H + y) * a.W + x;
    if (a.bytes == 1) return tile[idx];
    const uint32_t b0 = tile[2 * idx], b1 = tile[2 * idx + 1];
    return a.big_endian ? (b0 << 8) | b1 : (b1 << 8) | b0;
}

__global__ void __launch_bounds__(256) assemble_kernel(const AssembleArgs a)
{
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= (long long)a.n_tiles * a.H) return;
    const int t = (int)(row / a.H), y = (int)(row - (long long)t * a.H);
    const uint8_t *tile = a.raw + (size_t)t * a.H * a.W * a.c_in * a.bytes;
    const int chunk = (a.W + 31) >> 5, x0 = min(lane * chunk, a.W), x1 = min(x0 + chunk, a.W);
    const uint32_t smask = a.bytes == 1 ? 0xffu : 0xffffu;
    uint32_t run[4] = {0, 0, 0, 0};
    if (a.predictor == 2) {
        for (int c = 0; c < a.c_in; c++) {
            uint32_t sum = 0;
            for (int x = x0; x < x1; x++) sum += raw_sample(a, tile, y, x, c);
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            run[c] = incl - sum;                  // sum of the differences left of this lane's chunk
        }
    }
    for (int x = x0; x < x1; x++) {
        uint32_t v[4] = {0, 0, 0, 0};
        for (int c = 0; c < a.c_in; c++) {
            const uint32_t sm = raw_sample(a, tile, y, x, c);
            if (a.predictor == 2) { run[c] += sm; v[c] = run[c] & smask; }
            else v[c] = sm;
        }
        const size_t o = (((size_t)t * a.H + y) * a.W + x) * a.c_out;
        for (int c = 0; c < a.c_out; c++) {
            uint32_t s = v[a.bidx[c]];
            if (a.rescale == 1) s = rescale_one<false>(s, a.k[c], a.off[c]);
            else if (a.rescale == 2) s = rescale_one<true>(s, a.k[c], a.off[c]);
            if (a.bytes == 1 || a.rescale) ((uint8_t *)a.out)[o + c] = (uint8_t)s;
            else ((uint16_t *)a.out)[o + c] = (uint16_t)s;
        }
    }
}

int launch_assemble(rs_ctx *ctx, const uint8_t *raw, int n_tiles, int H, int W, int c_in, int planar, int predictor, int bytes,
                    int big_endian, int c_out, const int32_t *bidx_host, int rescale, const double *k_host, const double *off_host,
                    void *out, cudaStream_t st)
{
    if (n_tiles < 0 || H < 1 || W < 1 || c_in < 1 || c_in > 4 || c_out < 1 || c_out > 4) return RS_ERR_INVALID_ARG;
    if ((planar != 1 && planar != 2) || (predictor != 1 && predictor != 2) || (bytes != 1 && bytes != 2)) return RS_ERR_UNSUPPORTED;
    if (rescale < 0 || rescale > 2 || (rescale && (!k_host || !off_host))) return RS_ERR_INVALID_ARG;
    if (n_tiles == 0) return RS_OK;
    if (!raw || !out) return RS_ERR_INVALID_ARG;
    AssembleArgs a{};
    a.raw = raw; a.out = out; a.n_tiles = n_tiles; a.H = H; a.W = W; a.c_in = c_in; a.c_out = c_out;
    a.planar = planar; a.predictor = predictor; a.bytes = bytes; a.big_endian = big_endian; a.rescale = rescale;
    for (int c = 0; c < 4; c++) {
        a.bidx[c] = c < c_out ? (bidx_host ? bidx_host[c] : c) : 0;
        if (a.bidx[c] < 0 || a.bidx[c] >= c_in) return RS_ERR_INVALID_ARG;
        a.k[c] = (rescale && c < c_out) ? k_host[c] : 1.0;
        a.off[c] = (rescale && c < c_out) ? off_host[c] : 0.0;
    }
    const long long warps = (long long)n_tiles * H;
    const long long blocks = (warps * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
    assemble_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// synthetic tiles: counter-based, any shard regenerates its own tiles from (seed, tile_key)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 mix64(u64 z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void __launch_bounds__(256) synth_kernel(T *__restrict__ px, const int64_t *__restrict__ key, int n_tiles, int H, int W,
                                                    int C, int kind, u64 seed)
{
    const size_t per_tile = (size_t)H * W;
    const size_t total = per_tile * n_tiles;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i / per_tile, p = i - t * per_tile;
        const u64 tk = mix64(seed ^ ((u64)key[t] * 0xD6E8FEB86659FD93ull));
        const u64 hp = mix64(tk ^ (u64)p);
        const bool hole = (hp >> 40) % 100ull == 0ull;          // 1 % of pixels: every band 0
        T *o = px + i * C;
        if (kind == 2) {
            const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
            const u64 cell = mix64(tk ^ (0xC0FFEEull + (u64)(y >> 4) * 65537ull + (u64)(x >> 4)));
            const unsigned cls = (unsigned)(cell & 3ull);
            o[0] = (T)(cls == 3u ? 0u : cls);
            o[1] = (T)((hp >> 8) & 255ull);
            continue;
        }
        for (int c = 0; c < C; c++) {
            const u64 h = mix64(hp + 0x632BE59BD9B4E019ull * (u64)(c + 1));
            unsigned v;
            if (sizeof(T) == 2) v = (unsigned)(h >> 48);
            else if (kind == 1) {
                const int sum = (int)(h & 255) + (int)((h >> 8) & 255) + (int)((h >> 16) & 255) + (int)((h >> 24) & 255);
                int g = 110 + ((sum - 510) * 6) / 148;
                v = (unsigned)(g < 0 ? 0 : (g > 255 ? 255 : g));
            } else v = (unsigned)(h >> 56);
            o[c] = hole ? (T)0 : (T)v;
        }
    }
}

int launch_synth(rs_ctx *ctx, void *pixels, const int64_t *tile_key, int n_tiles, int H, int W, int C, int dtype,
                 int kind, uint64_t seed, cudaStream_t st)
{
    if (n_tiles < 0 || H < 1 || W < 1 || C < 1 || C > 4 || kind < 0 || kind > 2) return RS_ERR_INVALID_ARG;
    if (kind == 2 && (C != 2 || dtype != RS_U8)) return RS_ERR_INVALID_ARG;
    if (n_tiles == 0) return RS_OK;
    if (!pixels || !tile_key) return RS_ERR_INVALID_ARG;
    const int grid = ctx->sm_count * 16;
    if (dtype == RS_U8) synth_kernel<uint8_t><<<grid, 256, 0, st>>>((uint8_t *)pixels, tile_key, n_tiles, H, W, C, kind, seed);
    else if (dtype == RS_U16) synth_kernel<uint16_t><<<grid, 256, 0, st>>>((uint16_t *)pixels, tile_key, n_tiles, H, W, C, kind, seed);
    else return RS_ERR_INVALID_ARG;
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
