// Table-shaped entry points around the overlay kernels (sm_100a):
//   (ordered pixel extraction, fct_misc.get_pixel_values' return value, is zonal_kernel<PxExtract> in rs_zonal.cu)
//   group histograms           the groupby of fct_statistics.get_df_stats_groupby (scripts/functions/fct_statistics.py:55)
//                              on uint8 pixel tables: 256-bin histogram per group, finalized by rs_finalize_stats
//   table vote                 determine_class.determine_detected_class on the detection table
//                              (scripts/road_segmentation/determine_class.py:133-179)
//   band ratios                the per-pixel derived columns of scripts/statistical_analysis/statistical_analysis.py:279-293
//                              (ratios between bands rounded to 3 decimals, VgNIR-BI rounded to 5)
//   calibration bins           bin accuracy of the scores (scripts/road_segmentation/final_metrics.py:541-571)
//   confusion + metrics        final_metrics.get_tag / get_metrics from cover / ground-truth codes
//                              (scripts/road_segmentation/final_metrics.py:22-105)
#include <cuda_runtime.h>
#include <stdint.h>

#include "rs_internal.h"

namespace rs {


// ---------------------------------------------------------------------------------------------
// group histograms of a uint8 column
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_hist_kernel(const uint8_t *__restrict__ values, const int *__restrict__ group, long long n,
                                                         int n_groups, uint32_t *__restrict__ hist)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int g = group[i];
        if (g >= 0 && g < n_groups) atomicAdd(&hist[(size_t)g * 256 + values[i]], 1u);
    }
}

int launch_group_hist(rs_ctx *ctx, const uint8_t *values, const int *group, long long n, int n_groups, uint32_t *hist, cudaStream_t st)
{
    RS_CUDA_OK(ctx, cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 256 * (size_t)n_groups, st));
    if (n > 0) {
        const int blocks = (int)((n + 255) / 256 < (long long)ctx->sm_count * 8 ? (n + 255) / 256 : (long long)ctx->sm_count * 8);
        group_hist_kernel<<<blocks, 256, 0, st>>>(values, group, n, n_groups, hist);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
    }
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// table vote: thread per (road, threshold); rows of a road are contiguous and summed in row order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vote_table_kernel(const int *__restrict__ row_off, const int8_t *__restrict__ cls,
                                                         const double *__restrict__ score, const double *__restrict__ weighted,
                                                         const double *__restrict__ area, int n_roads, const double *__restrict__ thr,
                                                         int n_thr, int8_t *__restrict__ cover, double *__restrict__ scores)
{
    const int road = blockIdx.x * blockDim.x + threadIdx.x, ti = blockIdx.y;
    if (road >= n_roads || ti >= n_thr) return;
    const double th = thr[ti];
    // pandas' groupby(...).sum() (libgroupby.group_sum, pandas >= 1.2) is Kahan-compensated; same update order here, so the
    // exact tie test below sees the sums the reference sees
    double sw[2] = {0.0, 0.0}, sa[2] = {0.0, 0.0}, cw[2] = {0.0, 0.0}, ca[2] = {0.0, 0.0};
    int seen[2] = {0, 0}, any = 0;
    for (int i = row_off[road]; i < row_off[road + 1]; i++) {
        if (!(score[i] >= th)) continue;                       // valid_predictions: score >= threshold
        any = 1;
        const int k = cls[i];
        if (k == 0 || k == 1) {
            const double yw = __dsub_rn(weighted[i], cw[k]), tw = __dadd_rn(sw[k], yw);
            cw[k] = __dsub_rn(__dsub_rn(tw, sw[k]), yw);
            if (cw[k] != cw[k]) cw[k] = 0.0;                   // pandas resets a NaN compensation (inf sums)
            sw[k] = tw;
            const double ya = __dsub_rn(area[i], ca[k]), ta = __dadd_rn(sa[k], ya);
            ca[k] = __dsub_rn(__dsub_rn(ta, sa[k]), ya);
            if (ca[k] != ca[k]) ca[k] = 0.0;
            sa[k] = ta;
            seen[k] = 1;
        }
    }
    // index_k = sum(weighted) / sum(area), 0 when the class is absent or its weighted sum is 0
    const double ia = (seen[0] && sw[0] != 0.0) ? __ddiv_rn(sw[0], sa[0]) : 0.0;
    const double in_ = (seen[1] && sw[1] != 0.0) ? __ddiv_rn(sw[1], sa[1]) : 0.0;
    int cov;
    double diff = 0.0;
    if (!any) cov = RS_COVER_UNDETECTED;
    else if (ia == in_) cov = RS_COVER_UNDETERMINED;
    else {
        cov = ia > in_ ? RS_COVER_ARTIFICIAL : RS_COVER_NATURAL;
        diff = fabs(__dsub_rn(ia, in_));
    }
    const size_t o = (size_t)ti * n_roads + road;
    cover[o] = (int8_t)cov;
    if (scores) {
        scores[3 * o + 0] = any ? ia : 0.0;
        scores[3 * o + 1] = any ? in_ : 0.0;
        scores[3 * o + 2] = diff;
    }
}

int launch_vote_table(rs_ctx *ctx, const int *row_off, const int8_t *cls, const double *score, const double *weighted,
                      const double *area, int n_roads, const double *thr_dev, int n_thr, int8_t *cover, double *scores, cudaStream_t st)
{
    if (n_roads == 0) return RS_OK;
    dim3 grid((n_roads + 127) / 128, n_thr);
    vote_table_kernel<<<grid, 128, 0, st>>>(row_off, cls, score, weighted, area, n_roads, thr_dev, n_thr, cover, scores);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// band ratios: thread per pixel row of the uint8 table, one coalesced float64 column per ratio.
// numpy's round(x, d) is rint(x * 10^d) / 10^d in binary64 (multiply, rint, true_divide), restated with _rn intrinsics;
// 0/0 (NaN) -> 0 and x/0 (inf) -> 1 for the ratios (statistical_analysis.py:286-287); VgNIR-BI keeps its NaN (:289-292)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double np_round(double x, double f) { return __ddiv_rn(rint(__dmul_rn(x, f)), f); }

template <int C>
__global__ void __launch_bounds__(256) band_ratio_kernel(const uint8_t *__restrict__ values, long long n, double *__restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v[C];
        if constexpr (C == 4) {
            const uchar4 q = __ldg(reinterpret_cast<const uchar4 *>(values) + i);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int c = 0; c < C; c++) v[c] = (double)__ldg(values + i * C + c);
        }
        int k = 0;
#pragma unroll
        for (int a = 0; a < C; a++)
#pragma unroll
            for (int b = a + 1; b < C; b++, k++) {
                double r;
                if (v[b] == 0.0) r = v[a] == 0.0 ? 0.0 : 1.0;
                else r = np_round(__ddiv_rn(v[a], v[b]), 1000.0);
                __stcs(out + (size_t)k * n + i, r);
            }
        if constexpr (C == 4) {
            const double num = __dsub_rn(v[1], v[3]), den = __dadd_rn(v[1], v[3]);
            __stcs(out + (size_t)k * n + i, np_round(__ddiv_rn(num, den), 100000.0));      // 0/0 stays NaN
        }
    }
}

int launch_band_ratios(rs_ctx *ctx, const uint8_t *values, long long n, int channels, double *out, cudaStream_t st)
{
    if (n == 0) return RS_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
    switch (channels) {
        case 2: band_ratio_kernel<2><<<(int)blocks, 256, 0, st>>>(values, n, out); break;
        case 3: band_ratio_kernel<3><<<(int)blocks, 256, 0, st>>>(values, n, out); break;
        case 4: band_ratio_kernel<4><<<(int)blocks, 256, 0, st>>>(values, n, out); break;
        default: return RS_ERR_UNSUPPORTED;
    }
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// calibration bins (scripts/road_segmentation/final_metrics.py:541-571): per (group, column, threshold) the number of rows with
// lo[t] < value <= hi[t] that are selected, and how many of those are hits; block-private shared counters, then global atomics
// ---------------------------------------------------------------------------------------------
constexpr int BIN_SHARED = 4096;

__global__ void __launch_bounds__(256) bin_count_kernel(const double *__restrict__ values, const int8_t *__restrict__ sel,
                                                        const int8_t *__restrict__ hit, const int *__restrict__ group, int n, int n_cols,
                                                        int n_groups, const double *__restrict__ lo, const double *__restrict__ hi,
                                                        int n_thr, unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int sc[BIN_SHARED];
    const int total = n_groups * n_cols * n_thr * 2;
    const bool priv = total <= BIN_SHARED;
    if (priv) {
        for (int i = threadIdx.x; i < total; i += blockDim.x) sc[i] = 0;
        __syncthreads();
    }
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int g = group[r];
        if (g < 0 || g >= n_groups) continue;
        for (int k = 0; k < n_cols; k++) {
            if (!sel[(size_t)k * n + r]) continue;
            const double v = values[(size_t)k * n + r];
            const int h = hit[(size_t)k * n + r] != 0;
            for (int t = 0; t < n_thr; t++) {
                if (v > lo[t] && v <= hi[t]) {
                    const int o = ((g * n_cols + k) * n_thr + t) * 2;
                    if (priv) {
                        atomicAdd(&sc[o], 1u);
                        if (h) atomicAdd(&sc[o + 1], 1u);
                    } else {
                        atomicAdd(&counts[o], 1ull);
                        if (h) atomicAdd(&counts[o + 1], 1ull);
                    }
                }
            }
        }
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += blockDim.x)
            if (sc[i]) atomicAdd(&counts[i], (unsigned long long)sc[i]);
    }
}

int launch_bin_counts(rs_ctx *ctx, const double *values, const int8_t *sel, const int8_t *hit, const int *group, int n, int n_cols,
                      int n_groups, const double *lo, const double *hi, int n_thr, int64_t *counts, cudaStream_t st)
{
    RS_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(int64_t) * 2 * (size_t)n_groups * n_cols * n_thr, st));
    if (n == 0) return RS_OK;
    int bx = (n + 255) / 256;
    if (bx > ctx->sm_count * 4) bx = ctx->sm_count * 4;
    bin_count_kernel<<<bx, 256, 0, st>>>(values, sel, hit, group, n, n_cols, n_groups, lo, hi, n_thr, (unsigned long long *)counts);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// confusion counts of cover codes against ground-truth codes (then metrics_kernel of rs_tables.cu)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) confusion_kernel(const int8_t *__restrict__ cover, const int8_t *__restrict__ gt, int n_roads,
                                                        unsigned long long *__restrict__ confusion)
{
    __shared__ unsigned int c[8];
    if (threadIdx.x < 8) c[threadIdx.x] = 0;
    __syncthreads();
    const int ti = blockIdx.y;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_roads; r += gridDim.x * blockDim.x) {
        const int g = gt[r], cv = cover[(size_t)ti * n_roads + r];
        if ((g == 0 || g == 1) && cv >= 0 && cv < 4) atomicAdd(&c[g * 4 + cv], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 8 && c[threadIdx.x]) atomicAdd(&confusion[(size_t)ti * 8 + threadIdx.x], (unsigned long long)c[threadIdx.x]);
}

int launch_confusion(rs_ctx *ctx, const int8_t *cover, const int8_t *gt, int n_roads, int n_thr, int64_t *confusion, double *metrics,
                     cudaStream_t st)
{
    RS_CUDA_OK(ctx, cudaMemsetAsync(confusion, 0, sizeof(int64_t) * 8 * (size_t)n_thr, st));
    if (n_roads > 0) {
        int bx = (n_roads + 255) / 256;
        if (bx > ctx->sm_count * 4) bx = ctx->sm_count * 4;
        confusion_kernel<<<dim3(bx, n_thr), 256, 0, st>>>(cover, gt, n_roads, (unsigned long long *)confusion);
        ctx->launches++;
        RS_CUDA_OK(ctx, cudaGetLastError());
    }
    if (metrics) return launch_metrics(ctx, confusion, n_thr, metrics, st);
    return RS_OK;
}

}  // namespace rs
