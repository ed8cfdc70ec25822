// Zonal statistics over a float32 raster (sm_100a): rasterstats.zonal_stats(labels, dem_array, affine=affine,
// stats=['min','max','mean','median','std'], nodata=-9999)  --  scripts/functions/fct_rasters.py:147-163, the reference's one live
// zonal_stats call (a swissALTI3D DEM, float32).  The 256-bin histogram of the uint8 path does not apply to floats: the valid
// in-mask pixels of every feature (not NaN, not nodata) are compacted into the feature's slice of one value array
// (zonal_kernel<PxF32>: a count pass, a scan, a write pass), every slice is sorted (cub::DeviceSegmentedSort), and one warp per
// feature reads count / min / max / median / percentiles off the sorted slice and sums it in binary64 in a fixed order
// (deterministic).  rasterstats itself reduces in float32 (numpy masked arrays keep the raster dtype): its mean / std carry
// ~1e-7 relative rounding, inside the 1e-6 tolerance of the north star; min / max / count / median are exact.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "rs_internal.h"

namespace rs {

namespace {

typedef unsigned long long u64;

struct ToU64 {
    __host__ __device__ u64 operator()(const uint32_t &v) const { return (u64)v; }
};

__global__ void fstats_total_kernel(const uint32_t *__restrict__ cnt, u64 *__restrict__ off, int n)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n] = n ? off[n - 1] + cnt[n - 1] : 0ull;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per feature over its sorted values
__global__ void __launch_bounds__(256) fstats_kernel(const float *__restrict__ sorted, const u64 *__restrict__ off, int n_feat, int ddof,
                                                     int n_pct, const double *__restrict__ pct, double *__restrict__ stats)
{
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (f >= n_feat) return;
    const int NS = RS_NSTAT + n_pct;
    double *out = stats + (size_t)f * NS;
    const u64 b = off[f], n = off[f + 1] - b;
    const float *v = sorted + b;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (n == 0) {
        if (lane == 0) {
            out[RS_STAT_COUNT] = 0.0;
            for (int i = 1; i < NS; i++) out[i] = nan;
        }
        return;
    }
    double s1 = 0.0, s2 = 0.0;
    for (u64 i = lane; i < n; i += 32) {
        const double x = (double)v[i];
        s1 += x;
        s2 += x * x;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const double dn = (double)n, mean = s1 / dn;
    double ss = 0.0;                                     // numpy's std: mean of squared deviations
    for (u64 i = lane; i < n; i += 32) {
        const double d = (double)v[i] - mean;
        ss += d * d;
    }
    ss = warp_sum(ss);
    if (lane == 0) {
        const double sd = n > (u64)ddof ? sqrt(ss / (double)(n - (u64)ddof)) : nan;
        out[RS_STAT_COUNT] = dn;
        out[RS_STAT_MIN] = (double)v[0];
        out[RS_STAT_MAX] = (double)v[n - 1];
        out[RS_STAT_SUM] = s1;
        out[RS_STAT_SUMSQ] = s2;
        out[RS_STAT_MEAN] = mean;
        out[RS_STAT_STD] = sd;
        // np.median of a float32 array: the mean of the two middle elements, taken in float32
        out[RS_STAT_MEDIAN] = (double)__fmul_rn(__fadd_rn(v[(n - 1) >> 1], v[n >> 1]), 0.5f);
        out[RS_STAT_MARGIN] = 2.0 * sd / sqrt(dn);
        for (int i = 0; i < n_pct; i++) {                // numpy.percentile, method 'linear'
            const double vi = (double)(n - 1) * (pct[i] / 100.0);
            double fl = floor(vi);
            if (fl < 0.0) fl = 0.0;
            u64 k0 = (u64)fl;
            if (k0 > n - 1) k0 = n - 1;
            const u64 k1 = k0 + 1 > n - 1 ? n - 1 : k0 + 1;
            const double t = vi - fl, va = (double)v[k0], vb = (double)v[k1], d = vb - va;
            out[RS_NSTAT + i] = t >= 0.5 ? vb - d * (1.0 - t) : va + d * t;
        }
    }
}

}  // namespace

// counts uint32[n] -> offsets u64[n + 1]
int launch_fstats_offsets(rs_ctx *ctx, const uint32_t *counts, int n, unsigned long long *offsets, cudaStream_t st)
{
    if (n <= 0) return RS_OK;
    auto in = thrust::make_transform_iterator(counts, ToU64());
    size_t tmp = 0;
    RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, offsets, n, st));
    int rc = ensure(ctx, ctx->wide_tmp, tmp);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->wide_tmp.p, tmp, in, offsets, n, st));
    fstats_total_kernel<<<1, 32, 0, st>>>(counts, offsets, n);
    ctx->launches += 2;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

int launch_fstats_sort(rs_ctx *ctx, const float *values, float *sorted, long long total, int n, const unsigned long long *offsets,
                       cudaStream_t st)
{
    if (total <= 0 || n <= 0) return RS_OK;
    if (total > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
    size_t tmp = 0;
    RS_CUDA_OK(ctx, cub::DeviceSegmentedSort::SortKeys(nullptr, tmp, values, sorted, (int)total, n, offsets, offsets + 1, st));
    int rc = ensure(ctx, ctx->wide_tmp, tmp);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cub::DeviceSegmentedSort::SortKeys(ctx->wide_tmp.p, tmp, values, sorted, (int)total, n, offsets, offsets + 1, st));
    ctx->launches++;
    return RS_OK;
}

int launch_fstats(rs_ctx *ctx, const float *sorted, const unsigned long long *offsets, int n, int ddof, const double *pct_dev, int n_pct,
                  double *stats, cudaStream_t st)
{
    if (n <= 0) return RS_OK;
    fstats_kernel<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>(sorted, offsets, n, ddof, n_pct, pct_dev, stats);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
