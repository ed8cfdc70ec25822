// GPU broad phase: the (road, tile) pair list of
//   scripts/statistical_analysis/statistical_analysis.py:170-171   gpd.sjoin(tiles, roads) + drop_duplicates
// on a regular tile lattice, as road-major CSR.  The predicate is bounding-box overlap (closed), a superset of
// the reference's 'intersects': pairs whose polygon misses the tile contribute no pixel (zonal_kernel culls
// them after one 64-byte record), so downstream results are identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>

#include "rs_internal.h"

namespace rs {

struct LatticeArgs {
    const double *bbox;      // [R][4] xmin, ymin, xmax, ymax
    const double *ext;       // [T][4] tile extents
    const int *lut;          // [ny][nx] tile index or -1 (row iy = increasing y)
    double X0, Y0, tw, th;
    int nx, ny, n_roads;
};

__device__ __forceinline__ bool cell_range(const LatticeArgs &a, int r, int &ix0, int &ix1, int &iy0, int &iy1)
{
    const double *b = a.bbox + 4 * (size_t)r;
    if (!(b[0] <= b[2]) || !(b[1] <= b[3])) return false;            // empty / NaN bbox
    // candidate cells with one cell of slack; the exact overlap test against the tile extents decides
    const double fx0 = floor((b[0] - a.X0) / a.tw) - 1.0, fx1 = floor((b[2] - a.X0) / a.tw) + 1.0;
    const double fy0 = floor((b[1] - a.Y0) / a.th) - 1.0, fy1 = floor((b[3] - a.Y0) / a.th) + 1.0;
    if (fx1 < 0.0 || fy1 < 0.0 || fx0 > (double)(a.nx - 1) || fy0 > (double)(a.ny - 1)) return false;
    ix0 = (int)fmax(fx0, 0.0); ix1 = (int)fmin(fx1, (double)(a.nx - 1));
    iy0 = (int)fmax(fy0, 0.0); iy1 = (int)fmin(fy1, (double)(a.ny - 1));
    return true;
}

__device__ __forceinline__ int hit_tile(const LatticeArgs &a, const double *b, int ix, int iy)
{
    const int t = a.lut[(size_t)iy * a.nx + ix];
    if (t < 0) return -1;
    const double *e = a.ext + 4 * (size_t)t;
    return (b[0] <= e[2] && b[2] >= e[0] && b[1] <= e[3] && b[3] >= e[1]) ? t : -1;
}

__global__ void __launch_bounds__(256) pairs_count_kernel(const LatticeArgs a, int *__restrict__ cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_roads) return;
    int ix0, ix1, iy0, iy1, c = 0;
    if (cell_range(a, r, ix0, ix1, iy0, iy1)) {
        const double *b = a.bbox + 4 * (size_t)r;
        for (int iy = iy0; iy <= iy1; iy++)
            for (int ix = ix0; ix <= ix1; ix++) c += hit_tile(a, b, ix, iy) >= 0;
    }
    cnt[r] = c;
}

__global__ void __launch_bounds__(256) pairs_write_kernel(const LatticeArgs a, const int *__restrict__ off, long long capacity,
                                                          int *__restrict__ pair_tile)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_roads) return;
    int ix0, ix1, iy0, iy1;
    if (!cell_range(a, r, ix0, ix1, iy0, iy1)) return;
    const double *b = a.bbox + 4 * (size_t)r;
    const long long base = off[r];
    int k = 0;
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++) {
            const int t = hit_tile(a, b, ix, iy);
            if (t < 0 || base + k >= capacity) continue;
            // keep the road's tiles sorted by index (the order PairList.from_pairs / the reference loop gives)
            int j = k - 1;
            while (j >= 0 && pair_tile[base + j] > t) { pair_tile[base + j + 1] = pair_tile[base + j]; j--; }
            pair_tile[base + j + 1] = t;
            k++;
        }
}

__global__ void pairs_total_kernel(const int *__restrict__ cnt, int *__restrict__ off, int n_roads)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n_roads] = n_roads ? off[n_roads - 1] + cnt[n_roads - 1] : 0;
}

// road_pair_off_dev int32[R+1] is always written; pair_tile_dev only when phase == 1
int launch_pairs_bbox(rs_ctx *ctx, const double *bbox_dev, int n_roads, const double *ext_dev, const rs_lattice *lat, const int *lut_dev,
                      int *road_pair_off_dev, int *pair_tile_dev, long long capacity, int phase, cudaStream_t st)
{
    LatticeArgs a{};
    a.bbox = bbox_dev; a.ext = ext_dev; a.lut = lut_dev;
    a.X0 = lat->x0; a.Y0 = lat->y0; a.tw = lat->tile_w; a.th = lat->tile_h; a.nx = lat->nx; a.ny = lat->ny; a.n_roads = n_roads;
    const unsigned blocks = (unsigned)((n_roads + 255) / 256);
    int rc;
    if (phase == 0) {
        if ((rc = ensure(ctx, ctx->stage[12], sizeof(int) * ((size_t)n_roads + 1)))) return rc;
        int *cnt = (int *)ctx->stage[12].p;
        if (n_roads) {
            pairs_count_kernel<<<blocks, 256, 0, st>>>(a, cnt);
            ctx->launches++;
            size_t tmp = 0;
            RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, road_pair_off_dev, n_roads, st));
            if ((rc = ensure(ctx, ctx->stage[14], tmp))) return rc;
            RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->stage[14].p, tmp, cnt, road_pair_off_dev, n_roads, st));
        }
        pairs_total_kernel<<<1, 32, 0, st>>>(cnt, road_pair_off_dev, n_roads);
        ctx->launches++;
    } else if (n_roads) {
        pairs_write_kernel<<<blocks, 256, 0, st>>>(a, road_pair_off_dev, capacity, pair_tile_dev);
        ctx->launches++;
    }
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
