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


// ---------------------------------------------------------------------------------------------
// 'within' join: scripts/road_segmentation/determine_class.py:41-62 get_roads_in_quarries
//   gpd.sjoin(roads, buffered_quarries, predicate='within')
// One warp per (road a, polygon b).  a is within b iff (i) no vertex of a lies outside b (even-odd over all rings of b, a
// vertex on b's boundary counts as inside: 'within' allows touching), (ii) no edge of a properly crosses an edge of b,
// (iii) no vertex of b lies strictly inside a.  Plain binary64 orientation signs: results can differ from GEOS'
// robust predicates only for vertices within rounding distance of b's boundary.
// ---------------------------------------------------------------------------------------------
namespace {

struct PolySet {
    const double2 *xy;
    const int *ring_off;
    const int *road_ring_off;
    const double *bbox;
    int n;
};

__device__ __forceinline__ double orient(double2 p, double2 q, double2 r)
{
    return __dsub_rn(__dmul_rn(__dsub_rn(q.x, p.x), __dsub_rn(r.y, p.y)), __dmul_rn(__dsub_rn(q.y, p.y), __dsub_rn(r.x, p.x)));
}

// 0 outside, 1 inside, 2 on the boundary (even-odd over rings [g0, g1); ring i's edge k joins vertex k and its successor,
// the last vertex pairs with the first, so closed and unclosed rings both work)
__device__ int point_in_rings(const PolySet &s, int g0, int g1, double2 p)
{
    int inside = 0;
    for (int g = g0; g < g1; g++) {
        const int v0 = s.ring_off[g], v1 = s.ring_off[g + 1];
        for (int k = v0; k < v1; k++) {
            const double2 a = s.xy[k], b = s.xy[k + 1 < v1 ? k + 1 : v0];
            if (a.x == b.x && a.y == b.y) continue;
            const double o = orient(a, b, p);
            if (o == 0.0 && fmin(a.x, b.x) <= p.x && p.x <= fmax(a.x, b.x) && fmin(a.y, b.y) <= p.y && p.y <= fmax(a.y, b.y)) return 2;
            if ((a.y <= p.y) != (b.y <= p.y)) {            // the edge spans the horizontal line through p (half-open)
                const bool up = b.y > a.y;
                if ((o > 0.0) == up) inside ^= 1;           // p is left of an upward edge / right of a downward edge
            }
        }
    }
    return inside;
}

__global__ void __launch_bounds__(128) within_kernel(const PolySet A, const PolySet B, uint8_t *__restrict__ out)
{
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)A.n * B.n) return;
    const int ia = (int)(wid / B.n), ib = (int)(wid - (long long)ia * B.n);
    const double *ba = A.bbox + 4 * (size_t)ia, *bb = B.bbox + 4 * (size_t)ib;
    bool ok = ba[0] >= bb[0] && ba[1] >= bb[1] && ba[2] <= bb[2] && ba[3] <= bb[3];      // false for NaN / empty boxes
    const int ga0 = A.road_ring_off[ia], ga1 = A.road_ring_off[ia + 1];
    const int gb0 = B.road_ring_off[ib], gb1 = B.road_ring_off[ib + 1];
    const int va0 = A.ring_off[ga0], va1 = A.ring_off[ga1];
    const int vb0 = B.ring_off[gb0], vb1 = B.ring_off[gb1];
    if (va1 <= va0 || vb1 <= vb0) ok = false;
    if (ok) {
        // (i) vertices of a
        for (int k = va0 + lane; k < va1 && ok; k += 32)
            if (point_in_rings(B, gb0, gb1, A.xy[k]) == 0) ok = false;
        ok = __all_sync(0xffffffffu, ok);
    }
    if (ok) {
        // (ii) proper crossings: lanes take edges of a, every lane walks the edges of b
        for (int g = ga0; g < ga1 && ok; g++) {
            const int r0 = A.ring_off[g], r1 = A.ring_off[g + 1];
            for (int k = r0 + lane; k < r1 && ok; k += 32) {
                const double2 p = A.xy[k], q = A.xy[k + 1 < r1 ? k + 1 : r0];
                for (int h = gb0; h < gb1 && ok; h++) {
                    const int s0 = B.ring_off[h], s1 = B.ring_off[h + 1];
                    for (int j = s0; j < s1; j++) {
                        const double2 c = B.xy[j], d = B.xy[j + 1 < s1 ? j + 1 : s0];
                        const double o1 = orient(p, q, c), o2 = orient(p, q, d), o3 = orient(c, d, p), o4 = orient(c, d, q);
                        if (((o1 > 0.0 && o2 < 0.0) || (o1 < 0.0 && o2 > 0.0)) && ((o3 > 0.0 && o4 < 0.0) || (o3 < 0.0 && o4 > 0.0))) {
                            ok = false;
                            break;
                        }
                    }
                }
            }
        }
        ok = __all_sync(0xffffffffu, ok);
    }
    if (ok) {
        // (iii) with (i) and (ii) holding, a vertex of b strictly inside a means a piece of b's boundary (a hole, or
        // another part under the even-odd rule) lies in a's interior, so a has points outside b
        for (int k = vb0 + lane; k < vb1 && ok; k += 32)
            if (point_in_rings(A, ga0, ga1, B.xy[k]) == 1) ok = false;
        ok = __all_sync(0xffffffffu, ok);
    }
    if (lane == 0) out[wid] = ok ? 1 : 0;
}

}  // namespace

int launch_within(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, uint8_t *out, cudaStream_t st)
{
    if (a->n_roads == 0 || b->n_roads == 0) return RS_OK;
    PolySet A{(const double2 *)a->xy, a->ring_off, a->road_ring_off, a->road_bbox, a->n_roads};
    PolySet B{(const double2 *)b->xy, b->ring_off, b->road_ring_off, b->road_bbox, b->n_roads};
    const long long warps = (long long)a->n_roads * b->n_roads;
    const long long blocks = (warps * 32 + 127) / 128;
    if (blocks > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
    within_kernel<<<(unsigned)blocks, 128, 0, st>>>(A, B, out);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
