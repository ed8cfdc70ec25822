// GPU broad phase: the (road, tile) pair list of
//   scripts/statistical_analysis/statistical_analysis.py:170-171   gpd.sjoin(tiles, roads) + drop_duplicates
// on a regular tile lattice, as road-major CSR.  The predicate is bounding-box overlap (closed), a superset of
// the reference's 'intersects': pairs whose polygon misses the tile contribute no pixel (zonal_kernel culls
// them after one 64-byte record), so downstream results are identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
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

// ---------------------------------------------------------------------------------------------
// overlay areas: scripts/road_segmentation/determine_class.py:107-118 get_weighted_scores
//   gpd.overlay(ground_truth, predictions, how='intersection').area  and  ground_truth.area
// area(A and B) = closed integral of x dy over the boundary of the intersection = the pieces of A's boundary inside B plus the
// pieces of B's boundary inside A, every ring taken with its interior on the left.  One warp per candidate pair; lanes take
// edges, cut them at the crossings with the other polygon (next-cut search, no per-edge crossing capacity) and classify each
// piece by its midpoint (even-odd over all rings).  A piece lying ON the other boundary is counted once, from A's side, and
// only when both interiors lie on the same side of it.  Plain binary64: agrees with GEOS' noded overlay to rounding.
// ---------------------------------------------------------------------------------------------
namespace {

// interior-on-the-left sign of every ring: +1 when (counter-clockwise) == (even nesting depth inside its polygon)
__global__ void __launch_bounds__(128) ring_sign_kernel(const PolySet S, int n_rings_total, const int *__restrict__ ring_poly,
                                                        int8_t *__restrict__ sign)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rings_total) return;
    const int v0 = S.ring_off[g], v1 = S.ring_off[g + 1];
    if (v1 - v0 < 3) { sign[g] = 0; return; }
    double a2 = 0.0;
    for (int k = v0; k < v1; k++) {
        const double2 p = S.xy[k], q = S.xy[k + 1 < v1 ? k + 1 : v0];
        a2 = __dadd_rn(a2, __dsub_rn(__dmul_rn(p.x, q.y), __dmul_rn(q.x, p.y)));
    }
    const int poly = ring_poly[g];
    // nesting depth of the ring inside its polygon, from a point of the ring that lies on no other ring: a valid hole may
    // touch its shell (and parts may touch each other) at single vertices, where containment is undecided.  Candidates:
    // the ring's vertices in order, then its edge midpoints; the first one off every other ring decides.
    int depth = 0;
    const int nv = v1 - v0;
    for (int c = 0; c < 2 * nv; c++) {
        double2 rep = S.xy[v0 + (c < nv ? c : c - nv)];
        if (c >= nv) {
            const double2 q = S.xy[c - nv + 1 < nv ? v0 + c - nv + 1 : v0];
            rep.x = __dmul_rn(__dadd_rn(rep.x, q.x), 0.5);
            rep.y = __dmul_rn(__dadd_rn(rep.y, q.y), 0.5);
        }
        int d = 0;
        bool undecided = false;
        for (int h = S.road_ring_off[poly]; h < S.road_ring_off[poly + 1] && !undecided; h++) {
            if (h == g) continue;
            const int w = point_in_rings(S, h, h + 1, rep);
            if (w == 2) undecided = true;
            else if (w == 1) d ^= 1;
        }
        if (!undecided || c == 2 * nv - 1) { depth = d; break; }
    }
    const bool ccw = a2 > 0.0;
    sign[g] = a2 == 0.0 ? 0 : ((ccw != (depth != 0)) ? 1 : -1);
}

// 0 outside, 1 inside, 2 on the boundary; on_edge / on_ring: the first edge that carries the point
__device__ int locate(const PolySet &s, int g0, int g1, double2 p, int &on_edge, int &on_ring)
{
    int inside = 0;
    for (int g = g0; g < g1; g++) {
        const int v0 = s.ring_off[g], v1 = s.ring_off[g + 1];
        for (int k = v0; k < v1; k++) {
            const double2 a = s.xy[k], b = s.xy[k + 1 < v1 ? k + 1 : v0];
            if (a.x == b.x && a.y == b.y) continue;
            const double o = orient(a, b, p);
            if (o == 0.0 && fmin(a.x, b.x) <= p.x && p.x <= fmax(a.x, b.x) && fmin(a.y, b.y) <= p.y && p.y <= fmax(a.y, b.y)) {
                on_edge = k; on_ring = g;
                return 2;
            }
            if ((a.y <= p.y) != (b.y <= p.y)) {
                const bool up = b.y > a.y;
                if ((o > 0.0) == up) inside ^= 1;
            }
        }
    }
    return inside;
}

// integral of x dy over the pieces of the edges of P (rings [g0, g1)) that lie inside Q, interiors on the left.
// count_on: pieces on Q's boundary are counted when both interiors are on the same side (P = A), never (P = B).
__device__ double clipped_boundary_integral(const PolySet &P, int g0, int g1, const int8_t *signP, const PolySet &Q, int h0, int h1,
                                            const int8_t *signQ, bool count_on, int lane)
{
    double sum = 0.0;
    for (int g = g0; g < g1; g++) {
        const int r0 = P.ring_off[g], r1 = P.ring_off[g + 1];
        const double sp = (double)signP[g];
        if (sp == 0.0) continue;
        for (int k = r0 + lane; k < r1; k += 32) {
            const double2 p = P.xy[k], q = P.xy[k + 1 < r1 ? k + 1 : r0];
            if (p.y == q.y) continue;                         // x dy vanishes on horizontal edges
            const double ex = __dsub_rn(q.x, p.x), ey = __dsub_rn(q.y, p.y);
            double t0 = 0.0;
            while (t0 < 1.0) {
                // next cut after t0: the smallest crossing parameter with any edge of Q (collinear edges cut at their ends)
                double t1 = 1.0;
                for (int h = h0; h < h1; h++) {
                    const int s0 = Q.ring_off[h], s1 = Q.ring_off[h + 1];
                    for (int j = s0; j < s1; j++) {
                        const double2 c = Q.xy[j], d = Q.xy[j + 1 < s1 ? j + 1 : s0];
                        const double fx = __dsub_rn(d.x, c.x), fy = __dsub_rn(d.y, c.y);
                        const double den = __dsub_rn(__dmul_rn(ex, fy), __dmul_rn(ey, fx));
                        const double wx = __dsub_rn(c.x, p.x), wy = __dsub_rn(c.y, p.y);
                        if (den != 0.0) {
                            const double t = __ddiv_rn(__dsub_rn(__dmul_rn(wx, fy), __dmul_rn(wy, fx)), den);
                            const double u = __ddiv_rn(__dsub_rn(__dmul_rn(wx, ey), __dmul_rn(wy, ex)), den);
                            if (u >= 0.0 && u <= 1.0 && t > t0 && t < t1) t1 = t;
                        } else if (__dsub_rn(__dmul_rn(wx, ey), __dmul_rn(wy, ex)) == 0.0) {        // collinear
                            const double ee = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
                            const double tc = __ddiv_rn(__dadd_rn(__dmul_rn(wx, ex), __dmul_rn(wy, ey)), ee);
                            const double td = __ddiv_rn(__dadd_rn(__dmul_rn(__dsub_rn(d.x, p.x), ex), __dmul_rn(__dsub_rn(d.y, p.y), ey)), ee);
                            if (tc > t0 && tc < t1) t1 = tc;
                            if (td > t0 && td < t1) t1 = td;
                        }
                    }
                }
                const double tm = __dmul_rn(0.5, __dadd_rn(t0, t1));
                const double2 m = make_double2(__dadd_rn(p.x, __dmul_rn(tm, ex)), __dadd_rn(p.y, __dmul_rn(tm, ey)));
                int oe = 0, orr = 0;
                const int where = locate(Q, h0, h1, m, oe, orr);
                bool take = where == 1;
                if (where == 2 && count_on) {
                    const int s0 = Q.ring_off[orr], s1 = Q.ring_off[orr + 1];
                    const double2 c = Q.xy[oe], d = Q.xy[oe + 1 < s1 ? oe + 1 : s0];
                    const double dot = __dadd_rn(__dmul_rn(ex, __dsub_rn(d.x, c.x)), __dmul_rn(ey, __dsub_rn(d.y, c.y)));
                    take = dot * sp * (double)signQ[orr] > 0.0;
                }
                if (take) {
                    const double xa = __dadd_rn(p.x, __dmul_rn(t0, ex)), xb = __dadd_rn(p.x, __dmul_rn(t1, ex));
                    const double dy = __dmul_rn(__dsub_rn(t1, t0), ey);
                    sum = __dadd_rn(sum, __dmul_rn(sp, __dmul_rn(__dmul_rn(0.5, __dadd_rn(xa, xb)), dy)));
                }
                t0 = t1;
            }
        }
    }
    return sum;
}

__global__ void __launch_bounds__(128) overlay_area_kernel(const PolySet A, const int8_t *__restrict__ signA, const PolySet B,
                                                           const int8_t *__restrict__ signB, const int *__restrict__ pair_a,
                                                           const int *__restrict__ pair_b, int n_pairs, double *__restrict__ out)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const int ia = pair_a[w], ib = pair_b[w];
    const int ga0 = A.road_ring_off[ia], ga1 = A.road_ring_off[ia + 1];
    const int gb0 = B.road_ring_off[ib], gb1 = B.road_ring_off[ib + 1];
    double sum = clipped_boundary_integral(A, ga0, ga1, signA, B, gb0, gb1, signB, true, lane);
    sum = __dadd_rn(sum, clipped_boundary_integral(B, gb0, gb1, signB, A, ga0, ga1, signA, false, lane));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) out[w] = fmax(sum, 0.0);
}

// even-odd area of every polygon: the sum of the signed shoelace areas of its rings, interiors on the left
__global__ void __launch_bounds__(128) poly_area_kernel(const PolySet S, const int8_t *__restrict__ sign, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n) return;
    double tot = 0.0;
    for (int g = S.road_ring_off[i]; g < S.road_ring_off[i + 1]; g++) {
        const int v0 = S.ring_off[g], v1 = S.ring_off[g + 1];
        double a2 = 0.0;
        for (int k = v0; k < v1; k++) {
            const double2 p = S.xy[k], q = S.xy[k + 1 < v1 ? k + 1 : v0];
            a2 = __dadd_rn(a2, __dsub_rn(__dmul_rn(p.x, q.y), __dmul_rn(q.x, p.y)));
        }
        // |a2| / 2 for a ring at even depth, - |a2| / 2 for a hole
        const double mag = __dmul_rn(0.5, fabs(a2));
        const bool ccw = a2 > 0.0;
        tot = __dadd_rn(tot, (sign[g] > 0) == ccw ? mag : -mag);
    }
    out[i] = tot;
}

}  // namespace

// sign_a / sign_b: int8[n_rings] scratch; ring_poly_*: int32[n_rings] polygon of every ring (device)
int launch_overlay_area(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, const int *ring_poly_a, const int *ring_poly_b,
                        int8_t *sign_a, int8_t *sign_b, const int *pair_a, const int *pair_b, int n_pairs, double *area_pair,
                        double *area_a, cudaStream_t st)
{
    PolySet A{(const double2 *)a->xy, a->ring_off, a->road_ring_off, a->road_bbox, a->n_roads};
    PolySet B{(const double2 *)b->xy, b->ring_off, b->road_ring_off, b->road_bbox, b->n_roads};
    if (a->n_rings > 0) {
        ring_sign_kernel<<<(a->n_rings + 127) / 128, 128, 0, st>>>(A, a->n_rings, ring_poly_a, sign_a);
        ctx->launches++;
    }
    if (b->n_rings > 0) {
        ring_sign_kernel<<<(b->n_rings + 127) / 128, 128, 0, st>>>(B, b->n_rings, ring_poly_b, sign_b);
        ctx->launches++;
    }
    if (area_a && a->n_roads > 0) {
        poly_area_kernel<<<(a->n_roads + 127) / 128, 128, 0, st>>>(A, sign_a, area_a);
        ctx->launches++;
    }
    if (n_pairs > 0) {
        const long long blocks = ((long long)n_pairs * 32 + 127) / 128;
        if (blocks > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
        overlay_area_kernel<<<(unsigned)blocks, 128, 0, st>>>(A, sign_a, B, sign_b, pair_a, pair_b, n_pairs, area_pair);
        ctx->launches++;
    }
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// exact reject of the broad phase: gpd.sjoin(tiles, roads) keeps a (tile, road) pair when the geometries INTERSECT
// (scripts/statistical_analysis/statistical_analysis.py:170-171), the bounding-box broad phase keeps a superset.
// One warp per candidate pair; the road (all rings, even-odd) intersects the closed tile rectangle iff
//   an edge of the road touches the rectangle (an endpoint inside it, or -- separating axes of a segment and a box -- the
//   boxes overlap and the rectangle's corners are not strictly on one side of the edge's line), or
//   the rectangle lies inside the road: no edge touches it and one corner is inside (crossing parity over all edges).
// Binary64 orientation signs: can differ from GEOS only for contacts within rounding distance.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(128) intersects_kernel(const PolySet S, const double *__restrict__ tile_ext,
                                                         const int *__restrict__ road_pair_off, const int *__restrict__ pair_tile,
                                                         int n_pairs, uint8_t *__restrict__ keep)
{
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    int lo = 0, hi = S.n;                        // largest road with road_pair_off[road] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(road_pair_off + mid) <= p) lo = mid;
        else hi = mid;
    }
    const int road = lo;
    const double *e = tile_ext + 4 * (size_t)pair_tile[p];
    const double x0 = e[0], y0 = e[1], x1 = e[2], y1 = e[3];
    const int g0 = S.road_ring_off[road], g1 = S.road_ring_off[road + 1];
    bool hit = false;
    int parity = 0;                              // of the corner (x0, y0)
    for (int g = g0; g < g1 && !__any_sync(0xffffffffu, hit); g++) {
        const int v0 = S.ring_off[g], v1 = S.ring_off[g + 1];
        for (int base = v0; base < v1 && !__any_sync(0xffffffffu, hit); base += 32) {
            const int k = base + lane;
            if (k < v1) {
                const double2 a = S.xy[k], b = S.xy[k + 1 < v1 ? k + 1 : v0];
                const bool a_in = a.x >= x0 && a.x <= x1 && a.y >= y0 && a.y <= y1;
                if (a_in) hit = true;
                else if (fmin(a.x, b.x) <= x1 && fmax(a.x, b.x) >= x0 && fmin(a.y, b.y) <= y1 && fmax(a.y, b.y) >= y0 &&
                         !(a.x == b.x && a.y == b.y)) {
                    const double o1 = orient(a, b, make_double2(x0, y0)), o2 = orient(a, b, make_double2(x1, y0));
                    const double o3 = orient(a, b, make_double2(x1, y1)), o4 = orient(a, b, make_double2(x0, y1));
                    if (!((o1 > 0.0 && o2 > 0.0 && o3 > 0.0 && o4 > 0.0) || (o1 < 0.0 && o2 < 0.0 && o3 < 0.0 && o4 < 0.0))) hit = true;
                }
                if ((a.y <= y0) != (b.y <= y0)) {                 // the edge spans the horizontal line through the corner
                    const double o = orient(a, b, make_double2(x0, y0));
                    if ((o > 0.0) == (b.y > a.y)) parity ^= 1;
                }
            }
        }
    }
    hit = __any_sync(0xffffffffu, hit);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) parity ^= __shfl_xor_sync(0xffffffffu, parity, o);
    if (lane == 0) keep[p] = (hit || (parity & 1)) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// broad phase for tiles that are NOT on a lattice: the tiles are binned into a uniform grid (cell ~ a tile), sorted by
// cell, and every road looks up the cells under its bounding box.  A (road, tile) pair is reported from exactly one cell, the
// one holding the lower-left corner of the intersection of the two boxes.
// ---------------------------------------------------------------------------------------------
struct GridArgs {
    double X0, Y0, cw, ch;
    int nx, ny;
};

__device__ __forceinline__ int cell_of(double v, double o, double c, int n)
{
    const double f = floor((v - o) / c);
    return f < 0.0 ? 0 : (f > (double)(n - 1) ? n - 1 : (int)f);
}

__global__ void __launch_bounds__(256) grid_tile_kernel(const GridArgs G, const double *__restrict__ ext, int n_tiles,
                                                        const unsigned long long *__restrict__ off, int *__restrict__ cnt,
                                                        int *__restrict__ keys, int *__restrict__ vals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const double *e = ext + 4 * (size_t)t;
    if (!(e[0] <= e[2]) || !(e[1] <= e[3])) { if (cnt) cnt[t] = 0; return; }
    const int ix0 = cell_of(e[0], G.X0, G.cw, G.nx), ix1 = cell_of(e[2], G.X0, G.cw, G.nx);
    const int iy0 = cell_of(e[1], G.Y0, G.ch, G.ny), iy1 = cell_of(e[3], G.Y0, G.ch, G.ny);
    if (cnt) { cnt[t] = (ix1 - ix0 + 1) * (iy1 - iy0 + 1); return; }
    unsigned long long k = off[t];
    for (int iy = iy0; iy <= iy1; iy++)
        for (int ix = ix0; ix <= ix1; ix++, k++) { keys[k] = iy * G.nx + ix; vals[k] = t; }
}

__global__ void __launch_bounds__(256) grid_cell_start_kernel(const int *__restrict__ keys, long long n_entries, int n_cells,
                                                              int *__restrict__ start)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    long long lo = 0, hi = n_entries;            // first entry with key >= c
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < c) lo = mid + 1;
        else hi = mid;
    }
    start[c] = (int)lo;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) grid_road_kernel(const GridArgs G, const double *__restrict__ bbox, int n_roads,
                                                        const double *__restrict__ ext, const int *__restrict__ start,
                                                        const int *__restrict__ vals, int *__restrict__ cnt,
                                                        const int *__restrict__ off, long long capacity, int *__restrict__ pair_tile)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_roads) return;
    const double *b = bbox + 4 * (size_t)r;
    int c = 0;
    if ((b[0] <= b[2]) && (b[1] <= b[3])) {
        const int ix0 = cell_of(b[0], G.X0, G.cw, G.nx), ix1 = cell_of(b[2], G.X0, G.cw, G.nx);
        const int iy0 = cell_of(b[1], G.Y0, G.ch, G.ny), iy1 = cell_of(b[3], G.Y0, G.ch, G.ny);
        const long long base = WRITE ? off[r] : 0;
        for (int iy = iy0; iy <= iy1; iy++)
            for (int ix = ix0; ix <= ix1; ix++) {
                const int cell = iy * G.nx + ix;
                for (int k = start[cell]; k < start[cell + 1]; k++) {
                    const int t = vals[k];
                    const double *e = ext + 4 * (size_t)t;
                    if (!(b[0] <= e[2] && b[2] >= e[0] && b[1] <= e[3] && b[3] >= e[1])) continue;
                    // reported once: from the cell of the lower-left corner of the boxes' intersection
                    if (cell_of(fmax(b[0], e[0]), G.X0, G.cw, G.nx) != ix || cell_of(fmax(b[1], e[1]), G.Y0, G.ch, G.ny) != iy) continue;
                    if (WRITE && base + c < capacity) {
                        int j = c - 1;                         // keep the road's tiles sorted by index
                        while (j >= 0 && pair_tile[base + j] > t) { pair_tile[base + j + 1] = pair_tile[base + j]; j--; }
                        pair_tile[base + j + 1] = t;
                    }
                    c++;
                }
            }
    }
    if (!WRITE) cnt[r] = c;
}

}  // namespace

int launch_intersects(rs_ctx *ctx, const rs_roads *roads, const double *tile_ext_dev, const int *road_pair_off_dev,
                      const int *pair_tile_dev, int n_pairs, uint8_t *keep_dev, cudaStream_t st)
{
    if (n_pairs <= 0) return RS_OK;
    PolySet S{(const double2 *)roads->xy, roads->ring_off, roads->road_ring_off, roads->road_bbox, roads->n_roads};
    const long long blocks = ((long long)n_pairs * 32 + 127) / 128;
    if (blocks > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
    intersects_kernel<<<(unsigned)blocks, 128, 0, st>>>(S, tile_ext_dev, road_pair_off_dev, pair_tile_dev, n_pairs, keep_dev);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// phase 0: bins the tiles (stage[12..15], wide_tmp), counts per road -> road_pair_off_dev (exclusive scan, [n_roads] = total);
// phase 1: writes pair_tile_dev.  The grid is rebuilt in both phases (the call protocol is stateless).
int launch_pairs_grid(rs_ctx *ctx, const double *bbox_dev, int n_roads, const double *ext_dev, int n_tiles, double X0, double Y0,
                      double cw, double ch, int nx, int ny, int *road_pair_off_dev, int *pair_tile_dev, long long capacity, int phase,
                      cudaStream_t st)
{
    GridArgs G{X0, Y0, cw, ch, nx, ny};
    const int n_cells = nx * ny;
    int rc;
    if ((rc = ensure(ctx, ctx->stage[12], sizeof(int) * ((size_t)(n_tiles > n_roads ? n_tiles : n_roads) + 1)))) return rc;   // counts
    if ((rc = ensure(ctx, ctx->stage[13], sizeof(unsigned long long) * ((size_t)n_tiles + 1)))) return rc;                   // tile offsets
    if ((rc = ensure(ctx, ctx->stage[14], sizeof(int) * ((size_t)n_cells + 2)))) return rc;                                   // cell starts
    int *cnt = (int *)ctx->stage[12].p, *start = (int *)ctx->stage[14].p;
    unsigned long long *toff = (unsigned long long *)ctx->stage[13].p;
    long long n_entries = 0;
    if (n_tiles > 0) {
        grid_tile_kernel<<<(n_tiles + 255) / 256, 256, 0, st>>>(G, ext_dev, n_tiles, nullptr, cnt, nullptr, nullptr);
        ctx->launches++;
        if ((rc = launch_fstats_offsets(ctx, (const uint32_t *)cnt, n_tiles, toff, st))) return rc;
        unsigned long long tot = 0;
        RS_CUDA_OK(ctx, cudaMemcpyAsync(&tot, toff + n_tiles, sizeof(tot), cudaMemcpyDeviceToHost, st));
        RS_CUDA_OK(ctx, cudaStreamSynchronize(st));
        if (tot > 0x7fffffffull) return RS_ERR_UNSUPPORTED;
        n_entries = (long long)tot;
    }
    if ((rc = ensure(ctx, ctx->stage[15], sizeof(int) * 4 * ((size_t)n_entries + 1)))) return rc;    // keys | vals | sorted keys | sorted vals
    int *keys = (int *)ctx->stage[15].p, *vals = keys + n_entries, *skeys = vals + n_entries, *svals = skeys + n_entries;
    if (n_entries > 0) {
        grid_tile_kernel<<<(n_tiles + 255) / 256, 256, 0, st>>>(G, ext_dev, n_tiles, toff, nullptr, keys, vals);
        ctx->launches++;
        size_t tmp = 0;
        RS_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys, skeys, vals, svals, (int)n_entries, 0, 32, st));
        if ((rc = ensure(ctx, ctx->wide_tmp, tmp))) return rc;
        RS_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(ctx->wide_tmp.p, tmp, keys, skeys, vals, svals, (int)n_entries, 0, 32, st));
        ctx->launches++;
    }
    grid_cell_start_kernel<<<(n_cells + 256) / 256, 256, 0, st>>>(skeys, n_entries, n_cells, start);
    ctx->launches++;
    if (n_roads > 0) {
        if (phase == 0) {
            grid_road_kernel<false><<<(n_roads + 255) / 256, 256, 0, st>>>(G, bbox_dev, n_roads, ext_dev, start, svals, cnt, nullptr, 0, nullptr);
            ctx->launches++;
            size_t tmp = 0;
            RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, road_pair_off_dev, n_roads, st));
            if ((rc = ensure(ctx, ctx->wide_tmp, tmp))) return rc;
            RS_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->wide_tmp.p, tmp, cnt, road_pair_off_dev, n_roads, st));
        } else {
            grid_road_kernel<true><<<(n_roads + 255) / 256, 256, 0, st>>>(G, bbox_dev, n_roads, ext_dev, start, svals, nullptr, road_pair_off_dev,
                                                                          capacity, pair_tile_dev);
            ctx->launches++;
        }
    }
    if (phase == 0) {
        pairs_total_kernel<<<1, 32, 0, st>>>(cnt, road_pair_off_dev, n_roads);
        ctx->launches++;
    }
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

// ---------------------------------------------------------------------------------------------
// clip_labels: scripts/road_segmentation/determine_class.py:62-95 -- every label cut to every tile it intersects, the tile
// scaled by 0.99 about its centre (old_geo.intersection(scale(tile, 0.99, 0.99))).  One thread per (pair, ring): the
// re-entrant Sutherland-Hodgman pipeline (four half-plane stages, each keeping only its first and previous point) streams the
// ring's vertices through the rectangle without intermediate buffers; a count pass sizes the output, a write pass fills it.
// Pieces of a concave ring that leave and re-enter the rectangle stay connected by zero-width runs along its edge, which carry
// no area under the even-odd rule (GEOS returns them as separate parts: same point set, same areas, same raster).
// ---------------------------------------------------------------------------------------------
namespace {

struct ClipStage {
    double2 first, prev;
    bool has;
};

struct ClipSink {
    double2 *out;        // nullptr: count only
    int n;
    double2 first;
    __device__ __forceinline__ void put(double2 p)
    {
        if (n == 0) first = p;
        if (out) out[n] = p;
        n++;
    }
};

__device__ __forceinline__ bool clip_inside(int st, double2 p, const double *r)
{
    return st == 0 ? p.x >= r[0] : st == 1 ? p.x <= r[2] : st == 2 ? p.y >= r[1] : p.y <= r[3];
}

// the point where the edge p -> q meets the boundary line of stage st (same expression as the host clip: t along p -> q, the
// clipped coordinate set to the bound exactly)
__device__ __forceinline__ double2 clip_cross(int st, double2 p, double2 q, const double *r)
{
    const double bound = st == 0 ? r[0] : st == 1 ? r[2] : st == 2 ? r[1] : r[3];
    double2 c;
    if (st < 2) {
        const double t = __ddiv_rn(__dsub_rn(bound, p.x), __dsub_rn(q.x, p.x));
        c.x = bound;
        c.y = __dadd_rn(p.y, __dmul_rn(t, __dsub_rn(q.y, p.y)));
    } else {
        const double t = __ddiv_rn(__dsub_rn(bound, p.y), __dsub_rn(q.y, p.y));
        c.x = __dadd_rn(p.x, __dmul_rn(t, __dsub_rn(q.x, p.x)));
        c.y = bound;
    }
    return c;
}

__device__ void clip_feed(ClipStage *stg, int st, double2 p, const double *r, ClipSink &sink)
{
    // iterative form of the re-entrant pipeline: a point entering stage st may release up to two points into stage st + 1
    double2 queue[16];
    int qst[16], qn = 0;
    queue[qn] = p; qst[qn++] = st;
    while (qn > 0) {
        const double2 v = queue[--qn];
        const int s_ = qst[qn];
        if (s_ == 4) { sink.put(v); continue; }
        ClipStage &g = stg[s_];
        const bool vin = clip_inside(s_, v, r);
        double2 rel[2];
        int nr = 0;
        if (!g.has) { g.first = v; g.has = true; }
        else if (clip_inside(s_, g.prev, r) != vin) rel[nr++] = clip_cross(s_, g.prev, v, r);
        g.prev = v;
        if (vin) rel[nr++] = v;
        for (int k = nr - 1; k >= 0; k--) { queue[qn] = rel[k]; qst[qn++] = s_ + 1; }      // LIFO: push in reverse to keep the order
    }
}

__global__ void __launch_bounds__(128) clip_rings_kernel(const PolySet S, const int *__restrict__ pair_label, const double *__restrict__ rect,
                                                         const long long *__restrict__ pair_ring_off, int n_pairs, long long n_pair_rings,
                                                         int *__restrict__ ring_count, const long long *__restrict__ ring_vert_off,
                                                         double2 *__restrict__ xy_out)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pair_rings) return;
    int lo = 0, hi = n_pairs;                    // largest pair with pair_ring_off[pair] <= q
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pair_ring_off[mid] <= q) lo = mid;
        else hi = mid;
    }
    const int pair = lo, label = pair_label[pair];
    const int g = S.road_ring_off[label] + (int)(q - pair_ring_off[pair]);
    const double *r = rect + 4 * (size_t)pair;
    int v0 = S.ring_off[g], v1 = S.ring_off[g + 1];
    if (v1 - v0 > 1 && S.xy[v0].x == S.xy[v1 - 1].x && S.xy[v0].y == S.xy[v1 - 1].y) v1--;      // closed ring: the repeat is dropped
    ClipStage stg[4];
#pragma unroll
    for (int k = 0; k < 4; k++) stg[k].has = false;
    ClipSink sink{xy_out ? xy_out + ring_vert_off[q] : nullptr, 0, make_double2(0.0, 0.0)};
    for (int k = v0; k < v1; k++) clip_feed(stg, 0, S.xy[k], r, sink);
    // close the stages in order: the edge from a stage's last point back to its first one
    for (int st = 0; st < 4; st++) {
        ClipStage &gs = stg[st];
        if (!gs.has) continue;
        if (clip_inside(st, gs.prev, r) != clip_inside(st, gs.first, r)) {
            const double2 c = clip_cross(st, gs.prev, gs.first, r);
            if (st == 3) sink.put(c);
            else clip_feed(stg, st + 1, c, r, sink);
        }
    }
    if (sink.n >= 3) sink.put(sink.first);       // closed, like shapely's rings
    else sink.n = 0;                             // fewer than three points: the ring misses the rectangle
    if (!xy_out) ring_count[q] = sink.n;
}

}  // namespace

int launch_clip_rings(rs_ctx *ctx, const rs_roads *labels, const int *pair_label, const double *rect, const long long *pair_ring_off,
                      int n_pairs, long long n_pair_rings, int *ring_count, const long long *ring_vert_off, double *xy_out,
                      cudaStream_t st)
{
    if (n_pair_rings <= 0) return RS_OK;
    PolySet S{(const double2 *)labels->xy, labels->ring_off, labels->road_ring_off, labels->road_bbox, labels->n_roads};
    const long long blocks = (n_pair_rings + 127) / 128;
    if (blocks > 0x7fffffffLL) return RS_ERR_UNSUPPORTED;
    clip_rings_kernel<<<(unsigned)blocks, 128, 0, st>>>(S, pair_label, rect, pair_ring_off, n_pairs, n_pair_rings, ring_count,
                                                       ring_vert_off, (double2 *)xy_out);
    ctx->launches++;
    RS_CUDA_OK(ctx, cudaGetLastError());
    return RS_OK;
}

}  // namespace rs
