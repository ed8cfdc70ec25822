// Device helpers shared by the rasterizing kernels (rs_zonal.cu, rs_wide.cu): mbarrier / TMA bulk copy, the per-pair window
// geometry (rasterio geometry_window + window_transform + GDALInvGeoTransform), XU-free integer <-> binary64 conversions and
// the bit tricks of the even-odd mask.  Translation units that include this are compiled with -fmad=false: the rounding of
// every binary64 operation is part of the specification.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rs_internal.h"

namespace rs {

constexpr unsigned FULL = 0xffffffffu;
constexpr int ROWS_ITEM = 256;  // zonal_kernel: a window taller than this is split by rows over several items
constexpr int RS_MAX_WINDOW_W = 2048;   // widest window of one (road, raster) pair; the raster itself may be wider

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

// per-pair geometry: integer window inside the tile + world->window-pixel transform
struct alignas(16) PairGeom {       // 64 bytes, computed once per pair by pair_geom_kernel
    double inv0, inv1, inv3, inv5;
    int col_off, row_off, w, h;     // window inside the tile (what is masked and read)
    int xshift, yshift, wu;         // RS_WINDOW_BOUNDLESS: the rasterized window starts xshift columns / yshift rows
                                    // before the visible one and is wu columns wide (0, 0, w otherwise)
    int status;                     // 1: rasterize; 0: shapes do not overlap the raster; < 0: rs_status
};

// rasterio geometry_window + window_transform + GDALInvGeoTransform, from the road bbox
// (px/py are monotone in x/y for north-up transforms, so the vertex-wise bounds rasterio takes
// are attained at the bbox corners).  Returns 0 (shapes do not overlap raster), 1, or <0.
__device__ __forceinline__ int pair_geometry(const double *__restrict__ gt, const double *__restrict__ bb, int W, int H,
                                             int window_mode, int border, PairGeom &g)
{
    const double sa = gt[0], sb = gt[1], sc = gt[2], sd = gt[3], se = gt[4], sf = gt[5];
    if (sb != 0.0 || sd != 0.0 || sa == 0.0 || se == 0.0) return RS_ERR_ROTATED;
    if (window_mode == RS_WINDOW_FULL) {
        if (W > RS_MAX_WINDOW_W) return RS_ERR_UNSUPPORTED;
        g.col_off = 0; g.row_off = 0; g.w = W; g.h = H;
        g.xshift = 0; g.yshift = 0; g.wu = W;
        g.inv0 = __ddiv_rn(-sc, sa); g.inv1 = __ddiv_rn(1.0, sa);
        g.inv3 = __ddiv_rn(-sf, se); g.inv5 = __ddiv_rn(1.0, se);
        if (border > 0) {          // keep the transform of the whole raster, look at the inner rectangle only
            if (2 * border >= W || 2 * border >= H) return 0;
            g.col_off = border; g.row_off = border; g.w = W - 2 * border; g.h = H - 2 * border;
            g.xshift = border; g.yshift = border;
        }
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
    if (g.w > RS_MAX_WINDOW_W) return RS_ERR_UNSUPPORTED;      // a row's mask must fit the team's shared memory
    g.xshift = 0; g.yshift = 0; g.wu = g.w;
    double xo = (double)ic0, yo = (double)ir0;
    if (window_mode == RS_WINDOW_BOUNDLESS) {       // rasterstats: the window keeps its unclipped origin
        if (c0 < -1.0e9 || r0 < -1.0e9 || ww > 2.0e9) return RS_ERR_UNSUPPORTED;
        xo = c0; yo = r0;
        g.xshift = ic0 - (int)c0; g.yshift = ir0 - (int)r0; g.wu = (int)fmin(ww, 2.0e9);
    }
    // transform * Affine.translation(xo, yo), then GDALInvGeoTransform (north-up branch)
    const double wa = __dadd_rn(__dmul_rn(sa, 1.0), __dmul_rn(sb, 0.0));
    const double wc = __dadd_rn(__dadd_rn(__dmul_rn(sa, xo), __dmul_rn(sb, yo)), sc);
    const double we = __dadd_rn(__dmul_rn(sd, 0.0), __dmul_rn(se, 1.0));
    const double wf = __dadd_rn(__dadd_rn(__dmul_rn(sd, xo), __dmul_rn(se, yo)), sf);
    g.inv0 = __ddiv_rn(-wc, wa); g.inv1 = __ddiv_rn(1.0, wa);
    g.inv3 = __ddiv_rn(-wf, we); g.inv5 = __ddiv_rn(1.0, we);
    if (border > 0) {
        // the rasterized window keeps its origin and size; only the pixels at least `border` away from the tile edge
        // are looked at (determine_class.clip_labels: labels clipped to the tile scaled by 0.99)
        const int c_lo = max(g.col_off, border), c_hi = min(g.col_off + g.w, W - border);
        const int r_lo = max(g.row_off, border), r_hi = min(g.row_off + g.h, H - border);
        if (c_hi <= c_lo || r_hi <= r_lo) return 0;
        g.xshift += c_lo - g.col_off; g.yshift += r_lo - g.row_off;
        g.col_off = c_lo; g.row_off = r_lo; g.w = c_hi - c_lo; g.h = r_hi - r_lo;
    }
    return 1;
}

// Integer <-> binary64 without the conversion unit (F2I / I2F / FRND run on the quarter-rate XU pipe):
// adding 1.5 * 2^52 leaves rint(v) in the low mantissa word; every step is an exact or correctly rounded
// binary64 add, so the results below are the same integers floor()/ceil()/(int) casts would give.
constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
__device__ __forceinline__ int rint_magic(double v, double &t)      // |v| < 2^31; t = (double)result
{
    const double sft = __dadd_rn(v, MAGIC);
    t = __dsub_rn(sft, MAGIC);
    return __double2loint(sft);
}
__device__ __forceinline__ double int2double_magic(int y)          // exact for every int32
{
    return __dsub_rn(__hiloint2double(0x43300000, y ^ (int)0x80000000), 4503601774854144.0);   // 2^52 + 2^31
}
// smallest integer y with y + 0.5 >= v (exact comparisons); the largest y with y + 0.5 < v is that minus 1
__device__ __forceinline__ int first_row_ge(double v)
{
    const double vc = fmin(fmax(v, -4.0), 1.0e6);
    double t;
    const int ti = rint_magic(__dsub_rn(vc, 0.5), t);
    if (__dadd_rn(t, 0.5) < vc) return ti + 1;
    if (__dsub_rn(t, 0.5) >= vc) return ti - 1;
    return ti;
}
__device__ __forceinline__ int last_row_lt(double v) { return first_row_ge(v) - 1; }

__device__ __forceinline__ uint32_t prefix_xor32(uint32_t t)
{
    t ^= t << 1; t ^= t << 2; t ^= t << 4; t ^= t << 8; t ^= t << 16;
    return t;
}
// bit 7 of every non-zero byte
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t m) { return (((m & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m) & 0x80808080u; }


// launch of pair_geom_kernel (rs_zonal.cu): one PairGeom per pair into ctx->pgeom
int launch_pair_geom(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode, int border,
                     int tile_lo, int tile_hi, cudaStream_t st);

}  // namespace rs
