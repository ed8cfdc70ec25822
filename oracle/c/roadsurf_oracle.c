/*
 * CPU oracle, plain C restatement of the raster side of the hot path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker and the timed CPU
 * baseline.  Nothing shipped links or calls this file.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in third-party dependencies of
 * the reference (GDAL 3.0.4, rasterio 1.3.2, affine 2.3.1 -- requirements.txt:77,211,9)
 * that are neither vendored nor installable here, and the reference ships no golden
 * vectors.  Anchors: scripts/functions/fct_misc.py:77 (rasterio.mask.mask, crop=True),
 * scripts/sandbox/add_tile_mask.py:112-113 (features.rasterize).  Upstream routines
 * restated:
 *   GDAL alg/llrasterize.cpp     GDALdllImageFilledPolygon, gvBurnScanline
 *   GDAL gcore/gdal_misc.cpp     GDALInvGeoTransform (north-up branch)
 *   GDAL alg/gdaltransformer.cpp GDALGenImgProjTransform (dst geotransform only)
 *   affine                       Affine.__invert__ / __mul__
 *   rasterio features.py/mask.py bounds, geometry_window, raster_geometry_mask
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: x86-64 GDAL
 * builds evaluate these expressions with separate multiply and add).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

static void burn_scanline(uint8_t *mask, int W, int y, int xs, int xe)
{
    if (xs > xe) return;
    if (xs < 0) xs = 0;
    if (xe >= W) xe = W - 1;
    if (xs > xe) return;
    memset(mask + (size_t)y * W + xs, 1, (size_t)(xe - xs + 1));
}

/* GDALdllImageFilledPolygon: X/Y in pixel/line space, all rings concatenated. */
void orc_fill_polygon(int W, int H, int nparts, const int *part_size,
                      const double *X, const double *Y, uint8_t *mask)
{
    int n = 0;
    for (int p = 0; p < nparts; p++) n += part_size[p];
    if (n == 0 || W <= 0 || H <= 0) return;
    int *ints = (int *)malloc(sizeof(int) * (size_t)n);
    double dminy = Y[0], dmaxy = Y[0];
    for (int i = 1; i < n; i++) {
        if (Y[i] < dminy) dminy = Y[i];
        if (Y[i] > dmaxy) dmaxy = Y[i];
    }
    /* (int) casts of far-away values are undefined in C; clamp first, same result in range */
    if (dminy < -2e9) dminy = -2e9;
    if (dmaxy > 2e9) dmaxy = 2e9;
    if (dminy > 2e9 || dmaxy < -2e9) { free(ints); return; }
    int miny = (int)dminy, maxy = (int)dmaxy;
    if (miny < 0) miny = 0;
    if (maxy >= H) maxy = H - 1;
    const int minx = 0, maxx = W - 1;
    for (int y = miny; y <= maxy; y++) {
        const double dy = y + 0.5;
        int nints = 0, part = 0, partoffset = 0;
        for (int i = 0; i < n; i++) {
            while (part < nparts && i == partoffset + part_size[part]) {
                partoffset += part_size[part];
                part++;
            }
            int ind1, ind2;
            if (i == partoffset) { ind1 = partoffset + part_size[part] - 1; ind2 = partoffset; }
            else                 { ind1 = i - 1; ind2 = i; }
            double dy1 = Y[ind1], dy2 = Y[ind2];
            if ((dy1 < dy && dy2 < dy) || (dy1 > dy && dy2 > dy)) continue;
            double dx1, dx2;
            if (dy1 < dy2) { dx1 = X[ind1]; dx2 = X[ind2]; }
            else if (dy1 > dy2) { dy2 = Y[ind1]; dy1 = Y[ind2]; dx2 = X[ind1]; dx1 = X[ind2]; }
            else {
                if (X[ind1] > X[ind2]) {
                    double a = floor(X[ind2] + 0.5), b = floor(X[ind1] + 0.5);
                    if (a > maxx || b <= minx) continue;   /* same test, done in double */
                    int hx1 = (a < -1.0) ? -1 : (int)a;
                    int hx2 = (b > (double)W) ? W : (int)b;
                    burn_scanline(mask, W, y, hx1, hx2 - 1);
                }
                continue;
            }
            if (dy < dy2 && dy >= dy1) {
                double intersect = (dy - dy1) * (dx2 - dx1) / (dy2 - dy1) + dx1;
                double r = floor(intersect + 0.5);
                /* order-preserving clamp into [-1, W]: identical burns, no int overflow */
                if (r < -1.0) r = -1.0;
                if (r > (double)W) r = (double)W;
                ints[nints++] = (int)r;
            }
        }
        qsort(ints, (size_t)nints, sizeof(int), cmp_int);
        for (int i = 0; i + 1 < nints; i += 2)
            if (ints[i] <= maxx && ints[i + 1] > minx)
                burn_scanline(mask, W, y, ints[i], ints[i + 1] - 1);
    }
    free(ints);
}

/* Affine.__invert__ */
static void affine_invert(const double t[6], double r[6])
{
    double sa = t[0], sb = t[1], sc = t[2], sd = t[3], se = t[4], sf = t[5];
    double det = sa * se - sb * sd;
    double idet = 1.0 / det;
    double ra = se * idet, rb = -sb * idet, rd = -sd * idet, re = sa * idet;
    r[0] = ra; r[1] = rb; r[2] = -sc * ra - sf * rb;
    r[3] = rd; r[4] = re; r[5] = -sc * rd - sf * re;
}

/* rasterio geometry_window (pad 0, not boundless).  xy = all vertices of the geometry.
 * win = {col_off, row_off, w, h}.  Returns 0 where rasterio raises WindowError. */
int orc_geometry_window(const double t[6], int n, const double *xy, int W, int H, int win[4])
{
    if (n <= 0) return 0;
    double inv[6];
    affine_invert(t, inv);
    double left = 0, right = 0, top = 0, bottom = 0;
    for (int i = 0; i < n; i++) {
        double vx = xy[2 * i], vy = xy[2 * i + 1];
        double px = vx * inv[0] + vy * inv[1] + inv[2];
        double py = vx * inv[3] + vy * inv[4] + inv[5];
        if (i == 0) { left = right = px; top = bottom = py; }
        else {
            if (px < left) left = px;
            if (px > right) right = px;
            if (py < top) top = py;
            if (py > bottom) bottom = py;
        }
    }
    double lim = 1e9;
    if (left < -lim) left = -lim;
    if (right > lim) right = lim;
    if (top < -lim) top = -lim;
    if (bottom > lim) bottom = lim;
    if (left > lim || right < -lim || top > lim || bottom < -lim) return 0;
    long row_start = (long)floor(top), row_stop = (long)ceil(bottom);
    long col_start = (long)floor(left), col_stop = (long)ceil(right);
    long w = col_stop - col_start; if (w < 0) w = 0;
    long h = row_stop - row_start; if (h < 0) h = 0;
    long r0 = row_start, r1 = row_start + h, c0 = col_start, c1 = col_start + w;
    if (r0 >= H || r1 <= 0 || c0 >= W || c1 <= 0) return 0;
    if (r0 < 0) r0 = 0;
    if (r1 > H) r1 = H;
    if (c0 < 0) c0 = 0;
    if (c1 > W) c1 = W;
    win[0] = (int)c0; win[1] = (int)r0; win[2] = (int)(c1 - c0); win[3] = (int)(r1 - r0);
    return 1;
}

/* rasterio.features.rasterize(shapes, out_shape=(H,W), transform=t): north-up only. */
int orc_rasterize(const double t[6], int nparts, const int *part_size, const double *xy,
                  int W, int H, uint8_t *mask)
{
    if (t[1] != 0.0 || t[3] != 0.0 || t[0] == 0.0 || t[4] == 0.0) return -1;
    int n = 0;
    for (int p = 0; p < nparts; p++) n += part_size[p];
    /* gt = (c, a, b, f, d, e); GDALInvGeoTransform north-up branch */
    double inv0 = -t[2] / t[0], inv1 = 1.0 / t[0], inv2 = 0.0;
    double inv3 = -t[5] / t[4], inv4 = 0.0, inv5 = 1.0 / t[4];
    double *X = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *Y = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        X[i] = inv0 + x * inv1 + y * inv2;
        Y[i] = inv3 + x * inv4 + y * inv5;
    }
    orc_fill_polygon(W, H, nparts, part_size, X, Y, mask);
    free(X);
    free(Y);
    return 0;
}

/* rasterio.mask.raster_geometry_mask(crop=True) expanded to the full tile: mask[H][W]
 * gets 1 on selected pixels.  scratch must hold W*H bytes.  Returns #window pixels (0 if
 * the shapes do not overlap the raster). */
static int pair_mask(const double t[6], int nparts, const int *part_size, const double *xy, int nverts,
                     int W, int H, uint8_t *scratch, int win[4])
{
    if (!orc_geometry_window(t, nverts, xy, W, H, win)) return 0;
    int w = win[2], h = win[3];
    if (w <= 0 || h <= 0) return 0;
    /* transform * Affine.translation(col_off, row_off) */
    double wt[6];
    double xo = (double)win[0], yo = (double)win[1];
    wt[0] = t[0] * 1.0 + t[1] * 0.0;
    wt[1] = t[0] * 0.0 + t[1] * 1.0;
    wt[2] = t[0] * xo + t[1] * yo + t[2];
    wt[3] = t[3] * 1.0 + t[4] * 0.0;
    wt[4] = t[3] * 0.0 + t[4] * 1.0;
    wt[5] = t[3] * xo + t[4] * yo + t[5];
    memset(scratch, 0, (size_t)w * h);
    if (orc_rasterize(wt, nparts, part_size, xy, w, h, scratch) != 0) return -1;
    return w * h;
}

int orc_pair_mask_full(const double t[6], int nparts, const int *part_size, const double *xy,
                       int W, int H, uint8_t *mask /* [H][W], zeroed by caller */)
{
    int n = 0;
    for (int p = 0; p < nparts; p++) n += part_size[p];
    uint8_t *scratch = (uint8_t *)malloc((size_t)W * H + 1);
    int win[4];
    int k = pair_mask(t, nparts, part_size, xy, n, W, H, scratch, win);
    if (k > 0)
        for (int y = 0; y < win[3]; y++)
            memcpy(mask + (size_t)(win[1] + y) * W + win[0], scratch + (size_t)y * win[2], (size_t)win[2]);
    free(scratch);
    return k < 0 ? -1 : 0;
}

/*
 * Per-road accumulators over a road-major pair list (roads [road_begin, road_end)).
 *   xy[V][2], ring_off[NR+1] (vertex offsets), road_ring_off[R+1] (ring offsets)
 *   road_pair_off[R+1], pair_tile[P]
 *   tiles: [T][H][W][C], elem_bytes 1 (uint8) or 2 (uint16 rescaled with scale_k/scale_off,
 *          f32 working precision if rescale_f32)
 *   joint = 0: hist[R][C][256] per band; n_allzero[R] = in-mask pixels with all bands 0
 *   joint = 1 (C == 2, class/score planes): hist[R][3][256] indexed [min(class,2)... class][score]
 */
int orc_zonal_accumulate(const double *xy, const int *ring_off, const int *road_ring_off,
                         const int *road_pair_off, const int *pair_tile,
                         const void *tiles, const double *tile_gt, int H, int W, int C, int elem_bytes,
                         const double *scale_k, const double *scale_off, int rescale_f32, int joint,
                         int road_begin, int road_end, uint64_t *hist, uint64_t *n_allzero, uint64_t *min_zero)
{
    uint8_t *scratch = (uint8_t *)malloc((size_t)W * H + 1);
    int maxrings = 0;
    for (int r = road_begin; r < road_end; r++) {
        int k = road_ring_off[r + 1] - road_ring_off[r];
        if (k > maxrings) maxrings = k;
    }
    int *psize = (int *)malloc(sizeof(int) * (size_t)(maxrings + 1));
    const int HC = joint ? 3 : C;
    int rc = 0;
    for (int r = road_begin; r < road_end && rc == 0; r++) {
        int g0 = road_ring_off[r], g1 = road_ring_off[r + 1];
        int v0 = ring_off[g0], v1 = ring_off[g1];
        for (int g = g0; g < g1; g++) psize[g - g0] = ring_off[g + 1] - ring_off[g];
        uint64_t *hr = hist + (size_t)r * HC * 256;
        for (int p = road_pair_off[r]; p < road_pair_off[r + 1]; p++) {
            int t = pair_tile[p];
            int win[4];
            int k = pair_mask(tile_gt + 6 * (size_t)t, g1 - g0, psize, xy + 2 * (size_t)v0, v1 - v0, W, H, scratch, win);
            if (k < 0) { rc = -1; break; }
            if (k == 0) continue;
            /* zero-valued in-mask pixels of THIS (road, tile) call per band: fct_misc.py:95-111 pads per call */
            uint64_t zcall[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int y = 0; y < win[3]; y++)
                for (int x = 0; x < win[2]; x++) {
                    if (!scratch[(size_t)y * win[2] + x]) continue;
                    size_t pix = ((size_t)t * H + (size_t)(win[1] + y)) * W + (size_t)(win[0] + x);
                    int v[8];
                    int allzero = 1;
                    for (int c = 0; c < C; c++) {
                        if (elem_bytes == 1) v[c] = ((const uint8_t *)tiles)[pix * C + c];
                        else {
                            unsigned s = ((const uint16_t *)tiles)[pix * C + c];
                            double o;
                            if (rescale_f32) {
                                float f = (float)s * (float)scale_k[c] + (float)scale_off[c];
                                if (f < 0.0f) f = 0.0f;
                                if (f > 255.0f) f = 255.0f;
                                o = (double)(f + 0.5f);
                            } else {
                                double f = (double)s * scale_k[c] + scale_off[c];
                                if (f < 0.0) f = 0.0;
                                if (f > 255.0) f = 255.0;
                                o = f + 0.5;
                            }
                            v[c] = (int)o;
                        }
                        if (v[c] != 0) allzero = 0;
                    }
                    if (joint) {
                        int cls = v[0] > 2 ? 0 : v[0];       /* unknown class codes count as "none" */
                        hr[cls * 256 + v[1]]++;
                    } else {
                        for (int c = 0; c < C; c++) {
                            hr[c * 256 + v[c]]++;
                            if (v[c] == 0) zcall[c]++;
                        }
                    }
                    if (allzero) n_allzero[r]++;
                }
            if (min_zero && !joint) {
                uint64_t m = zcall[0];
                for (int c = 1; c < C; c++)
                    if (zcall[c] < m) m = zcall[c];
                min_zero[r] += m;
            }
        }
    }
    free(psize);
    free(scratch);
    return rc;
}
