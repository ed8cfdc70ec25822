/*
 * roadsurf_b200 -- C ABI of the B200-native raster-vector overlay hot path of proj-roadsurf.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.  The
 * reference has no FFI (it is Python calling rasterio/GDAL, rasterstats, pandas and
 * geopandas); each entry point below names the reference call it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *  - every function returns an int status (RS_OK == 0, negative = error); nothing throws;
 *  - "_dev" entry points take DEVICE pointers, enqueue on `stream` (a cudaStream_t passed as
 *    void*, NULL = default stream) and return without synchronising; kernel-side failures
 *    (a rotated tile transform ...) are latched in the context and read with rs_ctx_sync_status();
 *  - "_host" entry points take HOST pointers, do their own host<->device copies and return
 *    after the results are in the host buffers;
 *  - the caller owns every input and output buffer; the library only owns its context.
 *
 * Geometry layout (all entry points): a "road" is a (Multi)Polygon flattened to rings,
 *   xy            double[n_verts][2]   ring vertices in tile CRS, rings stored as given
 *                                      (GeoJSON rings are closed: last == first)
 *   ring_off      int32[n_rings+1]     vertex offset of each ring
 *   road_ring_off int32[n_roads+1]     ring offset of each road (exterior + holes of every part)
 *   road_bbox     double[n_roads][4]   xmin, ymin, xmax, ymax over the road's vertices
 * Tiles: pixels [n_tiles][height][width][channels], pixel-interleaved; gt double[n_tiles][6]
 * = affine (a, b, c, d, e, f) with x = a*col + b*row + c, y = d*col + e*row + f (rasterio
 * order); only north-up transforms (b == d == 0) are on the hot path.
 * Pairs: road-major CSR -- road_pair_off int32[n_roads+1], pair_tile int32[n_pairs]: the
 * tiles each road is masked against (reference: gpd.sjoin result consumed by the double loop
 * scripts/statistical_analysis/statistical_analysis.py:170-171,180-193).
 */
#ifndef ROADSURF_B200_H
#define ROADSURF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_VERSION 100

enum rs_status {
    RS_OK = 0,
    RS_ERR_INVALID_ARG = -1,   /* NULL pointer, negative size, misaligned xy, bad enum            */
    RS_ERR_CUDA = -2,          /* a CUDA runtime call failed (rs_ctx_last_cuda_error for details)  */
    RS_ERR_CAPACITY = -3,      /* reserved (the bit-mask fill has no per-scanline crossing limit)  */
    RS_ERR_ROTATED = -4,       /* a tile transform has b != 0 or d != 0                           */
    RS_ERR_NO_DEVICE = -5,     /* no CUDA device / device is not sm_100                           */
    RS_ERR_UNSUPPORTED = -6,   /* a (road, raster) window wider than 2048 px, channels not in 1..4, dtype/channels combination */
    RS_ERR_NOT_PINNED = -7,    /* rs_zonal_stats_mapped_host: tiles->pixels is not page-locked    */
    RS_ERR_NO_NCCL = -8,       /* rs_comm_*: libnccl.so.2 could not be loaded                     */
    RS_ERR_NCCL = -9,          /* an NCCL call failed                                             */
    RS_ERR_CODEC = -10         /* rs_decode_segments_*: a compressed segment is corrupt or does not decode to its expected size */
};

enum rs_dtype { RS_U8 = 0, RS_U16 = 1 };

enum rs_hist_mode {
    RS_HIST_BANDS = 0,         /* hist[slot][band][256]: one histogram per band                   */
    RS_HIST_CLASS_SCORE = 1    /* channels == 2 (class, score): hist[slot][3][256], class 0/1/2   */
};

enum rs_window_mode {
    RS_WINDOW_CROP = 0,        /* rasterio.mask.mask(crop=True): per-pair integer window and its   */
                               /* window transform (fct_misc.py:77)                                */
    RS_WINDOW_FULL = 1,        /* rasterio.features.rasterize(shapes, out_shape, transform): the   */
                               /* tile transform itself (add_tile_mask.py:112-113)                 */
    RS_WINDOW_BOUNDLESS = 2    /* rasterstats.zonal_stats: the window of the geometry bounds, NOT   */
                               /* clipped to the raster (its origin may be negative); pixels off   */
                               /* the raster are nodata (fct_rasters.py:162-163)                   */
};

enum rs_nodata_mode {
    RS_NODATA_RAW = 0,         /* statistics over every in-mask pixel                              */
    RS_NODATA_NONE = 1,        /* tile nodata is None: pixels with all bands 0 dropped (fct_misc.py:117-119) */
    RS_NODATA_ZERO = 2,        /* tile nodata == 0: per band zeros dropped, short bands zero-padded per (road, tile) call
                                  (fct_misc.py:95-111); rs_finalize_stats_* then takes rs_zonal_params::min_zero as `n_allzero` */
    RS_NODATA_ZERO_MASKED = 3  /* rasterstats nodata=0: per band zeros masked (statistical_analysis.py:221) */
};

typedef struct rs_ctx rs_ctx;

typedef struct rs_roads {
    const double  *xy;
    const int32_t *ring_off;
    const int32_t *road_ring_off;
    const double  *road_bbox;
    int32_t n_roads, n_rings, n_verts;
} rs_roads;

typedef struct rs_tiles {
    const void   *pixels;
    const double *gt;
    int32_t n_tiles, height, width, channels;
    int32_t dtype;              /* enum rs_dtype */
} rs_tiles;

typedef struct rs_pairs {
    const int32_t *road_pair_off;
    const int32_t *pair_tile;
    int32_t n_pairs;
} rs_pairs;

typedef struct rs_zonal_params {
    int32_t hist_mode;          /* enum rs_hist_mode */
    int32_t window_mode;        /* enum rs_window_mode */
    int32_t rescale;            /* RS_U16 only: 0 = none (illegal), 1 = float64, 2 = float32 working precision */
    int32_t border_px;          /* ignore the outermost border_px pixels of every tile: the raster form of determine_class.clip_labels,
                                   labels clipped to the tile scaled by 0.99 (determine_class.py:62-95); 0 = whole tile */
    double  scale_k[4];         /* dst = clamp(src*k + off, 0, 255) + 0.5 truncated -- gdal.Translate    */
    double  scale_off[4];       /* scaleParams (scripts/preprocessing/tif2cog.py:260-270)                */
    const int32_t *road_slot;   /* optional int32[n_roads]: output row of each road (NULL = identity)    */
    uint32_t *min_zero;         /* optional output uint32[n_slots] (RS_HIST_BANDS only; device pointer for _dev, host pointer for
                                   rs_zonal_hist_host): sum over the road's (road, tile) pairs of min over bands of the pair's
                                   zero-valued in-mask pixels -- the per-call zero padding of get_pixel_values when the tile's
                                   nodata is 0 (fct_misc.py:95-111); rs_finalize_stats_* takes it in RS_NODATA_ZERO mode */
} rs_zonal_params;

/* statistics row produced by rs_finalize_stats_*: doubles, RS_NSTAT fixed columns then the
 * requested percentiles */
enum rs_stat_col { RS_STAT_COUNT = 0, RS_STAT_MIN, RS_STAT_MAX, RS_STAT_SUM, RS_STAT_SUMSQ,
                   RS_STAT_MEAN, RS_STAT_STD, RS_STAT_MEDIAN, RS_STAT_MARGIN, RS_NSTAT };

/* per-threshold metrics row produced by rs_vote_metrics_*: doubles */
enum rs_metric_col { RS_MET_P0 = 0, RS_MET_R0, RS_MET_F0, RS_MET_P1, RS_MET_R1, RS_MET_F1,
                     RS_MET_PW, RS_MET_RW, RS_MET_F1W, RS_MET_PB, RS_MET_RB, RS_MET_F1B, RS_NMETRIC };

enum rs_vote_rule { RS_VOTE_COUNT = 0, RS_VOTE_SCORE = 1 };
enum rs_cover { RS_COVER_ARTIFICIAL = 0, RS_COVER_NATURAL = 1, RS_COVER_UNDETERMINED = 2, RS_COVER_UNDETECTED = 3 };

int         rs_version(void);
const char *rs_status_string(int status);

/* one context per device: owns the work counters, the status word, the work-item scratch of the zonal kernel and the
 * staging buffers of the _host entry points.  A context is used by ONE host thread at a time.  _dev calls on different
 * streams through one context are ordered by the library (each waits for the previous user of the scratch), so they do
 * not overlap; use one context per stream for concurrency.  A _dev call may synchronise the device once when its scratch
 * has to grow (first call, or a larger pair list than any before). */
int rs_ctx_create(int device, rs_ctx **out);
int rs_ctx_destroy(rs_ctx *ctx);
/* synchronise `stream`, return and clear the latched kernel-side status */
int rs_ctx_sync_status(rs_ctx *ctx, void *stream);
int rs_ctx_last_cuda_error(rs_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py: gpu_launches) */
int64_t rs_ctx_launch_count(rs_ctx *ctx);

/*
 * Page-lock a caller-owned host buffer (cudaHostRegister, portable + mapped) so that rs_zonal_stats_mapped_host can read
 * it in place and the other _host entry points copy from it at full link speed; a buffer that is already page-locked is
 * left alone.  The caller unregisters it before freeing the memory.
 */
int rs_host_register(rs_ctx *ctx, void *ptr, size_t bytes);
int rs_host_unregister(rs_ctx *ctx, void *ptr);

/* road_bbox from xy (device pointers).  Host code normally has it from geometry.bounds. */
int rs_road_bbox_dev(rs_ctx *ctx, const rs_roads *roads, double *road_bbox_out, void *stream);

/*
 * Fused rasterize + zonal accumulation.  Replaces, for the whole pair list at once,
 *   scripts/functions/fct_misc.py:57-123 get_pixel_values  (rasterio.mask.mask :77 + np.extract :95)
 *   scripts/statistical_analysis/statistical_analysis.py:180-193 (the road x tile loop)
 *   rasterstats.zonal_stats rasterize+mask (statistical_analysis.py:221, fct_rasters.py:162)
 * hist      uint32[n_slots][HC][256], HC = channels (RS_HIST_BANDS) or 3 (RS_HIST_CLASS_SCORE)
 * n_allzero uint32[n_slots]: in-mask pixels whose bands are all 0
 * Every road's slot is written exactly once (zeros if the road has no pixels): outputs need
 * no clearing (road_slot, when given, must be injective).  Integer outputs are bit-exact and independent of
 * scheduling.  Limits: the window of a pair (bounds of the road clipped to the raster) at most 2048 pixels wide, any height,
 * any raster size; 64/128-bit pixel loads need width % 8 == 0 and a
 * 16-byte aligned pixel base (other shapes take a byte-wise path).
 */
int rs_zonal_hist_dev(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                      const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, void *stream);
int rs_zonal_hist_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                       const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero);

/*
 * rs_zonal_stats_host for tile sets larger than the device (or than one wants resident): the tiles stay in (pinned) host
 * memory and are streamed through two device buffers of tiles_per_chunk tiles; the copy of chunk k+1 overlaps the
 * kernel of chunk k, the per-road histograms accumulate on the device (integer atomics, exact) and are finalized once.
 * Same arguments and results as rs_zonal_stats_host.
 */
int rs_zonal_stats_stream_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                               const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                               int32_t n_pct, int32_t tiles_per_chunk, double *stats, uint32_t *hist, uint32_t *n_allzero);

/*
 * rs_zonal_stats_host without a copy of the tiles: tiles->pixels must be page-locked host memory (cudaHostAlloc /
 * cudaHostRegister, else RS_ERR_NOT_PINNED) and zonal_kernel reads it in place through the unified address space.
 * The kernel is span-driven, so only the 32-byte sectors under road pixels cross the host link instead of every tile
 * byte.  Same arguments and results as rs_zonal_stats_host.
 */
int rs_zonal_stats_mapped_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                               const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                               int32_t n_pct, double *stats, uint32_t *hist, uint32_t *n_allzero);

/*
 * Pixel masks.  Replaces rasterio.features.rasterize (scripts/sandbox/add_tile_mask.py:112-113,
 * window_mode FULL) and the shape mask of rasterio.mask.mask (fct_misc.py:77, window_mode CROP).
 * masks uint8[n_pairs][height][width] (pair order = pair_tile order), 1 = selected; the
 * buffer must be zeroed by the caller.  `tiles->pixels` is not read.
 */
int rs_rasterize_pairs_dev(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                           int window_mode, uint8_t *masks, void *stream);
int rs_rasterize_pairs_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                            int window_mode, uint8_t *masks);

/*
 * Statistics from merged histograms.  Replaces pandas groupby.agg(['min','max','median','mean',
 * 'count','std']) + margin (scripts/functions/fct_statistics.py:55-63; ddof = 1) and the
 * rasterstats reductions (ddof = 0).  stats double[n_roads][channels][RS_NSTAT + n_pct];
 * rows without pixels get count 0 and NaN elsewhere.  percentiles in [0,100], numpy 'linear'.
 */
int rs_finalize_stats_dev(rs_ctx *ctx, const uint32_t *hist, const uint32_t *n_allzero, int32_t n_roads,
                          int32_t channels, int32_t nodata_mode, int32_t ddof, const double *percentiles_host,
                          int32_t n_pct, double *stats, void *stream);
int rs_finalize_stats_host(rs_ctx *ctx, const uint32_t *hist, const uint32_t *n_allzero, int32_t n_roads,
                           int32_t channels, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                           int32_t n_pct, double *stats);

/*
 * One call from host buffers to the per-road statistics table: rs_zonal_hist + rs_finalize_stats with
 * the histograms staying on the device in between.  This is the batched form of the whole
 * statistical_analysis.py:179-246 block (pixel loop + groupby stats) and what bench.py times as e2e.
 * stats double[n_roads][channels][RS_NSTAT + n_pct] (required); hist / n_allzero optional (NULL = not
 * copied back).  prm->road_slot must be NULL.
 */
int rs_zonal_stats_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                        const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                        int32_t n_pct, double *stats, uint32_t *hist, uint32_t *n_allzero);

/*
 * Zonal statistics of polygons over ONE float32 raster: rasterstats.zonal_stats(labels, dem_array, affine=affine,
 * stats=['min','max','mean','median','std'], nodata=-9999)  (scripts/functions/fct_rasters.py:147-163, the DEM call).
 * raster float[height][width] (host), gt = its affine; pixels that are NaN or (use_nodata != 0) equal to nodata are masked, as are
 * the parts of a feature's window that lie off the raster (rasterstats' boundless read).  stats double[n_features][RS_NSTAT +
 * n_pct], the columns of rs_finalize_stats_*; features without a valid pixel get count 0 and NaN elsewhere.  count / min / max /
 * median are exact (the median of an even count is the float32 mean of the two middle values, as np.median gives); sum, mean and
 * std are binary64 reductions in a fixed order (rasterstats reduces in float32: ~1e-7 relative apart).  ddof 0 = rasterstats.
 * Any raster size; a feature's window may be at most 2048 pixels wide; at most 2^31 - 1 valid pixels in total.
 */
int rs_zonal_stats_f32_host(rs_ctx *ctx, const rs_roads *features, const float *raster, int32_t height, int32_t width,
                            const double *gt, int32_t use_nodata, double nodata, int32_t ddof, const double *percentiles,
                            int32_t n_pct, double *stats);

/*
 * Multi-GPU merge (one process per GPU, tiles sharded over the GPUs of a box).  The reference is single-process: a road's
 * pixels are concatenated over all its tiles before the groupby (scripts/statistical_analysis/statistical_analysis.py:187-193);
 * with the tiles spread over devices, the roads that touch tiles of several ranks get a row in a boundary table that has the
 * same layout on every rank, and one all-reduce(SUM) of the integer accumulators over NVLink completes those rows on every
 * rank (exact and order-independent; statistics are finalized afterwards, so medians and percentiles are exact).
 *   rs_comm_unique_id   rank 0 creates the 128-byte rendezvous id and hands it to the other ranks (any channel)
 *   rs_comm_init        every rank, same id: builds the NCCL communicator of this context (collective call)
 *   rs_allreduce_accumulators_dev   in-place sum over ranks of hist_rows[n_hist] (the boundary rows of the histogram table,
 *                       n_hist = rows * HC * 256), n_allzero_rows[n_rows] and, when not NULL, min_zero_rows[n_rows], as ONE
 *                       grouped NCCL launch on `stream`; device pointers; a context without communicator (single GPU) returns
 *                       RS_OK without doing anything
 * NCCL is bound at run time (libnccl.so.2; inside a PyTorch process the copy torch loaded): RS_ERR_NO_NCCL when absent.
 */
#define RS_COMM_ID_BYTES 128
int rs_comm_unique_id(void *id_out);
int rs_comm_init(rs_ctx *ctx, const void *id, int32_t world, int32_t rank);
int rs_comm_destroy(rs_ctx *ctx);
int rs_comm_world(rs_ctx *ctx);
int rs_allreduce_accumulators_dev(rs_ctx *ctx, uint32_t *hist_rows, int64_t n_hist, uint32_t *n_allzero_rows,
                                  uint32_t *min_zero_rows, int64_t n_rows, void *stream);

/*
 * Per-road vote, tags, confusion counts and F1 for a list of score cut-offs in one launch.
 * Replaces determine_class.determine_detected_class (scripts/road_segmentation/determine_class.py:122-190),
 * final_metrics.get_tag / get_metrics (scripts/road_segmentation/final_metrics.py:91-105, :22-89)
 * and the threshold sweep (:277-316) on raster accumulators.
 * joint_hist uint32[n_roads][3][256] (RS_HIST_CLASS_SCORE); gt_class int8[n_roads] (0 artificial,
 * 1 natural, anything else = road not in the ground truth, skipped); cutoffs int32[n_thr] = smallest
 * uint8 score kept.  Outputs: cover int8[n_thr][n_roads] (enum rs_cover), scores double[n_thr][n_roads][3]
 * (art_score, nat_score, diff_score), confusion int64[n_thr][2][4], metrics double[n_thr][RS_NMETRIC].
 */
int rs_vote_metrics_dev(rs_ctx *ctx, const uint32_t *joint_hist, const int8_t *gt_class, int32_t n_roads,
                        const int32_t *cutoffs_host, int32_t n_thr, int32_t rule, double min_area_frac,
                        int8_t *cover, double *scores, int64_t *confusion, double *metrics, void *stream);
int rs_vote_metrics_host(rs_ctx *ctx, const uint32_t *joint_hist, const int8_t *gt_class, int32_t n_roads,
                         const int32_t *cutoffs, int32_t n_thr, int32_t rule, double min_area_frac,
                         int8_t *cover, double *scores, int64_t *confusion, double *metrics);

/*
 * Ordered pixel extraction: the in-mask pixels of every pair, row-major inside the pair, pairs in
 * pair_tile order -- what fct_misc.get_pixel_values collects with np.extract (fct_misc.py:87-99) before its
 * nodata handling.  pair_off int64[n_pairs+1] (always written; pair_off[n_pairs] = total pixel count,
 * also returned in *n_total); values [total][channels] of the tile dtype, written only when values != NULL
 * and capacity_pixels >= total (call once with values == NULL to size the buffer).
 */
int rs_extract_pixels_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                           int window_mode, int64_t *pair_off, void *values, int64_t capacity_pixels, int64_t *n_total);

/*
 * Per-pixel derived columns of scripts/statistical_analysis/statistical_analysis.py:279-293 on the uint8 pixel table
 * values[n][channels] (bands 1..channels): the ratios band_a / band_b for a < b in the reference's loop order
 * (1/2, 1/3, 1/4, 2/3, 2/4, 3/4 = R/G, R/B, R/NIR, G/B, G/NIR, B/NIR), float64, rounded to 3 decimals the numpy way,
 * NaN -> 0 then inf -> 1; with 4 bands one more column VgNIR-BI = (band2 - band4) / (band2 + band4) rounded to 5
 * decimals (0/0 stays NaN, as in the reference).  out is column-major double[rs_band_ratio_columns(channels)][n]
 * (a DataFrame column each).  channels in 2..4.
 */
int rs_band_ratio_columns(int32_t channels);
int rs_band_ratios_dev(rs_ctx *ctx, const uint8_t *values, int64_t n, int32_t channels, double *out, void *stream);
int rs_band_ratios_host(rs_ctx *ctx, const uint8_t *values, int64_t n, int32_t channels, double *out);

/*
 * 'within' join of two polygon sets: scripts/road_segmentation/determine_class.py:41-62 get_roads_in_quarries =
 * gpd.sjoin(roads, buffered_quarries, predicate='within').  within uint8[a->n_roads][b->n_roads], 1 where polygon i of `a`
 * lies within polygon j of `b` (touching allowed): no vertex of a outside b (even-odd over b's rings), no proper edge
 * crossing, no vertex of b strictly inside a.  road_bbox may be NULL in both sets.  Binary64 orientation signs; can differ from
 * GEOS' robust predicates only for vertices within rounding distance of b's boundary.
 */
int rs_within_host(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, uint8_t *within);

/*
 * Overlay areas: scripts/road_segmentation/determine_class.py:107-118 get_weighted_scores =
 * gpd.overlay(ground_truth, predictions, how='intersection').area and ground_truth.area.  For every candidate pair
 * (pair_a[k], pair_b[k]) the area of polygon pair_a[k] of `a` intersected with polygon pair_b[k] of `b` (0 when they do not
 * overlap) -> area_pair[k]; the area of every polygon of `a` -> area_a (may be NULL).  Rings may have either orientation;
 * holes and parts follow the even-odd rule; the rings of one polygon must not cross each other (valid polygons, as GEOS
 * requires too).  Binary64 boundary integrals: agree with GEOS' noded overlay to rounding.
 */
int rs_overlay_area_host(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, const int32_t *pair_a, const int32_t *pair_b,
                         int32_t n_pairs, double *area_pair, double *area_a);

/*
 * Calibration bins of scripts/road_segmentation/final_metrics.py:541-571: for every group g (gt_type), value column k
 * and threshold t, counts[g][k][t] = { rows with sel[k][r] != 0 and lo[t] < values[k][r] <= hi[t],  those of them with
 * hit[k][r] != 0 }; the bin accuracy is their quotient where the first is non-zero.  values double[n_cols][n],
 * sel / hit int8[n_cols][n], group int32[n] in [0, n_groups) (other rows skipped), counts int64[n_groups][n_cols][n_thr][2].
 * The reference's bounds are lo = threshold - 0.5, hi = threshold (the 0.5 is the reference's, :557).
 */
int rs_bin_counts_host(rs_ctx *ctx, const double *values, const int8_t *sel, const int8_t *hit, const int32_t *group, int32_t n,
                       int32_t n_cols, int32_t n_groups, const double *lo, const double *hi, int32_t n_thr, int64_t *counts);

/*
 * 256-bin histogram per group of a uint8 column: the groupby of fct_statistics.get_df_stats_groupby
 * (fct_statistics.py:55) / the single group of get_df_stats_no_group (:89-94); finalize with
 * rs_finalize_stats_*.  group int32[n] in [0, n_groups) (other values are skipped); hist uint32[n_groups][256].
 */
int rs_group_hist_host(rs_ctx *ctx, const uint8_t *values, const int32_t *group, int64_t n, int32_t n_groups, uint32_t *hist);

/*
 * determine_class.determine_detected_class on the detection table (determine_class.py:133-179).  The rows
 * of road r are row_off[r] .. row_off[r+1] (the table sorted by road, original order kept inside a road);
 * cls int8 (0 artificial, 1 natural), score / weighted_score / area_pred_in_label float64; thresholds
 * float64[n_thr] (rows with score >= threshold vote).  cover int8[n_thr][n_roads] (enum rs_cover);
 * scores double[n_thr][n_roads][3] = artificial index, natural index (unrounded), diff_score.
 */
int rs_vote_table_host(rs_ctx *ctx, const int32_t *row_off, const int8_t *cls, const double *score, const double *weighted,
                       const double *area, int32_t n_roads, const double *thresholds, int32_t n_thr, int8_t *cover,
                       double *scores);

/*
 * final_metrics.get_tag + get_metrics (final_metrics.py:22-105) from cover codes and ground-truth classes:
 * cover int8[n_thr][n_roads], gt_class int8[n_roads] (0 / 1, others skipped); confusion int64[n_thr][2][4],
 * metrics double[n_thr][RS_NMETRIC].
 */
int rs_confusion_metrics_host(rs_ctx *ctx, const int8_t *cover, const int8_t *gt_class, int32_t n_roads, int32_t n_thr,
                              int64_t *confusion, double *metrics);

/*
 * GPU broad phase: the (road, tile) pair list of gpd.sjoin(tiles, roads) + drop_duplicates
 * (scripts/statistical_analysis/statistical_analysis.py:170-171) for tiles on a regular lattice, as road-major CSR
 * with each road's tiles sorted by index.  Predicate: closed bounding-box overlap of road_bbox[r] with
 * tile_ext[t] (xmin, ymin, xmax, ymax) -- a superset of 'intersects'; the extra pairs contribute no pixel.
 * Lattice cell (ix, iy) covers [x0 + ix*tile_w, +tile_w] x [y0 + iy*tile_h, +tile_h]; lut int32[ny][nx] gives the
 * tile index of a cell or -1.  road_pair_off int32[n_roads+1] is always written and *n_pairs returned; pair_tile
 * is written when it is non-NULL and capacity >= *n_pairs (call once with NULL to size it).
 */
typedef struct rs_lattice {
    double x0, y0, tile_w, tile_h;
    int32_t nx, ny;
    const int32_t *lut;
} rs_lattice;
int rs_pairs_bbox_host(rs_ctx *ctx, const double *road_bbox, int32_t n_roads, const double *tile_ext, int32_t n_tiles,
                       const rs_lattice *lattice, int32_t *road_pair_off, int32_t *pair_tile, int64_t capacity, int64_t *n_pairs);

/*
 * The same broad phase for tiles that are NOT on a lattice (any extents, any order): the tiles are binned into a uniform grid
 * on the device, every road looks up the cells under its bounding box.  Same predicate, outputs and two-call protocol as
 * rs_pairs_bbox_host.
 */
int rs_pairs_bbox_grid_host(rs_ctx *ctx, const double *road_bbox, int32_t n_roads, const double *tile_ext, int32_t n_tiles,
                            int32_t *road_pair_off, int32_t *pair_tile, int64_t capacity, int64_t *n_pairs);

/*
 * Exact reject of a candidate pair list: keep[p] = 1 iff the road polygon of pair p (all rings, even-odd) INTERSECTS the closed
 * rectangle tile_ext[pair_tile[p]] -- the predicate of gpd.sjoin(tiles, roads) (statistical_analysis.py:170-171; touching
 * counts).  An edge touches the rectangle, or the rectangle lies inside the polygon.  Binary64 orientation signs: can differ
 * from GEOS only for contacts within rounding distance.  (Pairs the bounding-box phase keeps in excess contribute no pixel, so
 * this filter changes no statistic; it makes the pair table the reference's and saves their work items.)
 */
int rs_pairs_intersect_host(rs_ctx *ctx, const rs_roads *roads, const double *tile_ext, int32_t n_tiles,
                            const int32_t *road_pair_off, const int32_t *pair_tile, int32_t n_pairs, uint8_t *keep);

/*
 * clip_labels (scripts/road_segmentation/determine_class.py:62-95): the rings of label pair_label[p] clipped to the closed
 * rectangle rect[p] = (xmin, ymin, xmax, ymax) -- the tile scaled by 0.99 about its centre -- for every pair of the
 * labels x tiles 'intersects' join (rs_pairs_bbox_* + rs_pairs_intersect_host).  A pair-ring is ring k of the label of pair p,
 * numbered pair_ring_off[p] + k (pair_ring_off int64[n_pairs + 1], the prefix sum of the labels' ring counts).  Two calls:
 *   xy_out == NULL   ring_count[q] = vertices of the clipped pair-ring q, closed (0 when it misses the rectangle);
 *   xy_out != NULL   the clipped rings written at xy_out[ring_vert_off[q]] (ring_vert_off int64[n_pair_rings], from ring_count).
 * Re-entrant Sutherland-Hodgman: parts of a concave ring that leave and re-enter the rectangle stay joined by zero-width runs
 * along its edge (no area under the even-odd rule); GEOS returns them as separate parts -- same point set, same areas.
 */
int rs_clip_rings_host(rs_ctx *ctx, const rs_roads *labels, const int32_t *pair_label, const double *rect, int32_t n_pairs,
                       const int64_t *pair_ring_off, int32_t *ring_count, const int64_t *ring_vert_off, double *xy_out);

/*
 * Two-sample Kolmogorov-Smirnov statistic of every road's pixel values on one band against a reference distribution,
 * from histograms: scipy.stats.kstest(road_values, general_values) of statistical_analysis.py:441-451 (the pixels of the
 * road against all pixels of its road type).  hist uint32[n_roads][256] (one band), ref_hist uint64[n_refs][256],
 * ref_of_road int32[n_roads] (NULL = reference 0; negative = skip, D = NaN).  D double[n_roads], n double[n_roads]
 * (sample size of the road, for the p-value: scipy evaluates kstwo.sf(D, round(m n / (m + n))) for large samples).
 */
int rs_ks_hist_host(rs_ctx *ctx, const uint32_t *hist, const int32_t *ref_of_road, const uint64_t *ref_hist, int32_t n_roads,
                    int32_t n_refs, double *D, double *n);

/*
 * 16 -> 8 bit rescale as a materialising pass: gdal.Translate(outputType=GDT_Byte, scaleParams=[[smin, smax, 0, 255]...])
 * (scripts/preprocessing/tif2cog.py:260-270) plus the band selection of the tile URL (config/config_stats.yaml:39).
 * dst[px][c] = byte(clamp(src[px][bidx[c]] * k[c] + off[c], 0, 255) + 0.5); k / off / bidx (NULL = identity) are host
 * arrays of c_out entries; f32 != 0 evaluates in float32 (GDAL's working precision is not pinned).
 * (rs_zonal_hist with RS_U16 applies the same formula on the fly without materialising the 8-bit tiles.)
 */
int rs_rescale_u16_dev(rs_ctx *ctx, const uint16_t *src, int64_t n_pixels, int32_t c_in, int32_t c_out, const int32_t *bidx,
                       const double *k, const double *off, int32_t f32, uint8_t *dst, void *stream);
int rs_rescale_u16_host(rs_ctx *ctx, const uint16_t *src, int64_t n_pixels, int32_t c_in, int32_t c_out, const int32_t *bidx,
                        const double *k, const double *off, int32_t f32, uint8_t *dst);

/*
 * Tile ingest: decompressed TIFF segments of n_tiles equally shaped tiles -> the pixel-interleaved batch rs_zonal_* reads
 * (what rasterio's src.read() + np.moveaxis deliver in fct_misc.py:76-77, and the band selection bidx=2,3,4,1 of
 * config/config_stats.yaml:39).  raw holds, per tile, height*width*c_in samples of sample_bytes (1 | 2) bytes:
 * planar 1 = [H][W][c_in] (PlanarConfiguration chunky), 2 = [c_in][H][W] (band-sequential); predictor 2 = TIFF horizontal
 * differencing (undone per row and sample, modulo the sample width), 1 = none; big_endian: byte order of 16-bit samples.
 * out[n_tiles][H][W][c_out], band c = input band bidx[c] (NULL = identity); rescale 0 keeps the sample width, 1 / 2 apply
 * dst = clamp_round(src * k[c] + off[c]) in float64 / float32 to uint8 (tif2cog.py:260-270, as rs_rescale_u16_*).
 */
int rs_assemble_tiles_dev(rs_ctx *ctx, const uint8_t *raw, int32_t n_tiles, int32_t height, int32_t width, int32_t c_in,
                          int32_t planar, int32_t predictor, int32_t sample_bytes, int32_t big_endian, int32_t c_out,
                          const int32_t *bidx, int32_t rescale, const double *k, const double *off, void *out, void *stream);
int rs_assemble_tiles_host(rs_ctx *ctx, const uint8_t *raw, int32_t n_tiles, int32_t height, int32_t width, int32_t c_in,
                           int32_t planar, int32_t predictor, int32_t sample_bytes, int32_t big_endian, int32_t c_out,
                           const int32_t *bidx, int32_t rescale, const double *k, const double *off, void *out);

/*
 * Tile ingest, decompression on the device: the compressed segments (strips / internal tiles) of a batch of TIFFs, concatenated
 * in comp[comp_off[s] .. comp_off[s + 1]), are decoded one decoder per segment (DEFLATE: table-driven, the 32 decoders of a warp
 * stepped together; RS_INFLATE = lut | warp | bits selects the decoder) into raw[raw_off[s] .. raw_off[s + 1]); a segment
 * must decode to exactly that many bytes (libtiff: rows * row bytes), otherwise RS_ERR_CODEC (latched for _dev, returned by
 * _host).  codec = the TIFF Compression tag: 1 none, 5 LZW (MSB-first, early change), 8 / 32946 zlib-wrapped DEFLATE.  The
 * compressed bytes are what crosses the host link; `raw` then feeds rs_assemble_tiles_* (predictor, byte order, bands, rescale).
 * Replaces the libtiff decode under rasterio's src.read() (scripts/functions/fct_misc.py:76-77).  comp_off / raw_off int64[n + 1].
 */
int rs_decode_segments_dev(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec,
                           uint8_t *raw, const int64_t *raw_off, void *stream);
int rs_decode_segments_host(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec,
                            uint8_t *raw, const int64_t *raw_off);

/*
 * The whole ingest of a batch of equally shaped TIFF tiles in one call: compressed segments H2D -> rs_decode_segments ->
 * rs_assemble_tiles -> out[n_tiles][H][W][c_out] D2H.  raw_off must tile the sample buffer of rs_assemble_tiles exactly
 * (n_tiles * H * W * c_in * sample_bytes bytes: per tile [H][W][c_in] or [c_in][H][W], strips in row order).  With
 * keep_on_device != 0 the assembled batch also stays in the context's staging buffer (*device_out, valid until the next _host
 * call on this context) so that a caller can run rs_zonal_hist_dev on it without a second upload; out may then be NULL.
 */
int rs_ingest_tiles_host(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec,
                         const int64_t *raw_off, int32_t n_tiles, int32_t height, int32_t width, int32_t c_in, int32_t planar,
                         int32_t predictor, int32_t sample_bytes, int32_t big_endian, int32_t c_out, const int32_t *bidx,
                         int32_t rescale, const double *k, const double *off, void *out, int32_t keep_on_device, void **device_out);

/*
 * rs_zonal_stats_host for uint8 tiles that are still COMPRESSED (the strips of GeoTIFF files as they are on disk): the compressed
 * bytes are uploaded, decoded and assembled on the device (rs_decode_segments + rs_assemble_tiles) and the statistics computed
 * from there -- the reference's rasterio.open(tile).read() + mask + statistics for a whole batch without the decoded tiles ever
 * crossing the host link.  tiles->pixels is ignored; raw_off must tile n_tiles * H * W * channels bytes.
 */
int rs_zonal_stats_compressed_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                                   const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                                   int32_t n_pct, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec,
                                   const int64_t *raw_off, int32_t planar, int32_t predictor, int32_t big_endian, double *stats);

/*
 * Deterministic synthetic tiles (bench / tests only; the reference ships no imagery,
 * data/readme.md:20-21).  value = f(seed, tile_key[t], pixel, band), see DESIGN.md.
 * kind 0: iid uniform; 1: low-entropy "asphalt"; 2: class/score planes (channels == 2).
 */
int rs_synth_tiles_dev(rs_ctx *ctx, void *pixels, const int64_t *tile_key, int32_t n_tiles, int32_t height,
                       int32_t width, int32_t channels, int32_t dtype, int32_t kind, uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ROADSURF_B200_H */
