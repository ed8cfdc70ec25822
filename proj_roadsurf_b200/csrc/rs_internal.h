// Internal declarations shared by the translation units of libroadsurf_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "roadsurf_b200.h"

namespace rs {

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void  *p = nullptr;
    size_t cap = 0;
};
struct HostCopier;                // page-locked slot ring + copy threads for pageable sources (rs_hostcopy.cu)

}  // namespace rs

struct rs_ctx {
    int device = 0;
    int sm_count = 0;
    int last_cuda_error = 0;
    int64_t launches = 0;
    int *d_status = nullptr;      // latched kernel-side status (min over failures)
    int *d_counters = nullptr;    // work counters / list sizes, RS_NCOUNTERS ints
    int *h_status_pinned = nullptr;
    cudaStream_t host_stream = nullptr;   // stream of the _host entry points
    cudaStream_t copy_stream = nullptr;   // H2D stream of the streaming entry point (created on first use)
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
    rs::DevBuf stage[16];                 // grow-only device staging of the _host entry points
    rs::DevBuf items;                     // grow-only work-item list of the zonal kernel
    rs::DevBuf pgeom;                     // grow-only per-pair geometry records of the zonal kernel
    rs::DevBuf pair_zero;                 // grow-only per-pair zero counts (min_zero of row-split pairs on tall tiles)
    rs::DevBuf pool, heads, ov_items;     // grow-only entry pool / per-item segment heads / overflow items of the two-kernel form
    rs::DevBuf wide_cnt, wide_off, wide_bounds, wide_flags, wide_pair_road, wide_tmp;   // per-launch tables of the wide-window kernel (rs_wide.cu)
    cudaEvent_t ev_scratch = nullptr;     // recorded after every launch that uses the scratch above
    cudaStream_t scratch_stream = nullptr;
    bool scratch_used = false;
    rs::DevBuf lzw_scratch;               // LZW string tables of the resident decoders (rs_codec.cu)
    rs::DevBuf lut_dev;                   // 16 -> 8 bit rescale thresholds of the last scale parameters (rs_zonal.cu, PxU16x4Lut)
    bool lut_valid = false, lut_ok = false, lut_f32_same = false;
    int lut_f32 = 0;
    double lut_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t lut_k[4] = {0, 0, 0, 0};
    long long lut_b[4] = {0, 0, 0, 0};
    rs::HostCopier *copier = nullptr;     // created by the first large copy out of pageable memory
    void *comm = nullptr;                 // ncclComm_t of rs_comm_init (rs_comm.cu)
    int comm_world = 0, comm_rank = 0;
};

namespace rs {

enum { RS_NCOUNTERS = 16 };

// Make `b` hold at least `bytes` bytes of device memory.
int ensure(rs_ctx *ctx, DevBuf &b, size_t bytes);

#define RS_CUDA_OK(ctx, call)                                  \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) {                              \
            (ctx)->last_cuda_error = (int)e__;                 \
            return RS_ERR_CUDA;                                \
        }                                                      \
    } while (0)

// host -> device copy queued on `st`; large pageable sources are staged by several host threads through page-locked slots
int copy_h2d(rs_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st);
void copier_destroy(rs_ctx *ctx);

// launches (defined in the .cu files)
int launch_zonal(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                 const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks,
                 int window_mode, cudaStream_t st);
int launch_zonal_chunk(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                       const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, uint8_t *masks, int window_mode,
                       int tile_lo, int tile_hi, int accumulate, cudaStream_t st);
int launch_zonal_f32(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode, double nodata,
                     int has_nodata, uint32_t *count, uint32_t *cursor, const unsigned long long *offset, float *values, int write,
                     cudaStream_t st);
int launch_zonal_extract(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode,
                         uint32_t *count, const unsigned long long *offset, uint8_t *values, int bpp, int write, cudaStream_t st);
int launch_fstats_offsets(rs_ctx *ctx, const uint32_t *counts, int n, unsigned long long *offsets, cudaStream_t st);
int launch_fstats_sort(rs_ctx *ctx, const float *values, float *sorted, long long total, int n, const unsigned long long *offsets,
                       cudaStream_t st);
int launch_fstats(rs_ctx *ctx, const float *sorted, const unsigned long long *offsets, int n, int ddof, const double *pct_dev, int n_pct,
                  double *stats, cudaStream_t st);
int launch_road_bbox(rs_ctx *ctx, const rs_roads *roads, double *out, cudaStream_t st);
// wide-window form (rs_wide.cu): tiles of 512 .. 2048 pixels under wide polygons with long edge lists
bool wide_eligible(const rs_tiles *tiles, const rs_zonal_params *prm, bool resident);
int launch_zonal_wide(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, const rs_zonal_params *prm,
                      uint32_t *hist, uint32_t *n_allzero, int window_mode, int tile_lo, int tile_hi, int accumulate, cudaStream_t st);
int launch_finalize(rs_ctx *ctx, const uint32_t *hist, const uint32_t *n_allzero, int n_roads, int channels,
                    int nodata_mode, int ddof, const double *pct_host, int n_pct, double *stats, cudaStream_t st);
int launch_vote(rs_ctx *ctx, const uint32_t *joint_hist, const int8_t *gt_class, int n_roads,
                const int32_t *cutoffs_host, int n_thr, int rule, double min_area_frac, int8_t *cover,
                double *scores, int64_t *confusion, double *metrics, cudaStream_t st);
int launch_metrics(rs_ctx *ctx, const int64_t *confusion, int n_thr, double *metrics, cudaStream_t st);
int launch_group_hist(rs_ctx *ctx, const uint8_t *values, const int *group, long long n, int n_groups, uint32_t *hist, cudaStream_t st);
int launch_vote_table(rs_ctx *ctx, const int *row_off, const int8_t *cls, const double *score, const double *weighted,
                      const double *area, int n_roads, const double *thr_dev, int n_thr, int8_t *cover, double *scores, cudaStream_t st);
int launch_confusion(rs_ctx *ctx, const int8_t *cover, const int8_t *gt, int n_roads, int n_thr, int64_t *confusion, double *metrics,
                     cudaStream_t st);
int launch_band_ratios(rs_ctx *ctx, const uint8_t *values, long long n, int channels, double *out, cudaStream_t st);
int launch_bin_counts(rs_ctx *ctx, const double *values, const int8_t *sel, const int8_t *hit, const int *group, int n, int n_cols,
                      int n_groups, const double *lo, const double *hi, int n_thr, int64_t *counts, cudaStream_t st);
int launch_overlay_area(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, const int *ring_poly_a, const int *ring_poly_b,
                        int8_t *sign_a, int8_t *sign_b, const int *pair_a, const int *pair_b, int n_pairs, double *area_pair,
                        double *area_a, cudaStream_t st);
int launch_intersects(rs_ctx *ctx, const rs_roads *roads, const double *tile_ext_dev, const int *road_pair_off_dev,
                      const int *pair_tile_dev, int n_pairs, uint8_t *keep_dev, cudaStream_t st);
int launch_pairs_grid(rs_ctx *ctx, const double *bbox_dev, int n_roads, const double *ext_dev, int n_tiles, double X0, double Y0,
                      double cw, double ch, int nx, int ny, int *road_pair_off_dev, int *pair_tile_dev, long long capacity, int phase,
                      cudaStream_t st);
int launch_clip_rings(rs_ctx *ctx, const rs_roads *labels, const int *pair_label, const double *rect, const long long *pair_ring_off,
                      int n_pairs, long long n_pair_rings, int *ring_count, const long long *ring_vert_off, double *xy_out,
                      cudaStream_t st);
int launch_decode_segments(rs_ctx *ctx, const uint8_t *comp, const long long *comp_off, int n_seg, int codec, uint8_t *raw,
                           const long long *raw_off, cudaStream_t st);
int launch_within(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, uint8_t *out, cudaStream_t st);
int launch_pairs_bbox(rs_ctx *ctx, const double *bbox_dev, int n_roads, const double *ext_dev, const rs_lattice *lat, const int *lut_dev,
                      int *road_pair_off_dev, int *pair_tile_dev, long long capacity, int phase, cudaStream_t st);
int launch_rescale(rs_ctx *ctx, const uint16_t *src, long long n_px, int c_in, int c_out, const int32_t *bidx_host, const double *k_host,
                   const double *off_host, int f32, uint8_t *dst, cudaStream_t st);
int launch_assemble(rs_ctx *ctx, const uint8_t *raw, int n_tiles, int H, int W, int c_in, int planar, int predictor, int bytes,
                    int big_endian, int c_out, const int32_t *bidx_host, int rescale, const double *k_host, const double *off_host,
                    void *out, cudaStream_t st);
int launch_ks(rs_ctx *ctx, const uint32_t *hist, const int *ref_of_road, const unsigned long long *ref_hist, int n_roads, int stride,
              double *d_out, double *n_out, cudaStream_t st);
int launch_synth(rs_ctx *ctx, void *pixels, const int64_t *tile_key, int n_tiles, int H, int W, int C, int dtype,
                 int kind, uint64_t seed, cudaStream_t st);

}  // namespace rs
