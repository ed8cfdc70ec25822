// C ABI of libroadsurf_b200.so (include/roadsurf_b200.h): context, argument checks, and the
// _host entry points (host buffers in, host buffers out, copies inside).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "rs_internal.h"

namespace rs {

int ensure(rs_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return RS_OK;
    if (b.p) RS_CUDA_OK(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    RS_CUDA_OK(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return RS_OK;
}

static int bind(rs_ctx *ctx)
{
    if (!ctx) return RS_ERR_INVALID_ARG;
    RS_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return RS_OK;
}

static int up(rs_ctx *ctx, DevBuf &b, const void *src, size_t bytes)
{
    int rc = ensure(ctx, b, bytes);
    if (rc) return rc;
    return copy_h2d(ctx, b.p, src, bytes, ctx->host_stream);
}

static size_t elem_bytes(int dtype) { return dtype == RS_U16 ? 2 : 1; }

// Copies roads / tiles / pairs to the device staging buffers and returns device-side descriptors.
static int stage_inputs(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, bool need_pixels,
                        rs_roads &dr, rs_tiles &dt, rs_pairs &dp)
{
    if (!roads || !tiles || !pairs) return RS_ERR_INVALID_ARG;
    if (roads->n_roads < 0 || roads->n_rings < 0 || roads->n_verts < 0 || tiles->n_tiles < 0 || pairs->n_pairs < 0)
        return RS_ERR_INVALID_ARG;
    if (tiles->channels < 1 || tiles->channels > 4) return RS_ERR_UNSUPPORTED;
    if (tiles->dtype != RS_U8 && tiles->dtype != RS_U16) return RS_ERR_INVALID_ARG;
    if (roads->n_roads > 0 && (!roads->xy || !roads->ring_off || !roads->road_ring_off || !pairs->road_pair_off))
        return RS_ERR_INVALID_ARG;
    int rc;
    dr = *roads;
    dt = *tiles;
    dp = *pairs;
    if ((rc = up(ctx, ctx->stage[0], roads->xy, sizeof(double) * 2 * (size_t)roads->n_verts))) return rc;
    if ((rc = up(ctx, ctx->stage[1], roads->ring_off, sizeof(int32_t) * ((size_t)roads->n_rings + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[2], roads->road_ring_off, sizeof(int32_t) * ((size_t)roads->n_roads + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[3], pairs->road_pair_off, sizeof(int32_t) * ((size_t)roads->n_roads + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[4], pairs->pair_tile, sizeof(int32_t) * (size_t)pairs->n_pairs))) return rc;
    if ((rc = up(ctx, ctx->stage[5], tiles->gt, sizeof(double) * 6 * (size_t)tiles->n_tiles))) return rc;
    dr.xy = (const double *)ctx->stage[0].p;
    dr.ring_off = (const int32_t *)ctx->stage[1].p;
    dr.road_ring_off = (const int32_t *)ctx->stage[2].p;
    dp.road_pair_off = (const int32_t *)ctx->stage[3].p;
    dp.pair_tile = (const int32_t *)ctx->stage[4].p;
    dt.gt = (const double *)ctx->stage[5].p;
    if ((rc = ensure(ctx, ctx->stage[6], sizeof(double) * 4 * (size_t)roads->n_roads))) return rc;
    if (roads->road_bbox) {
        if ((rc = up(ctx, ctx->stage[6], roads->road_bbox, sizeof(double) * 4 * (size_t)roads->n_roads))) return rc;
    } else if ((rc = launch_road_bbox(ctx, &dr, (double *)ctx->stage[6].p, ctx->host_stream)))
        return rc;
    dr.road_bbox = (const double *)ctx->stage[6].p;
    if (need_pixels) {
        if (tiles->n_tiles > 0 && !tiles->pixels) return RS_ERR_INVALID_ARG;
        const size_t nb = (size_t)tiles->n_tiles * tiles->height * tiles->width * tiles->channels * elem_bytes(tiles->dtype);
        if ((rc = up(ctx, ctx->stage[7], tiles->pixels, nb))) return rc;
        dt.pixels = ctx->stage[7].p;
    } else
        dt.pixels = nullptr;
    return RS_OK;
}

static int finish(rs_ctx *ctx)
{
    RS_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_status_pinned, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->host_stream));
    RS_CUDA_OK(ctx, cudaStreamSynchronize(ctx->host_stream));
    const int st = *ctx->h_status_pinned;
    if (st != 0) RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->host_stream));
    return st;
}

}  // namespace rs

using namespace rs;

extern "C" {

int rs_version(void) { return RS_VERSION; }

const char *rs_status_string(int status)
{
    switch (status) {
        case RS_OK: return "ok";
        case RS_ERR_INVALID_ARG: return "invalid argument";
        case RS_ERR_CUDA: return "CUDA runtime error";
        case RS_ERR_CAPACITY: return "scanline crossing capacity exceeded";
        case RS_ERR_ROTATED: return "rotated tile transform is not supported on this path";
        case RS_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        case RS_ERR_UNSUPPORTED: return "unsupported tile shape or dtype";
        case RS_ERR_NOT_PINNED: return "the tile buffer is not page-locked host memory";
        case RS_ERR_NO_NCCL: return "libnccl.so.2 could not be loaded";
        case RS_ERR_NCCL: return "an NCCL call failed";
        case RS_ERR_CODEC: return "a compressed tile segment is corrupt or has an unexpected size";
        default: return "unknown status";
    }
}

int rs_ctx_create(int device, rs_ctx **out)
{
    if (!out) return RS_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RS_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RS_ERR_NO_DEVICE;
    if (prop.major != 10) return RS_ERR_NO_DEVICE;       // the kernels are built for sm_100a only
    rs_ctx *ctx = new (std::nothrow) rs_ctx();
    if (!ctx) return RS_ERR_INVALID_ARG;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_status, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_counters, sizeof(int) * RS_NCOUNTERS);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_counters, 0, sizeof(int) * RS_NCOUNTERS);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_status_pinned, sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_scratch, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        rs_ctx_destroy(ctx);
        return RS_ERR_CUDA;
    }
    *out = ctx;
    return RS_OK;
}

int rs_ctx_destroy(rs_ctx *ctx)
{
    if (!ctx) return RS_OK;
    cudaSetDevice(ctx->device);
    rs_comm_destroy(ctx);
    copier_destroy(ctx);
    for (auto &b : ctx->stage)
        if (b.p) cudaFree(b.p);
    if (ctx->items.p) cudaFree(ctx->items.p);
    if (ctx->pgeom.p) cudaFree(ctx->pgeom.p);
    if (ctx->pair_zero.p) cudaFree(ctx->pair_zero.p);
    for (rs::DevBuf *b : {&ctx->wide_cnt, &ctx->wide_off, &ctx->wide_bounds, &ctx->wide_flags, &ctx->wide_pair_road, &ctx->wide_tmp})
        if (b->p) cudaFree(b->p);
    if (ctx->lut_dev.p) cudaFree(ctx->lut_dev.p);
    if (ctx->lzw_scratch.p) cudaFree(ctx->lzw_scratch.p);
    if (ctx->pool.p) cudaFree(ctx->pool.p);
    if (ctx->heads.p) cudaFree(ctx->heads.p);
    if (ctx->ov_items.p) cudaFree(ctx->ov_items.p);
    if (ctx->ev_scratch) cudaEventDestroy(ctx->ev_scratch);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_status_pinned) cudaFreeHost(ctx->h_status_pinned);
    if (ctx->host_stream) cudaStreamDestroy(ctx->host_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_used[i]) cudaEventDestroy(ctx->ev_used[i]);
    }
    delete ctx;
    return RS_OK;
}

int rs_ctx_sync_status(rs_ctx *ctx, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_status_pinned, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    RS_CUDA_OK(ctx, cudaStreamSynchronize(st));
    const int s = *ctx->h_status_pinned;
    if (s != 0) {
        RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(int), st));
        RS_CUDA_OK(ctx, cudaStreamSynchronize(st));
    }
    return s;
}

int rs_ctx_last_cuda_error(rs_ctx *ctx) { return ctx ? ctx->last_cuda_error : 0; }
int64_t rs_ctx_launch_count(rs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int rs_host_register(rs_ctx *ctx, void *ptr, size_t bytes)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!ptr || bytes == 0) return RS_ERR_INVALID_ARG;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost) return RS_OK;    // already page-locked
    cudaGetLastError();
    RS_CUDA_OK(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return RS_OK;
}

int rs_host_unregister(rs_ctx *ctx, void *ptr)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!ptr) return RS_ERR_INVALID_ARG;
    RS_CUDA_OK(ctx, cudaHostUnregister(ptr));
    return RS_OK;
}

int rs_road_bbox_dev(rs_ctx *ctx, const rs_roads *roads, double *road_bbox_out, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!roads || roads->n_roads < 0) return RS_ERR_INVALID_ARG;
    if (roads->n_roads > 0 && (!roads->xy || !roads->ring_off || !roads->road_ring_off || !road_bbox_out))
        return RS_ERR_INVALID_ARG;
    return launch_road_bbox(ctx, roads, road_bbox_out, (cudaStream_t)stream);
}

int rs_zonal_hist_dev(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                      const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!prm) return RS_ERR_INVALID_ARG;
    return launch_zonal(ctx, roads, tiles, pairs, prm, hist, n_allzero, nullptr, prm->window_mode, (cudaStream_t)stream);
}

int rs_rasterize_pairs_dev(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                           int window_mode, uint8_t *masks, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!masks && pairs && pairs->n_pairs > 0) return RS_ERR_INVALID_ARG;
    if (pairs && pairs->n_pairs == 0) return RS_OK;
    return launch_zonal(ctx, roads, tiles, pairs, nullptr, nullptr, nullptr, masks, window_mode, (cudaStream_t)stream);
}

int rs_zonal_hist_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                       const rs_zonal_params *prm, uint32_t *hist, uint32_t *n_allzero)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!prm || !hist || !n_allzero) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, true, dr, dt, dp))) return rc;
    rs_zonal_params p = *prm;
    int n_slots = prm->road_slot ? 0 : roads->n_roads;      /* rows of the output buffers */
    if (prm->road_slot) {
        for (int i = 0; i < roads->n_roads; i++) {
            if (prm->road_slot[i] < 0) return RS_ERR_INVALID_ARG;
            n_slots = prm->road_slot[i] + 1 > n_slots ? prm->road_slot[i] + 1 : n_slots;
        }
        if ((rc = up(ctx, ctx->stage[8], prm->road_slot, sizeof(int32_t) * (size_t)roads->n_roads))) return rc;
        p.road_slot = (const int32_t *)ctx->stage[8].p;
    }
    const int HC = prm->hist_mode == RS_HIST_CLASS_SCORE ? 3 : tiles->channels;
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)HC * n_slots, zb = sizeof(uint32_t) * (size_t)n_slots;
    if ((rc = ensure(ctx, ctx->stage[9], hb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], zb))) return rc;
    if (prm->min_zero) {
        if ((rc = ensure(ctx, ctx->stage[12], zb))) return rc;
        p.min_zero = (uint32_t *)ctx->stage[12].p;
    }
    if (prm->road_slot) {      // slots no road maps to stay zero
        RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[9].p, 0, hb, ctx->host_stream));
        RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[10].p, 0, zb, ctx->host_stream));
        if (prm->min_zero) RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[12].p, 0, zb, ctx->host_stream));
    }
    rc = launch_zonal(ctx, &dr, &dt, &dp, &p, (uint32_t *)ctx->stage[9].p, (uint32_t *)ctx->stage[10].p, nullptr,
                      prm->window_mode, ctx->host_stream);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(hist, ctx->stage[9].p, hb, cudaMemcpyDeviceToHost, ctx->host_stream));
    RS_CUDA_OK(ctx, cudaMemcpyAsync(n_allzero, ctx->stage[10].p, zb, cudaMemcpyDeviceToHost, ctx->host_stream));
    if (prm->min_zero) RS_CUDA_OK(ctx, cudaMemcpyAsync(prm->min_zero, ctx->stage[12].p, zb, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_zonal_stats_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                        const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                        int32_t n_pct, double *stats, uint32_t *hist, uint32_t *n_allzero)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!prm || !stats || prm->road_slot || n_pct < 0 || n_pct > 16) return RS_ERR_INVALID_ARG;
    if (prm->hist_mode != RS_HIST_BANDS) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, true, dr, dt, dp))) return rc;
    const int R = roads->n_roads, C = tiles->channels;
    if (R == 0) return RS_OK;
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)C * R, zb = sizeof(uint32_t) * (size_t)R;
    const size_t sb = sizeof(double) * (size_t)(RS_NSTAT + n_pct) * C * R;
    if ((rc = ensure(ctx, ctx->stage[9], hb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], zb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sb))) return rc;
    cudaStream_t st = ctx->host_stream;
    rs_zonal_params p = *prm;
    const uint32_t *aux = (const uint32_t *)ctx->stage[10].p;
    p.min_zero = nullptr;
    if (nodata_mode == RS_NODATA_ZERO) {        // the per-call zero padding of get_pixel_values needs the per-pair minima
        if ((rc = ensure(ctx, ctx->stage[12], zb))) return rc;
        aux = p.min_zero = (uint32_t *)ctx->stage[12].p;
    }
    rc = launch_zonal(ctx, &dr, &dt, &dp, &p, (uint32_t *)ctx->stage[9].p, (uint32_t *)ctx->stage[10].p, nullptr,
                      prm->window_mode, st);
    if (rc) return rc;
    rc = launch_finalize(ctx, (const uint32_t *)ctx->stage[9].p, aux, R, C, nodata_mode, ddof,
                         percentiles, n_pct, (double *)ctx->stage[11].p, st);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(stats, ctx->stage[11].p, sb, cudaMemcpyDeviceToHost, st));
    if (hist) RS_CUDA_OK(ctx, cudaMemcpyAsync(hist, ctx->stage[9].p, hb, cudaMemcpyDeviceToHost, st));
    if (n_allzero) RS_CUDA_OK(ctx, cudaMemcpyAsync(n_allzero, ctx->stage[10].p, zb, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_zonal_stats_stream_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                               const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                               int32_t n_pct, int32_t tiles_per_chunk, double *stats, uint32_t *hist, uint32_t *n_allzero)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!prm || !stats || prm->road_slot || n_pct < 0 || n_pct > 16 || tiles_per_chunk < 1) return RS_ERR_INVALID_ARG;
    if (prm->hist_mode != RS_HIST_BANDS) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, false, dr, dt, dp))) return rc;     // geometry + pairs + transforms, no pixels
    if (tiles->n_tiles > 0 && !tiles->pixels) return RS_ERR_INVALID_ARG;
    const int R = roads->n_roads, C = tiles->channels;
    if (R == 0) return RS_OK;
    if (!ctx->copy_stream) {
        RS_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            RS_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            RS_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_used[i], cudaEventDisableTiming));
        }
    }
    const size_t tile_bytes = (size_t)tiles->height * tiles->width * C * elem_bytes(tiles->dtype);
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)C * R, zb = sizeof(uint32_t) * (size_t)R;
    const size_t sb = sizeof(double) * (size_t)(RS_NSTAT + n_pct) * C * R;
    const int per = tiles_per_chunk < tiles->n_tiles ? tiles_per_chunk : (tiles->n_tiles > 0 ? tiles->n_tiles : 1);
    if ((rc = ensure(ctx, ctx->stage[7], tile_bytes * per))) return rc;
    if ((rc = ensure(ctx, ctx->stage[8], tile_bytes * per))) return rc;
    if ((rc = ensure(ctx, ctx->stage[9], hb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], zb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sb))) return rc;
    cudaStream_t st = ctx->host_stream, cs = ctx->copy_stream;
    RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[9].p, 0, hb, st));
    RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[10].p, 0, zb, st));
    rs_zonal_params p = *prm;
    const uint32_t *aux = (const uint32_t *)ctx->stage[10].p;
    p.min_zero = nullptr;
    if (nodata_mode == RS_NODATA_ZERO) {
        if ((rc = ensure(ctx, ctx->stage[12], zb))) return rc;
        aux = p.min_zero = (uint32_t *)ctx->stage[12].p;
        RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[12].p, 0, zb, st));
    }
    int k = 0;
    for (int lo = 0; lo < tiles->n_tiles; lo += per, k++) {
        const int hi = lo + per < tiles->n_tiles ? lo + per : tiles->n_tiles, b = k & 1;
        void *buf = b ? ctx->stage[8].p : ctx->stage[7].p;
        if (k >= 2) RS_CUDA_OK(ctx, cudaStreamWaitEvent(cs, ctx->ev_used[b], 0));          // the kernel of chunk k-2 is done with it
        if ((rc = copy_h2d(ctx, buf, (const uint8_t *)tiles->pixels + (size_t)lo * tile_bytes, (size_t)(hi - lo) * tile_bytes, cs)))
            return rc;
        RS_CUDA_OK(ctx, cudaEventRecord(ctx->ev_copied[b], cs));
        RS_CUDA_OK(ctx, cudaStreamWaitEvent(st, ctx->ev_copied[b], 0));
        rs_tiles chunk = dt;
        chunk.pixels = (const uint8_t *)buf - (size_t)lo * tile_bytes;                     // indexed with the global tile index
        rc = launch_zonal_chunk(ctx, &dr, &chunk, &dp, &p, (uint32_t *)ctx->stage[9].p, (uint32_t *)ctx->stage[10].p, nullptr,
                                prm->window_mode, lo, hi, 1, st);
        if (rc) return rc;
        RS_CUDA_OK(ctx, cudaEventRecord(ctx->ev_used[b], st));
    }
    rc = launch_finalize(ctx, (const uint32_t *)ctx->stage[9].p, aux, R, C, nodata_mode, ddof,
                         percentiles, n_pct, (double *)ctx->stage[11].p, st);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(stats, ctx->stage[11].p, sb, cudaMemcpyDeviceToHost, st));
    if (hist) RS_CUDA_OK(ctx, cudaMemcpyAsync(hist, ctx->stage[9].p, hb, cudaMemcpyDeviceToHost, st));
    if (n_allzero) RS_CUDA_OK(ctx, cudaMemcpyAsync(n_allzero, ctx->stage[10].p, zb, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_zonal_stats_mapped_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                               const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles,
                               int32_t n_pct, double *stats, uint32_t *hist, uint32_t *n_allzero)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!prm || !stats || prm->road_slot || n_pct < 0 || n_pct > 16) return RS_ERR_INVALID_ARG;
    if (prm->hist_mode != RS_HIST_BANDS) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if (!roads || !tiles || !pairs) return RS_ERR_INVALID_ARG;
    // the tiles must be page-locked (cudaHostAlloc / cudaHostRegister): the kernel reads them in place over PCIe / C2C
    // (checked before anything is staged, so that a caller probing the transport pays nothing for a refusal)
    cudaPointerAttributes attr;
    attr.devicePointer = nullptr;
    if (tiles->n_tiles > 0 && roads->n_roads > 0) {
        if (!tiles->pixels) return RS_ERR_INVALID_ARG;
        if (cudaPointerGetAttributes(&attr, tiles->pixels) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
            cudaGetLastError();
            return RS_ERR_NOT_PINNED;
        }
    }
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, false, dr, dt, dp))) return rc;     // geometry + pairs + transforms, no pixels
    const int R = roads->n_roads, C = tiles->channels;
    if (R == 0) return RS_OK;
    dt.pixels = attr.devicePointer;
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)C * R, zb = sizeof(uint32_t) * (size_t)R;
    const size_t sb = sizeof(double) * (size_t)(RS_NSTAT + n_pct) * C * R;
    if ((rc = ensure(ctx, ctx->stage[9], hb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], zb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sb))) return rc;
    cudaStream_t st = ctx->host_stream;
    rs_zonal_params p = *prm;
    const uint32_t *aux = (const uint32_t *)ctx->stage[10].p;
    p.min_zero = nullptr;
    if (nodata_mode == RS_NODATA_ZERO) {        // the per-call zero padding of get_pixel_values needs the per-pair minima
        if ((rc = ensure(ctx, ctx->stage[12], zb))) return rc;
        aux = p.min_zero = (uint32_t *)ctx->stage[12].p;
    }
    rc = launch_zonal(ctx, &dr, &dt, &dp, &p, (uint32_t *)ctx->stage[9].p, (uint32_t *)ctx->stage[10].p, nullptr,
                      prm->window_mode, st);
    if (rc) return rc;
    rc = launch_finalize(ctx, (const uint32_t *)ctx->stage[9].p, aux, R, C, nodata_mode, ddof,
                         percentiles, n_pct, (double *)ctx->stage[11].p, st);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(stats, ctx->stage[11].p, sb, cudaMemcpyDeviceToHost, st));
    if (hist) RS_CUDA_OK(ctx, cudaMemcpyAsync(hist, ctx->stage[9].p, hb, cudaMemcpyDeviceToHost, st));
    if (n_allzero) RS_CUDA_OK(ctx, cudaMemcpyAsync(n_allzero, ctx->stage[10].p, zb, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_rasterize_pairs_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                            int window_mode, uint8_t *masks)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!pairs || !tiles) return RS_ERR_INVALID_ARG;
    if (pairs->n_pairs == 0) return RS_OK;
    if (!masks) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, false, dr, dt, dp))) return rc;
    const size_t mb = (size_t)pairs->n_pairs * tiles->height * tiles->width;
    if ((rc = ensure(ctx, ctx->stage[9], mb))) return rc;
    RS_CUDA_OK(ctx, cudaMemsetAsync(ctx->stage[9].p, 0, mb, ctx->host_stream));
    rc = launch_zonal(ctx, &dr, &dt, &dp, nullptr, nullptr, nullptr, (uint8_t *)ctx->stage[9].p, window_mode, ctx->host_stream);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(masks, ctx->stage[9].p, mb, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_finalize_stats_dev(rs_ctx *ctx, const uint32_t *hist, const uint32_t *n_allzero, int32_t n_roads, int32_t channels,
                          int32_t nodata_mode, int32_t ddof, const double *percentiles_host, int32_t n_pct, double *stats,
                          void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    return launch_finalize(ctx, hist, n_allzero, n_roads, channels, nodata_mode, ddof, percentiles_host, n_pct, stats,
                           (cudaStream_t)stream);
}

int rs_finalize_stats_host(rs_ctx *ctx, const uint32_t *hist, const uint32_t *n_allzero, int32_t n_roads, int32_t channels,
                           int32_t nodata_mode, int32_t ddof, const double *percentiles, int32_t n_pct, double *stats)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || channels < 1 || channels > 4 || n_pct < 0 || n_pct > 16) return RS_ERR_INVALID_ARG;
    if (n_roads == 0) return RS_OK;
    if (!hist || !stats) return RS_ERR_INVALID_ARG;
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)channels * n_roads;
    const size_t sb = sizeof(double) * (size_t)(RS_NSTAT + n_pct) * channels * n_roads;
    if ((rc = up(ctx, ctx->stage[9], hist, hb))) return rc;
    if (n_allzero && (rc = up(ctx, ctx->stage[10], n_allzero, sizeof(uint32_t) * (size_t)n_roads))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sb))) return rc;
    rc = launch_finalize(ctx, (const uint32_t *)ctx->stage[9].p, n_allzero ? (const uint32_t *)ctx->stage[10].p : nullptr,
                         n_roads, channels, nodata_mode, ddof, percentiles, n_pct, (double *)ctx->stage[11].p, ctx->host_stream);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(stats, ctx->stage[11].p, sb, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_vote_metrics_dev(rs_ctx *ctx, const uint32_t *joint_hist, const int8_t *gt_class, int32_t n_roads,
                        const int32_t *cutoffs_host, int32_t n_thr, int32_t rule, double min_area_frac, int8_t *cover,
                        double *scores, int64_t *confusion, double *metrics, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    return launch_vote(ctx, joint_hist, gt_class, n_roads, cutoffs_host, n_thr, rule, min_area_frac, cover, scores, confusion,
                       metrics, (cudaStream_t)stream);
}

int rs_vote_metrics_host(rs_ctx *ctx, const uint32_t *joint_hist, const int8_t *gt_class, int32_t n_roads,
                         const int32_t *cutoffs, int32_t n_thr, int32_t rule, double min_area_frac, int8_t *cover,
                         double *scores, int64_t *confusion, double *metrics)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || n_thr < 1 || n_thr > 32 || !cutoffs || !confusion) return RS_ERR_INVALID_ARG;
    const size_t R = (size_t)n_roads, T = (size_t)n_thr;
    if ((rc = up(ctx, ctx->stage[9], joint_hist, sizeof(uint32_t) * 768 * R))) return rc;
    if (gt_class && (rc = up(ctx, ctx->stage[10], gt_class, R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], T * R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[12], sizeof(double) * 3 * T * R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[13], sizeof(int64_t) * 8 * T))) return rc;
    if ((rc = ensure(ctx, ctx->stage[14], sizeof(double) * RS_NMETRIC * T))) return rc;
    rc = launch_vote(ctx, (const uint32_t *)ctx->stage[9].p, gt_class ? (const int8_t *)ctx->stage[10].p : nullptr, n_roads,
                     cutoffs, n_thr, rule, min_area_frac, cover ? (int8_t *)ctx->stage[11].p : nullptr,
                     scores ? (double *)ctx->stage[12].p : nullptr, (int64_t *)ctx->stage[13].p,
                     metrics ? (double *)ctx->stage[14].p : nullptr, ctx->host_stream);
    if (rc) return rc;
    cudaStream_t st = ctx->host_stream;
    if (cover && R) RS_CUDA_OK(ctx, cudaMemcpyAsync(cover, ctx->stage[11].p, T * R, cudaMemcpyDeviceToHost, st));
    if (scores && R) RS_CUDA_OK(ctx, cudaMemcpyAsync(scores, ctx->stage[12].p, sizeof(double) * 3 * T * R, cudaMemcpyDeviceToHost, st));
    RS_CUDA_OK(ctx, cudaMemcpyAsync(confusion, ctx->stage[13].p, sizeof(int64_t) * 8 * T, cudaMemcpyDeviceToHost, st));
    if (metrics) RS_CUDA_OK(ctx, cudaMemcpyAsync(metrics, ctx->stage[14].p, sizeof(double) * RS_NMETRIC * T, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_extract_pixels_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs, int window_mode,
                           int64_t *pair_off, void *values, int64_t capacity_pixels, int64_t *n_total)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!pairs || !tiles || !roads || !pair_off || !n_total) return RS_ERR_INVALID_ARG;
    *n_total = 0;
    const int P = pairs->n_pairs;
    if (P == 0) { pair_off[0] = 0; return RS_OK; }
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, values != nullptr, dr, dt, dp))) return rc;
    cudaStream_t st = ctx->host_stream;
    // count pass -> per-pair offsets (exclusive scan); the write pass walks every pair in raster order behind its offset.
    // No P x H x W mask is materialised: scratch is 12 bytes per pair.
    if ((rc = ensure(ctx, ctx->stage[9], sizeof(uint32_t) * (size_t)P))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], sizeof(unsigned long long) * ((size_t)P + 1)))) return rc;
    uint32_t *cnt = (uint32_t *)ctx->stage[9].p;
    unsigned long long *off = (unsigned long long *)ctx->stage[10].p;
    const int bpp = tiles->channels * (int)elem_bytes(tiles->dtype);
    RS_CUDA_OK(ctx, cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)P, st));
    if ((rc = launch_zonal_extract(ctx, &dr, &dt, &dp, window_mode, cnt, off, nullptr, bpp, 0, st))) return rc;
    if ((rc = launch_fstats_offsets(ctx, cnt, P, off, st))) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(pair_off, off, sizeof(int64_t) * ((size_t)P + 1), cudaMemcpyDeviceToHost, st));
    if ((rc = finish(ctx))) return rc;
    const int64_t total = pair_off[P];
    *n_total = total;
    if (!values || capacity_pixels < total || total == 0) return RS_OK;
    if ((rc = ensure(ctx, ctx->stage[11], (size_t)total * bpp))) return rc;
    if ((rc = launch_zonal_extract(ctx, &dr, &dt, &dp, window_mode, cnt, off, (uint8_t *)ctx->stage[11].p, bpp, 1, st))) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(values, ctx->stage[11].p, (size_t)total * bpp, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_group_hist_host(rs_ctx *ctx, const uint8_t *values, const int32_t *group, int64_t n, int32_t n_groups, uint32_t *hist)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n < 0 || n_groups < 0 || (n > 0 && (!values || !group)) || (n_groups > 0 && !hist)) return RS_ERR_INVALID_ARG;
    if (n_groups == 0) return RS_OK;
    if ((rc = up(ctx, ctx->stage[9], values, (size_t)n))) return rc;
    if ((rc = up(ctx, ctx->stage[10], group, sizeof(int32_t) * (size_t)n))) return rc;
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)n_groups;
    if ((rc = ensure(ctx, ctx->stage[11], hb))) return rc;
    if ((rc = launch_group_hist(ctx, (const uint8_t *)ctx->stage[9].p, (const int *)ctx->stage[10].p, n, n_groups,
                                (uint32_t *)ctx->stage[11].p, ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(hist, ctx->stage[11].p, hb, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_band_ratio_columns(int32_t channels) { return channels < 2 || channels > 4 ? 0 : channels * (channels - 1) / 2 + (channels == 4); }

int rs_band_ratios_dev(rs_ctx *ctx, const uint8_t *values, int64_t n, int32_t channels, double *out, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!values || !out))) return RS_ERR_INVALID_ARG;
    return launch_band_ratios(ctx, values, n, channels, out, (cudaStream_t)stream);
}

int rs_band_ratios_host(rs_ctx *ctx, const uint8_t *values, int64_t n, int32_t channels, double *out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!values || !out))) return RS_ERR_INVALID_ARG;
    const int K = rs_band_ratio_columns(channels);
    if (K == 0) return RS_ERR_UNSUPPORTED;
    if (n == 0) return RS_OK;
    if ((rc = up(ctx, ctx->stage[9], values, (size_t)n * channels))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], sizeof(double) * (size_t)K * n))) return rc;
    if ((rc = launch_band_ratios(ctx, (const uint8_t *)ctx->stage[9].p, n, channels, (double *)ctx->stage[10].p, ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->stage[10].p, sizeof(double) * (size_t)K * n, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_bin_counts_host(rs_ctx *ctx, const double *values, const int8_t *sel, const int8_t *hit, const int32_t *group, int32_t n,
                       int32_t n_cols, int32_t n_groups, const double *lo, const double *hi, int32_t n_thr, int64_t *counts)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n < 0 || n_cols < 1 || n_groups < 1 || n_thr < 1 || !lo || !hi || !counts) return RS_ERR_INVALID_ARG;
    if (n > 0 && (!values || !sel || !hit || !group)) return RS_ERR_INVALID_ARG;
    const size_t nk = (size_t)n * n_cols, nc = 2 * (size_t)n_groups * n_cols * n_thr;
    if ((rc = up(ctx, ctx->stage[9], values, sizeof(double) * nk))) return rc;
    if ((rc = up(ctx, ctx->stage[10], sel, nk))) return rc;
    if ((rc = up(ctx, ctx->stage[11], hit, nk))) return rc;
    if ((rc = up(ctx, ctx->stage[12], group, sizeof(int32_t) * (size_t)n))) return rc;
    if ((rc = up(ctx, ctx->stage[13], lo, sizeof(double) * (size_t)n_thr))) return rc;
    if ((rc = up(ctx, ctx->stage[14], hi, sizeof(double) * (size_t)n_thr))) return rc;
    if ((rc = ensure(ctx, ctx->stage[15], sizeof(int64_t) * nc))) return rc;
    rc = launch_bin_counts(ctx, (const double *)ctx->stage[9].p, (const int8_t *)ctx->stage[10].p, (const int8_t *)ctx->stage[11].p,
                           (const int *)ctx->stage[12].p, n, n_cols, n_groups, (const double *)ctx->stage[13].p,
                           (const double *)ctx->stage[14].p, n_thr, (int64_t *)ctx->stage[15].p, ctx->host_stream);
    if (rc) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(counts, ctx->stage[15].p, sizeof(int64_t) * nc, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_vote_table_host(rs_ctx *ctx, const int32_t *row_off, const int8_t *cls, const double *score, const double *weighted,
                       const double *area, int32_t n_roads, const double *thresholds, int32_t n_thr, int8_t *cover, double *scores)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || n_thr < 1 || n_thr > 32 || !thresholds || (n_roads > 0 && (!row_off || !cover))) return RS_ERR_INVALID_ARG;
    if (n_roads == 0) return RS_OK;
    const size_t N = (size_t)row_off[n_roads], R = (size_t)n_roads, T = (size_t)n_thr;
    if (N > 0 && (!cls || !score || !weighted || !area)) return RS_ERR_INVALID_ARG;
    if ((rc = up(ctx, ctx->stage[0], row_off, sizeof(int32_t) * (R + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[1], cls, N))) return rc;
    if ((rc = up(ctx, ctx->stage[2], score, sizeof(double) * N))) return rc;
    if ((rc = up(ctx, ctx->stage[3], weighted, sizeof(double) * N))) return rc;
    if ((rc = up(ctx, ctx->stage[4], area, sizeof(double) * N))) return rc;
    if ((rc = up(ctx, ctx->stage[5], thresholds, sizeof(double) * T))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], T * R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[12], sizeof(double) * 3 * T * R))) return rc;
    if ((rc = launch_vote_table(ctx, (const int *)ctx->stage[0].p, (const int8_t *)ctx->stage[1].p, (const double *)ctx->stage[2].p,
                                (const double *)ctx->stage[3].p, (const double *)ctx->stage[4].p, n_roads,
                                (const double *)ctx->stage[5].p, n_thr, (int8_t *)ctx->stage[11].p,
                                scores ? (double *)ctx->stage[12].p : nullptr, ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(cover, ctx->stage[11].p, T * R, cudaMemcpyDeviceToHost, ctx->host_stream));
    if (scores)
        RS_CUDA_OK(ctx, cudaMemcpyAsync(scores, ctx->stage[12].p, sizeof(double) * 3 * T * R, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_confusion_metrics_host(rs_ctx *ctx, const int8_t *cover, const int8_t *gt_class, int32_t n_roads, int32_t n_thr,
                              int64_t *confusion, double *metrics)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || n_thr < 1 || n_thr > 32 || !confusion || (n_roads > 0 && (!cover || !gt_class))) return RS_ERR_INVALID_ARG;
    const size_t R = (size_t)n_roads, T = (size_t)n_thr;
    if ((rc = up(ctx, ctx->stage[11], cover, T * R))) return rc;
    if ((rc = up(ctx, ctx->stage[10], gt_class, R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[13], sizeof(int64_t) * 8 * T))) return rc;
    if ((rc = ensure(ctx, ctx->stage[14], sizeof(double) * RS_NMETRIC * T))) return rc;
    if ((rc = launch_confusion(ctx, (const int8_t *)ctx->stage[11].p, (const int8_t *)ctx->stage[10].p, n_roads, n_thr,
                               (int64_t *)ctx->stage[13].p, metrics ? (double *)ctx->stage[14].p : nullptr, ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(confusion, ctx->stage[13].p, sizeof(int64_t) * 8 * T, cudaMemcpyDeviceToHost, ctx->host_stream));
    if (metrics)
        RS_CUDA_OK(ctx, cudaMemcpyAsync(metrics, ctx->stage[14].p, sizeof(double) * RS_NMETRIC * T, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_pairs_bbox_host(rs_ctx *ctx, const double *road_bbox, int32_t n_roads, const double *tile_ext, int32_t n_tiles,
                       const rs_lattice *lattice, int32_t *road_pair_off, int32_t *pair_tile, int64_t capacity, int64_t *n_pairs)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || n_tiles < 0 || !lattice || !road_pair_off || !n_pairs) return RS_ERR_INVALID_ARG;
    if (lattice->nx < 1 || lattice->ny < 1 || !(lattice->tile_w > 0.0) || !(lattice->tile_h > 0.0) || !lattice->lut) return RS_ERR_INVALID_ARG;
    if (n_roads > 0 && !road_bbox) return RS_ERR_INVALID_ARG;
    if (n_tiles > 0 && !tile_ext) return RS_ERR_INVALID_ARG;
    *n_pairs = 0;
    const size_t R = (size_t)n_roads;
    if ((rc = up(ctx, ctx->stage[0], road_bbox, sizeof(double) * 4 * R))) return rc;
    if ((rc = up(ctx, ctx->stage[1], tile_ext, sizeof(double) * 4 * (size_t)n_tiles))) return rc;
    if ((rc = up(ctx, ctx->stage[2], lattice->lut, sizeof(int32_t) * (size_t)lattice->nx * lattice->ny))) return rc;
    if ((rc = ensure(ctx, ctx->stage[3], sizeof(int32_t) * (R + 1)))) return rc;
    cudaStream_t st = ctx->host_stream;
    if ((rc = launch_pairs_bbox(ctx, (const double *)ctx->stage[0].p, n_roads, (const double *)ctx->stage[1].p, lattice,
                                (const int *)ctx->stage[2].p, (int *)ctx->stage[3].p, nullptr, 0, 0, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(road_pair_off, ctx->stage[3].p, sizeof(int32_t) * (R + 1), cudaMemcpyDeviceToHost, st));
    if ((rc = finish(ctx))) return rc;
    const int64_t total = road_pair_off[n_roads];
    *n_pairs = total;
    if (!pair_tile || capacity < total || total == 0) return RS_OK;
    if ((rc = ensure(ctx, ctx->stage[4], sizeof(int32_t) * (size_t)total))) return rc;
    if ((rc = launch_pairs_bbox(ctx, (const double *)ctx->stage[0].p, n_roads, (const double *)ctx->stage[1].p, lattice,
                                (const int *)ctx->stage[2].p, (int *)ctx->stage[3].p, (int *)ctx->stage[4].p, total, 1, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(pair_tile, ctx->stage[4].p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_pairs_bbox_grid_host(rs_ctx *ctx, const double *road_bbox, int32_t n_roads, const double *tile_ext, int32_t n_tiles,
                            int32_t *road_pair_off, int32_t *pair_tile, int64_t capacity, int64_t *n_pairs)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || n_tiles < 0 || !road_pair_off || !n_pairs) return RS_ERR_INVALID_ARG;
    if ((n_roads > 0 && !road_bbox) || (n_tiles > 0 && !tile_ext)) return RS_ERR_INVALID_ARG;
    *n_pairs = 0;
    // uniform grid over the tiles: a cell is about one (mean) tile
    double X0 = 0, Y0 = 0, X1 = 0, Y1 = 0, sw = 0, sh = 0;
    int nv = 0;
    for (int t = 0; t < n_tiles; t++) {
        const double *e = tile_ext + 4 * (size_t)t;
        if (!(e[0] <= e[2]) || !(e[1] <= e[3])) continue;
        if (!nv) { X0 = e[0]; Y0 = e[1]; X1 = e[2]; Y1 = e[3]; }
        X0 = e[0] < X0 ? e[0] : X0; Y0 = e[1] < Y0 ? e[1] : Y0;
        X1 = e[2] > X1 ? e[2] : X1; Y1 = e[3] > Y1 ? e[3] : Y1;
        sw += e[2] - e[0]; sh += e[3] - e[1];
        nv++;
    }
    double cw = nv ? sw / nv : 1.0, ch = nv ? sh / nv : 1.0;
    if (!(cw > 0.0)) cw = (X1 - X0) > 0.0 ? (X1 - X0) : 1.0;
    if (!(ch > 0.0)) ch = (Y1 - Y0) > 0.0 ? (Y1 - Y0) : 1.0;
    double fx = (X1 - X0) / cw + 1.0, fy = (Y1 - Y0) / ch + 1.0;
    while (fx * fy > 4.0e6) { cw *= 1.5; ch *= 1.5; fx = (X1 - X0) / cw + 1.0; fy = (Y1 - Y0) / ch + 1.0; }
    const int nx = (int)fx < 1 ? 1 : (int)fx, ny = (int)fy < 1 ? 1 : (int)fy;
    const size_t R = (size_t)n_roads;
    if ((rc = up(ctx, ctx->stage[0], road_bbox, sizeof(double) * 4 * R))) return rc;
    if ((rc = up(ctx, ctx->stage[1], tile_ext, sizeof(double) * 4 * (size_t)n_tiles))) return rc;
    if ((rc = ensure(ctx, ctx->stage[3], sizeof(int32_t) * (R + 1)))) return rc;
    cudaStream_t st = ctx->host_stream;
    if ((rc = launch_pairs_grid(ctx, (const double *)ctx->stage[0].p, n_roads, (const double *)ctx->stage[1].p, n_tiles, X0, Y0, cw, ch, nx,
                                ny, (int *)ctx->stage[3].p, nullptr, 0, 0, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(road_pair_off, ctx->stage[3].p, sizeof(int32_t) * (R + 1), cudaMemcpyDeviceToHost, st));
    if ((rc = finish(ctx))) return rc;
    const int64_t total = road_pair_off[n_roads];
    *n_pairs = total;
    if (!pair_tile || capacity < total || total == 0) return RS_OK;
    if ((rc = ensure(ctx, ctx->stage[4], sizeof(int32_t) * (size_t)total))) return rc;
    if ((rc = launch_pairs_grid(ctx, (const double *)ctx->stage[0].p, n_roads, (const double *)ctx->stage[1].p, n_tiles, X0, Y0, cw, ch, nx,
                                ny, (int *)ctx->stage[3].p, (int *)ctx->stage[4].p, total, 1, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(pair_tile, ctx->stage[4].p, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

static int stage_polys(rs_ctx *ctx, const rs_roads *r, int base, rs_roads &d);

int rs_pairs_intersect_host(rs_ctx *ctx, const rs_roads *roads, const double *tile_ext, int32_t n_tiles, const int32_t *road_pair_off,
                            const int32_t *pair_tile, int32_t n_pairs, uint8_t *keep)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!roads || n_tiles < 0 || n_pairs < 0) return RS_ERR_INVALID_ARG;
    if (n_pairs == 0) return RS_OK;
    if (!tile_ext || !road_pair_off || !pair_tile || !keep) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    if ((rc = stage_polys(ctx, roads, 0, dr))) return rc;
    if ((rc = up(ctx, ctx->stage[4], tile_ext, sizeof(double) * 4 * (size_t)n_tiles))) return rc;
    if ((rc = up(ctx, ctx->stage[5], road_pair_off, sizeof(int32_t) * ((size_t)roads->n_roads + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[6], pair_tile, sizeof(int32_t) * (size_t)n_pairs))) return rc;
    if ((rc = ensure(ctx, ctx->stage[7], (size_t)n_pairs))) return rc;
    cudaStream_t st = ctx->host_stream;
    if ((rc = launch_intersects(ctx, &dr, (const double *)ctx->stage[4].p, (const int *)ctx->stage[5].p, (const int *)ctx->stage[6].p,
                                n_pairs, (uint8_t *)ctx->stage[7].p, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(keep, ctx->stage[7].p, (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_decode_segments_dev(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec, uint8_t *raw,
                           const int64_t *raw_off, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_segments < 0) return RS_ERR_INVALID_ARG;
    if (n_segments == 0) return RS_OK;
    if (!comp || !comp_off || !raw || !raw_off) return RS_ERR_INVALID_ARG;
    return launch_decode_segments(ctx, comp, (const long long *)comp_off, n_segments, codec, raw, (const long long *)raw_off,
                                  (cudaStream_t)stream);
}

int rs_decode_segments_host(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec, uint8_t *raw,
                            const int64_t *raw_off)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_segments < 0) return RS_ERR_INVALID_ARG;
    if (n_segments == 0) return RS_OK;
    if (!comp || !comp_off || !raw || !raw_off) return RS_ERR_INVALID_ARG;
    const size_t cb = (size_t)comp_off[n_segments], rb = (size_t)raw_off[n_segments];
    if ((rc = up(ctx, ctx->stage[0], comp, cb))) return rc;
    if ((rc = up(ctx, ctx->stage[1], comp_off, sizeof(int64_t) * ((size_t)n_segments + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[2], raw_off, sizeof(int64_t) * ((size_t)n_segments + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->stage[3], rb))) return rc;
    cudaStream_t st = ctx->host_stream;
    if ((rc = launch_decode_segments(ctx, (const uint8_t *)ctx->stage[0].p, (const long long *)ctx->stage[1].p, n_segments, codec,
                                     (uint8_t *)ctx->stage[3].p, (const long long *)ctx->stage[2].p, st)))
        return rc;
    if (rb) RS_CUDA_OK(ctx, cudaMemcpyAsync(raw, ctx->stage[3].p, rb, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_clip_rings_host(rs_ctx *ctx, const rs_roads *labels, const int32_t *pair_label, const double *rect, int32_t n_pairs,
                       const int64_t *pair_ring_off, int32_t *ring_count, const int64_t *ring_vert_off, double *xy_out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!labels || n_pairs < 0) return RS_ERR_INVALID_ARG;
    if (n_pairs == 0) return RS_OK;
    if (!pair_label || !rect || !pair_ring_off || !ring_count) return RS_ERR_INVALID_ARG;
    if ((xy_out != nullptr) != (ring_vert_off != nullptr)) return RS_ERR_INVALID_ARG;
    const int64_t nq = pair_ring_off[n_pairs];
    if (nq < 0) return RS_ERR_INVALID_ARG;
    if (nq == 0) return RS_OK;
    rs_roads dl;
    if ((rc = stage_polys(ctx, labels, 0, dl))) return rc;
    if ((rc = up(ctx, ctx->stage[4], pair_label, sizeof(int32_t) * (size_t)n_pairs))) return rc;
    if ((rc = up(ctx, ctx->stage[5], rect, sizeof(double) * 4 * (size_t)n_pairs))) return rc;
    if ((rc = up(ctx, ctx->stage[6], pair_ring_off, sizeof(int64_t) * ((size_t)n_pairs + 1)))) return rc;
    cudaStream_t st = ctx->host_stream;
    if (!xy_out) {
        if ((rc = ensure(ctx, ctx->stage[7], sizeof(int32_t) * (size_t)nq))) return rc;
        if ((rc = launch_clip_rings(ctx, &dl, (const int *)ctx->stage[4].p, (const double *)ctx->stage[5].p, (const long long *)ctx->stage[6].p,
                                    n_pairs, nq, (int *)ctx->stage[7].p, nullptr, nullptr, st)))
            return rc;
        RS_CUDA_OK(ctx, cudaMemcpyAsync(ring_count, ctx->stage[7].p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
        return finish(ctx);
    }
    int64_t total = 0;
    for (int64_t q = 0; q < nq; q++) total = ring_vert_off[q] + ring_count[q] > total ? ring_vert_off[q] + ring_count[q] : total;
    if ((rc = up(ctx, ctx->stage[8], ring_vert_off, sizeof(int64_t) * (size_t)nq))) return rc;
    if ((rc = ensure(ctx, ctx->stage[9], sizeof(double) * 2 * (size_t)total))) return rc;
    if ((rc = launch_clip_rings(ctx, &dl, (const int *)ctx->stage[4].p, (const double *)ctx->stage[5].p, (const long long *)ctx->stage[6].p,
                                n_pairs, nq, nullptr, (const long long *)ctx->stage[8].p, (double *)ctx->stage[9].p, st)))
        return rc;
    if (total > 0) RS_CUDA_OK(ctx, cudaMemcpyAsync(xy_out, ctx->stage[9].p, sizeof(double) * 2 * (size_t)total, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

// stage one polygon set in stage[base .. base + 3] (xy, ring_off, road_ring_off, bbox)
static int stage_polys(rs_ctx *ctx, const rs_roads *r, int base, rs_roads &d)
{
    if (!r || r->n_roads < 0 || r->n_rings < 0 || r->n_verts < 0) return RS_ERR_INVALID_ARG;
    if (r->n_roads > 0 && (!r->xy || !r->ring_off || !r->road_ring_off)) return RS_ERR_INVALID_ARG;
    int rc;
    d = *r;
    if ((rc = up(ctx, ctx->stage[base], r->xy, sizeof(double) * 2 * (size_t)r->n_verts))) return rc;
    if ((rc = up(ctx, ctx->stage[base + 1], r->ring_off, sizeof(int32_t) * ((size_t)r->n_rings + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[base + 2], r->road_ring_off, sizeof(int32_t) * ((size_t)r->n_roads + 1)))) return rc;
    d.xy = (const double *)ctx->stage[base].p;
    d.ring_off = (const int32_t *)ctx->stage[base + 1].p;
    d.road_ring_off = (const int32_t *)ctx->stage[base + 2].p;
    if ((rc = ensure(ctx, ctx->stage[base + 3], sizeof(double) * 4 * (size_t)r->n_roads))) return rc;
    if (r->road_bbox) {
        if ((rc = up(ctx, ctx->stage[base + 3], r->road_bbox, sizeof(double) * 4 * (size_t)r->n_roads))) return rc;
    } else if ((rc = launch_road_bbox(ctx, &d, (double *)ctx->stage[base + 3].p, ctx->host_stream)))
        return rc;
    d.road_bbox = (const double *)ctx->stage[base + 3].p;
    return RS_OK;
}

int rs_zonal_stats_f32_host(rs_ctx *ctx, const rs_roads *features, const float *raster, int32_t height, int32_t width, const double *gt,
                            int32_t use_nodata, double nodata, int32_t ddof, const double *percentiles, int32_t n_pct, double *stats)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!features || height < 1 || width < 1 || !gt || ddof < 0 || n_pct < 0 || n_pct > 16 || (n_pct > 0 && !percentiles))
        return RS_ERR_INVALID_ARG;
    const int n = features->n_roads;
    if (n == 0) return RS_OK;
    if (!raster || !stats) return RS_ERR_INVALID_ARG;
    cudaStream_t st = ctx->host_stream;
    rs_roads dr;
    if ((rc = stage_polys(ctx, features, 0, dr))) return rc;
    const size_t px = (size_t)height * width;
    if ((rc = up(ctx, ctx->stage[7], raster, sizeof(float) * px))) return rc;
    if ((rc = up(ctx, ctx->stage[5], gt, sizeof(double) * 6))) return rc;
    // every feature against the one raster: pairs = identity
    int32_t *h = (int32_t *)malloc(sizeof(int32_t) * (2 * (size_t)n + 1));
    if (!h) return RS_ERR_INVALID_ARG;
    for (int i = 0; i <= n; i++) h[i] = i;
    for (int i = 0; i < n; i++) h[n + 1 + i] = 0;
    rc = up(ctx, ctx->stage[4], h, sizeof(int32_t) * (2 * (size_t)n + 1));
    if (!rc) RS_CUDA_OK(ctx, cudaStreamSynchronize(st));
    free(h);
    if (rc) return rc;
    rs_tiles dt{ctx->stage[7].p, (const double *)ctx->stage[5].p, 1, height, width, 1, RS_U8};
    rs_pairs dp{(const int32_t *)ctx->stage[4].p, (const int32_t *)ctx->stage[4].p + n + 1, n};
    const size_t NS = (size_t)RS_NSTAT + n_pct;
    if ((rc = ensure(ctx, ctx->stage[8], sizeof(uint32_t) * 2 * (size_t)n))) return rc;                    // counts | cursors
    if ((rc = ensure(ctx, ctx->stage[9], sizeof(unsigned long long) * ((size_t)n + 1)))) return rc;        // offsets
    if ((rc = ensure(ctx, ctx->stage[12], sizeof(double) * NS * n))) return rc;
    if ((rc = up(ctx, ctx->stage[13], percentiles, sizeof(double) * (size_t)n_pct))) return rc;
    uint32_t *counts = (uint32_t *)ctx->stage[8].p, *cursor = counts + n;
    unsigned long long *off = (unsigned long long *)ctx->stage[9].p;
    RS_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * 2 * (size_t)n, st));
    // rasterstats reads a boundless window and masks what lies off the raster: same pixels as the window clipped to the raster
    if ((rc = launch_zonal_f32(ctx, &dr, &dt, &dp, RS_WINDOW_BOUNDLESS, nodata, use_nodata, counts, cursor, off, nullptr, 0, st))) return rc;
    if ((rc = launch_fstats_offsets(ctx, counts, n, off, st))) return rc;
    unsigned long long total = 0;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(&total, off + n, sizeof(total), cudaMemcpyDeviceToHost, st));
    if ((rc = finish(ctx))) return rc;
    if (total > 0x7fffffffull) return RS_ERR_UNSUPPORTED;
    if ((rc = ensure(ctx, ctx->stage[10], sizeof(float) * (size_t)total))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sizeof(float) * (size_t)total))) return rc;
    if (total > 0) {
        if ((rc = launch_zonal_f32(ctx, &dr, &dt, &dp, RS_WINDOW_BOUNDLESS, nodata, use_nodata, counts, cursor, off,
                                   (float *)ctx->stage[10].p, 1, st)))
            return rc;
        if ((rc = launch_fstats_sort(ctx, (const float *)ctx->stage[10].p, (float *)ctx->stage[11].p, (long long)total, n, off, st))) return rc;
    }
    if ((rc = launch_fstats(ctx, (const float *)ctx->stage[11].p, off, n, ddof, (const double *)ctx->stage[13].p, n_pct,
                            (double *)ctx->stage[12].p, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(stats, ctx->stage[12].p, sizeof(double) * NS * n, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_within_host(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, uint8_t *within)
{
    int rc = bind(ctx);
    if (rc) return rc;
    rs_roads da, db;
    if ((rc = stage_polys(ctx, a, 0, da))) return rc;
    if ((rc = stage_polys(ctx, b, 4, db))) return rc;
    const size_t n = (size_t)a->n_roads * b->n_roads;
    if (n == 0) return RS_OK;
    if (!within) return RS_ERR_INVALID_ARG;
    if ((rc = ensure(ctx, ctx->stage[8], n))) return rc;
    if ((rc = launch_within(ctx, &da, &db, (uint8_t *)ctx->stage[8].p, ctx->host_stream))) return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(within, ctx->stage[8].p, n, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

// polygon index of every ring, built on the host from the CSR offsets and staged in stage[slot]
static int stage_ring_poly(rs_ctx *ctx, const rs_roads *r, int slot)
{
    int32_t *h = (int32_t *)malloc(sizeof(int32_t) * ((size_t)r->n_rings + 1));
    if (!h) return RS_ERR_INVALID_ARG;
    for (int i = 0; i < r->n_roads; i++)
        for (int g = r->road_ring_off[i]; g < r->road_ring_off[i + 1] && g < r->n_rings; g++) h[g] = i;
    int rc = ensure(ctx, ctx->stage[slot], sizeof(int32_t) * ((size_t)r->n_rings + 1));
    if (!rc && r->n_rings > 0) {
        cudaError_t e = cudaMemcpyAsync(ctx->stage[slot].p, h, sizeof(int32_t) * (size_t)r->n_rings, cudaMemcpyHostToDevice, ctx->host_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->host_stream);            // h is freed below
        if (e != cudaSuccess) { ctx->last_cuda_error = (int)e; rc = RS_ERR_CUDA; }
    }
    free(h);
    return rc;
}

int rs_overlay_area_host(rs_ctx *ctx, const rs_roads *a, const rs_roads *b, const int32_t *pair_a, const int32_t *pair_b,
                         int32_t n_pairs, double *area_pair, double *area_a)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_a || !pair_b || !area_pair))) return RS_ERR_INVALID_ARG;
    rs_roads da, db;
    if ((rc = stage_polys(ctx, a, 0, da))) return rc;
    if ((rc = stage_polys(ctx, b, 4, db))) return rc;
    for (int i = 0; i < n_pairs; i++)
        if (pair_a[i] < 0 || pair_a[i] >= a->n_roads || pair_b[i] < 0 || pair_b[i] >= b->n_roads) return RS_ERR_INVALID_ARG;
    if ((rc = stage_ring_poly(ctx, a, 8))) return rc;
    if ((rc = stage_ring_poly(ctx, b, 9))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], (size_t)a->n_rings + 1))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], (size_t)b->n_rings + 1))) return rc;
    if ((rc = up(ctx, ctx->stage[12], pair_a, sizeof(int32_t) * (size_t)n_pairs))) return rc;
    if ((rc = up(ctx, ctx->stage[13], pair_b, sizeof(int32_t) * (size_t)n_pairs))) return rc;
    if ((rc = ensure(ctx, ctx->stage[14], sizeof(double) * (size_t)n_pairs))) return rc;
    if ((rc = ensure(ctx, ctx->stage[15], sizeof(double) * (size_t)a->n_roads))) return rc;
    rc = launch_overlay_area(ctx, &da, &db, (const int *)ctx->stage[8].p, (const int *)ctx->stage[9].p, (int8_t *)ctx->stage[10].p,
                             (int8_t *)ctx->stage[11].p, (const int *)ctx->stage[12].p, (const int *)ctx->stage[13].p, n_pairs,
                             (double *)ctx->stage[14].p, area_a ? (double *)ctx->stage[15].p : nullptr, ctx->host_stream);
    if (rc) return rc;
    if (n_pairs > 0)
        RS_CUDA_OK(ctx, cudaMemcpyAsync(area_pair, ctx->stage[14].p, sizeof(double) * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->host_stream));
    if (area_a && a->n_roads > 0)
        RS_CUDA_OK(ctx, cudaMemcpyAsync(area_a, ctx->stage[15].p, sizeof(double) * (size_t)a->n_roads, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_ks_hist_host(rs_ctx *ctx, const uint32_t *hist, const int32_t *ref_of_road, const uint64_t *ref_hist, int32_t n_roads,
                    int32_t n_refs, double *D, double *n)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_roads < 0 || n_refs < 1 || !ref_hist || (n_roads > 0 && (!hist || !D))) return RS_ERR_INVALID_ARG;
    if (n_roads == 0) return RS_OK;
    const size_t R = (size_t)n_roads;
    if (ref_of_road)
        for (size_t i = 0; i < R; i++)
            if (ref_of_road[i] >= n_refs) return RS_ERR_INVALID_ARG;
    if ((rc = up(ctx, ctx->stage[9], hist, sizeof(uint32_t) * 256 * R))) return rc;
    if ((rc = up(ctx, ctx->stage[10], ref_hist, sizeof(uint64_t) * 256 * (size_t)n_refs))) return rc;
    if (ref_of_road && (rc = up(ctx, ctx->stage[8], ref_of_road, sizeof(int32_t) * R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sizeof(double) * R))) return rc;
    if ((rc = ensure(ctx, ctx->stage[12], sizeof(double) * R))) return rc;
    if ((rc = launch_ks(ctx, (const uint32_t *)ctx->stage[9].p, ref_of_road ? (const int *)ctx->stage[8].p : nullptr,
                        (const unsigned long long *)ctx->stage[10].p, n_roads, 256, (double *)ctx->stage[11].p,
                        (double *)ctx->stage[12].p, ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(D, ctx->stage[11].p, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->host_stream));
    if (n) RS_CUDA_OK(ctx, cudaMemcpyAsync(n, ctx->stage[12].p, sizeof(double) * R, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_rescale_u16_dev(rs_ctx *ctx, const uint16_t *src, int64_t n_pixels, int32_t c_in, int32_t c_out, const int32_t *bidx,
                       const double *k, const double *off, int32_t f32, uint8_t *dst, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    return launch_rescale(ctx, src, n_pixels, c_in, c_out, bidx, k, off, f32, dst, (cudaStream_t)stream);
}

int rs_rescale_u16_host(rs_ctx *ctx, const uint16_t *src, int64_t n_pixels, int32_t c_in, int32_t c_out, const int32_t *bidx,
                        const double *k, const double *off, int32_t f32, uint8_t *dst)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_pixels < 0 || c_in < 1 || c_in > 4 || c_out < 1 || c_out > 4) return RS_ERR_INVALID_ARG;
    if (n_pixels == 0) return RS_OK;
    if (!src || !dst) return RS_ERR_INVALID_ARG;
    const size_t ib = sizeof(uint16_t) * (size_t)n_pixels * c_in, ob = (size_t)n_pixels * c_out;
    if ((rc = up(ctx, ctx->stage[7], src, ib))) return rc;
    if ((rc = ensure(ctx, ctx->stage[9], ob))) return rc;
    if ((rc = launch_rescale(ctx, (const uint16_t *)ctx->stage[7].p, n_pixels, c_in, c_out, bidx, k, off, f32, (uint8_t *)ctx->stage[9].p,
                             ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(dst, ctx->stage[9].p, ob, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

int rs_assemble_tiles_dev(rs_ctx *ctx, const uint8_t *raw, int32_t n_tiles, int32_t height, int32_t width, int32_t c_in,
                          int32_t planar, int32_t predictor, int32_t sample_bytes, int32_t big_endian, int32_t c_out,
                          const int32_t *bidx, int32_t rescale, const double *k, const double *off, void *out, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    return launch_assemble(ctx, raw, n_tiles, height, width, c_in, planar, predictor, sample_bytes, big_endian, c_out, bidx, rescale,
                           k, off, out, (cudaStream_t)stream);
}

int rs_assemble_tiles_host(rs_ctx *ctx, const uint8_t *raw, int32_t n_tiles, int32_t height, int32_t width, int32_t c_in,
                           int32_t planar, int32_t predictor, int32_t sample_bytes, int32_t big_endian, int32_t c_out,
                           const int32_t *bidx, int32_t rescale, const double *k, const double *off, void *out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_tiles < 0 || height < 1 || width < 1 || c_in < 1 || c_out < 1 || (sample_bytes != 1 && sample_bytes != 2)) return RS_ERR_INVALID_ARG;
    if (n_tiles == 0) return RS_OK;
    if (!raw || !out) return RS_ERR_INVALID_ARG;
    const size_t npx = (size_t)n_tiles * height * width;
    const size_t in_b = npx * c_in * sample_bytes, out_b = npx * c_out * ((sample_bytes == 1 || rescale) ? 1 : 2);
    if ((rc = up(ctx, ctx->stage[7], raw, in_b))) return rc;
    if ((rc = ensure(ctx, ctx->stage[9], out_b))) return rc;
    if ((rc = launch_assemble(ctx, (const uint8_t *)ctx->stage[7].p, n_tiles, height, width, c_in, planar, predictor, sample_bytes,
                              big_endian, c_out, bidx, rescale, k, off, ctx->stage[9].p, ctx->host_stream)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->stage[9].p, out_b, cudaMemcpyDeviceToHost, ctx->host_stream));
    return finish(ctx);
}

}  // extern "C" (helper below is internal)

// Compressed segments -> decoded samples with the upload and the decoding overlapped: the compressed bytes cross the host link in
// up to eight pieces on the copy stream, and the decoder of a piece is queued on the host stream behind the event of its copy
// -- while it runs, the next piece is staged and copied.  d_comp / d_coff / d_roff: device buffers (offsets already uploaded).
static int decode_overlapped(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec,
                             uint8_t *d_comp, const long long *d_coff, uint8_t *d_raw, const long long *d_roff)
{
    int rc;
    if (!ctx->copy_stream) {
        RS_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            RS_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            RS_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_used[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t st = ctx->host_stream, cs = ctx->copy_stream;
    // a piece is at least a full wave of decoders (a launch lasts as long as its slowest segment, however few it has) and at
    // most an eighth of the call
    const int per_piece = (n_segments + 7) / 8 > 65536 ? (n_segments + 7) / 8 : 65536;
    // the copy stream starts behind whatever the host stream still does with these buffers (a previous call's decoder)
    RS_CUDA_OK(ctx, cudaEventRecord(ctx->ev_used[0], st));
    RS_CUDA_OK(ctx, cudaStreamWaitEvent(cs, ctx->ev_used[0], 0));
    int a = 0, k = 0;
    while (a < n_segments) {
        const int b = n_segments - a <= per_piece + per_piece / 4 ? n_segments : a + per_piece;      // segments [a, b)
        const int64_t lo = comp_off[a], hi = comp_off[b];
        if ((rc = copy_h2d(ctx, d_comp + lo, comp + lo, (size_t)(hi - lo), cs))) return rc;
        // two events in turn: a record waits (host side) until the decoder that waited on its previous use was queued -- it was,
        // two pieces ago, in this same thread
        RS_CUDA_OK(ctx, cudaEventRecord(ctx->ev_copied[k & 1], cs));
        RS_CUDA_OK(ctx, cudaStreamWaitEvent(st, ctx->ev_copied[k & 1], 0));
        if ((rc = launch_decode_segments(ctx, d_comp, d_coff + a, b - a, codec, d_raw, d_roff + a, st))) return rc;
        a = b;
        k++;
    }
    return RS_OK;
}

extern "C" {

int rs_ingest_tiles_host(rs_ctx *ctx, const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec,
                         const int64_t *raw_off, int32_t n_tiles, int32_t height, int32_t width, int32_t c_in, int32_t planar,
                         int32_t predictor, int32_t sample_bytes, int32_t big_endian, int32_t c_out, const int32_t *bidx, int32_t rescale,
                         const double *k, const double *off, void *out, int32_t keep_on_device, void **device_out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_tiles < 0 || n_segments < 0 || height < 1 || width < 1 || c_in < 1 || c_out < 1 || (sample_bytes != 1 && sample_bytes != 2))
        return RS_ERR_INVALID_ARG;
    if (n_tiles == 0) return RS_OK;
    if (!comp || !comp_off || !raw_off || (!out && !keep_on_device)) return RS_ERR_INVALID_ARG;
    const size_t npx = (size_t)n_tiles * height * width;
    const size_t in_b = npx * c_in * sample_bytes, out_b = npx * c_out * ((sample_bytes == 1 || rescale) ? 1 : 2);
    if ((size_t)raw_off[n_segments] != in_b) return RS_ERR_INVALID_ARG;           // the segments must tile the sample buffer exactly
    const size_t cb = (size_t)comp_off[n_segments];
    if ((rc = ensure(ctx, ctx->stage[0], cb))) return rc;                          // only the COMPRESSED bytes cross the host link
    if ((rc = up(ctx, ctx->stage[1], comp_off, sizeof(int64_t) * ((size_t)n_segments + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[2], raw_off, sizeof(int64_t) * ((size_t)n_segments + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->stage[7], in_b))) return rc;
    if ((rc = ensure(ctx, ctx->stage[9], out_b))) return rc;
    cudaStream_t st = ctx->host_stream;
    if ((rc = decode_overlapped(ctx, comp, comp_off, n_segments, codec, (uint8_t *)ctx->stage[0].p, (const long long *)ctx->stage[1].p,
                                (uint8_t *)ctx->stage[7].p, (const long long *)ctx->stage[2].p)))
        return rc;
    if ((rc = launch_assemble(ctx, (const uint8_t *)ctx->stage[7].p, n_tiles, height, width, c_in, planar, predictor, sample_bytes,
                              big_endian, c_out, bidx, rescale, k, off, ctx->stage[9].p, st)))
        return rc;
    if (out) RS_CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->stage[9].p, out_b, cudaMemcpyDeviceToHost, st));
    if (device_out) *device_out = keep_on_device ? ctx->stage[9].p : nullptr;      // valid until the next _host call on this context
    return finish(ctx);
}

int rs_zonal_stats_compressed_host(rs_ctx *ctx, const rs_roads *roads, const rs_tiles *tiles, const rs_pairs *pairs,
                                   const rs_zonal_params *prm, int32_t nodata_mode, int32_t ddof, const double *percentiles, int32_t n_pct,
                                   const uint8_t *comp, const int64_t *comp_off, int32_t n_segments, int32_t codec, const int64_t *raw_off,
                                   int32_t planar, int32_t predictor, int32_t big_endian, double *stats)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!prm || !stats || prm->road_slot || n_pct < 0 || n_pct > 16 || n_segments < 0) return RS_ERR_INVALID_ARG;
    if (prm->hist_mode != RS_HIST_BANDS || !tiles || !roads || !pairs) return RS_ERR_INVALID_ARG;
    if (tiles->dtype != RS_U8 || !comp || !comp_off || !raw_off) return RS_ERR_INVALID_ARG;
    rs_roads dr;
    rs_tiles dt;
    rs_pairs dp;
    if ((rc = stage_inputs(ctx, roads, tiles, pairs, false, dr, dt, dp))) return rc;     // geometry + pairs + transforms, no pixels
    const int R = roads->n_roads, C = tiles->channels;
    if (R == 0) return RS_OK;
    const size_t npx = (size_t)tiles->n_tiles * tiles->height * tiles->width, in_b = npx * C;
    if ((size_t)raw_off[n_segments] != in_b) return RS_ERR_INVALID_ARG;
    const size_t hb = sizeof(uint32_t) * 256 * (size_t)C * R, zb = sizeof(uint32_t) * (size_t)R;
    const size_t sb = sizeof(double) * (size_t)(RS_NSTAT + n_pct) * C * R;
    // stage[12..14]: compressed bytes + offsets (only these cross the host link); stage[7]: decoded samples; stage[8]: tiles
    if ((rc = ensure(ctx, ctx->stage[12], (size_t)comp_off[n_segments]))) return rc;
    if ((rc = up(ctx, ctx->stage[13], comp_off, sizeof(int64_t) * ((size_t)n_segments + 1)))) return rc;
    if ((rc = up(ctx, ctx->stage[14], raw_off, sizeof(int64_t) * ((size_t)n_segments + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->stage[7], in_b))) return rc;
    if ((rc = ensure(ctx, ctx->stage[8], in_b))) return rc;
    if ((rc = ensure(ctx, ctx->stage[9], hb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[10], zb))) return rc;
    if ((rc = ensure(ctx, ctx->stage[11], sb))) return rc;
    cudaStream_t st = ctx->host_stream;
    if ((rc = decode_overlapped(ctx, comp, comp_off, n_segments, codec, (uint8_t *)ctx->stage[12].p, (const long long *)ctx->stage[13].p,
                                (uint8_t *)ctx->stage[7].p, (const long long *)ctx->stage[14].p)))
        return rc;
    if ((rc = launch_assemble(ctx, (const uint8_t *)ctx->stage[7].p, tiles->n_tiles, tiles->height, tiles->width, C, planar, predictor, 1,
                              big_endian, C, nullptr, 0, nullptr, nullptr, ctx->stage[8].p, st)))
        return rc;
    dt.pixels = ctx->stage[8].p;
    rs_zonal_params p = *prm;
    const uint32_t *aux = (const uint32_t *)ctx->stage[10].p;
    p.min_zero = nullptr;
    if (nodata_mode == RS_NODATA_ZERO) {
        if ((rc = ensure(ctx, ctx->stage[15], zb))) return rc;
        aux = p.min_zero = (uint32_t *)ctx->stage[15].p;
    }
    if ((rc = launch_zonal(ctx, &dr, &dt, &dp, &p, (uint32_t *)ctx->stage[9].p, (uint32_t *)ctx->stage[10].p, nullptr, prm->window_mode, st)))
        return rc;
    if ((rc = launch_finalize(ctx, (const uint32_t *)ctx->stage[9].p, aux, R, C, nodata_mode, ddof, percentiles, n_pct,
                              (double *)ctx->stage[11].p, st)))
        return rc;
    RS_CUDA_OK(ctx, cudaMemcpyAsync(stats, ctx->stage[11].p, sb, cudaMemcpyDeviceToHost, st));
    return finish(ctx);
}

int rs_synth_tiles_dev(rs_ctx *ctx, void *pixels, const int64_t *tile_key, int32_t n_tiles, int32_t height, int32_t width,
                       int32_t channels, int32_t dtype, int32_t kind, uint64_t seed, void *stream)
{
    int rc = bind(ctx);
    if (rc) return rc;
    return launch_synth(ctx, pixels, tile_key, n_tiles, height, width, channels, dtype, kind, seed, (cudaStream_t)stream);
}

}  // extern "C"
