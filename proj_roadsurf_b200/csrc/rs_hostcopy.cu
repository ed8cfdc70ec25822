// Host -> device copies out of PAGEABLE caller memory (the numpy array an ordinary caller of the _host entry points holds --
// what rasterio.open(tile).read() returns in the reference's pair loop, fct_misc.py:76-77).  cudaMemcpy from pageable memory is
// staged by the driver through one small buffer on one thread (~11 GB/s measured, profiles/README.md); here a few host threads
// copy 16 MiB pieces into a ring of page-locked slots and every slot goes to the device with its own asynchronous DMA, so the
// host copy of piece k+1 runs under the DMA of piece k and the link, not one core, sets the rate.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "rs_internal.h"

namespace rs {

namespace {

constexpr int NSLOT = 4;
constexpr size_t SLOT_BYTES = (size_t)16 << 20;
constexpr size_t MIN_STAGED = (size_t)8 << 20;        // smaller copies are left to cudaMemcpyAsync

}  // namespace

struct HostCopier {
    void *slot[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    bool in_flight[NSLOT] = {false, false, false, false};
    int next = 0;
    int n_threads = 1;
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv, cv_done;
    uint64_t gen = 0;
    bool stop = false;
    int pending = 0;
    uint8_t *job_dst = nullptr;
    const uint8_t *job_src = nullptr;
    size_t job_len = 0;

    void slice(int t) const
    {
        // 4 KiB aligned cuts
        const size_t pages = (job_len + 4095) / 4096;
        const size_t a = pages * (size_t)t / (size_t)n_threads * 4096, b = pages * (size_t)(t + 1) / (size_t)n_threads * 4096;
        const size_t lo = a < job_len ? a : job_len, hi = b < job_len ? b : job_len;
        if (hi > lo) memcpy(job_dst + lo, job_src + lo, hi - lo);
    }
    void work(int t)
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
            }
            slice(t);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    void parallel_copy(void *dst, const void *src, size_t len)
    {
        job_dst = (uint8_t *)dst;
        job_src = (const uint8_t *)src;
        job_len = len;
        if (n_threads > 1) {
            {
                std::lock_guard<std::mutex> lk(m);
                pending = n_threads - 1;
                gen++;
            }
            cv.notify_all();
        }
        slice(0);
        if (n_threads > 1) {
            std::unique_lock<std::mutex> lk(m);
            cv_done.wait(lk, [&] { return pending == 0; });
        }
    }
};

static HostCopier *copier_of(rs_ctx *ctx)
{
    if (ctx->copier) return ctx->copier;
    HostCopier *h = new (std::nothrow) HostCopier();
    if (!h) return nullptr;
    for (int i = 0; i < NSLOT; i++) {
        if (cudaHostAlloc(&h->slot[i], SLOT_BYTES, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k <= i; k++) {
                if (h->slot[k]) cudaFreeHost(h->slot[k]);
                if (h->ev[k]) cudaEventDestroy(h->ev[k]);
            }
            delete h;
            return nullptr;
        }
    }
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = CPU_COUNT(&set);   // the cores this process may use
    if (n > 8) n = 8;
    const char *env = getenv("RS_STAGE_THREADS");
    if (env && atoi(env) >= 1 && atoi(env) <= 64) n = atoi(env);
    h->n_threads = n < 1 ? 1 : n;
    for (int t = 1; t < h->n_threads; t++) h->workers.emplace_back([h, t] { h->work(t); });
    ctx->copier = h;
    return h;
}

void copier_destroy(rs_ctx *ctx)
{
    HostCopier *h = ctx->copier;
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(h->m);
        h->stop = true;
    }
    h->cv.notify_all();
    for (auto &w : h->workers) w.join();
    for (int i = 0; i < NSLOT; i++) {
        if (h->in_flight[i]) cudaEventSynchronize(h->ev[i]);
        if (h->slot[i]) cudaFreeHost(h->slot[i]);
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    }
    delete h;
    ctx->copier = nullptr;
}

bool host_pageable(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// dst (device) <- src (host), queued on `st`.  Pageable sources of at least MIN_STAGED bytes go through the slot ring: when the call
// returns the source has been read completely (like cudaMemcpyAsync from pageable memory); page-locked sources are one DMA.
int copy_h2d(rs_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return RS_OK;
    const char *env = getenv("RS_STAGE_COPY");           // 0: always plain cudaMemcpyAsync
    HostCopier *h = nullptr;
    if (bytes >= MIN_STAGED && !(env && atoi(env) == 0) && host_pageable(src)) h = copier_of(ctx);
    if (!h) {
        RS_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return RS_OK;
    }
    for (size_t off = 0; off < bytes; off += SLOT_BYTES) {
        const size_t len = bytes - off < SLOT_BYTES ? bytes - off : SLOT_BYTES;
        const int i = h->next;
        h->next = (i + 1) % NSLOT;
        if (h->in_flight[i]) RS_CUDA_OK(ctx, cudaEventSynchronize(h->ev[i]));        // the DMA out of this slot, NSLOT pieces ago
        h->parallel_copy(h->slot[i], (const uint8_t *)src + off, len);
        RS_CUDA_OK(ctx, cudaMemcpyAsync((uint8_t *)dst + off, h->slot[i], len, cudaMemcpyHostToDevice, st));
        RS_CUDA_OK(ctx, cudaEventRecord(h->ev[i], st));
        h->in_flight[i] = true;
    }
    return RS_OK;
}

}  // namespace rs
