// How does the B200 fetch page-locked host memory over PCIe?  Sector reads per second for four access shapes:
//   A  one 8-byte load per lane, every lane in its own 128-byte line                    (1 sector / request)
//   B  lanes 2k, 2k+1 load the two sectors of one 64-byte half line in ONE instruction  (2 sectors / request if merged)
//   C  lanes 4k..4k+3 load the four sectors of one line in ONE instruction              (4 sectors / request if merged)
//   D  one lane loads two adjacent sectors with TWO instructions                         (what the span-driven kernel does)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sysmem_read sysmem_read.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <string.h>

template <int SHAPE>
__global__ void reader(const uint8_t *base, size_t n_lines, int iters, unsigned long long *sink)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (int it = 0; it < iters; it++) {
        // a warp's lane groups read different rows (768-byte stride) of ONE 192 KiB tile at a time, like the rows under a
        // road (page-local, TLB-friendly); no line is touched twice
        const size_t g = SHAPE == 1 ? tid >> 1 : SHAPE == 2 ? tid >> 2 : tid;
        const size_t gpw = SHAPE == 1 ? 16 : SHAPE == 2 ? 8 : 32;                   // groups per warp
        const size_t w = g / gpw, gl = g % gpw, nw = nthr / 32;
        const size_t rows_per_it = gpw, its_per_tile = 256 / rows_per_it;
        const size_t tile = w + (size_t)(it / its_per_tile) * nw;
        const size_t off = ((tile * 196608ull + ((size_t)(it % its_per_tile) * rows_per_it + gl) * 768ull) % (n_lines * 128)) & ~127ull;
        const uint8_t *p = base + off;
        if (SHAPE == 0) acc += *reinterpret_cast<const unsigned long long *>(p);
        if (SHAPE == 1) acc += *reinterpret_cast<const unsigned long long *>(p + 32 * (tid & 1));
        if (SHAPE == 2) acc += *reinterpret_cast<const unsigned long long *>(p + 32 * (tid & 3));
        if (SHAPE == 3) {
            acc += *reinterpret_cast<const unsigned long long *>(p);
            acc += *reinterpret_cast<const unsigned long long *>(p + 32);
        }
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
}

static void run_all(const char *what, uint8_t *h, size_t bytes, unsigned long long *sink)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256, iters = 64;
    const char *names[4] = {"A 1 sector / lane", "B 2 lanes -> 64 B", "C 4 lanes -> 128 B", "D 1 lane, 2 loads"};
    for (int s = 0; s < 4; s++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (s == 0) reader<0><<<blocks, threads>>>(h, bytes / 128, iters, sink);
            if (s == 1) reader<1><<<blocks, threads>>>(h, bytes / 128, iters, sink);
            if (s == 2) reader<2><<<blocks, threads>>>(h, bytes / 128, iters, sink);
            if (s == 3) reader<3><<<blocks, threads>>>(h, bytes / 128, iters, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double sectors = (double)blocks * threads * iters * (s == 3 ? 2 : 1);
            if (rep) printf("%-12s %-20s %8.2f ms  %7.1f M sectors/s  %6.2f GB/s of sectors\n", what, names[s], ms, sectors / ms / 1e3, sectors * 32 / ms / 1e6);
        }
    }
}

int main()
{
    const size_t bytes = 32ull << 30;
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    uint8_t *h;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    for (size_t i = 0; i < bytes; i += 4096) h[i] = (uint8_t)i;
    run_all("cudaHostAlloc", h, bytes, sink);
    cudaFreeHost(h);
    // the same buffer backed by transparent huge pages (2 MiB), page-locked with cudaHostRegister
    void *m = mmap(nullptr, bytes + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) { printf("mmap failed\n"); return 1; }
    uint8_t *hp = (uint8_t *)(((uintptr_t)m + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
    printf("madvise(MADV_HUGEPAGE) -> %d\n", madvise(hp, bytes, MADV_HUGEPAGE));
    for (size_t i = 0; i < bytes; i += 4096) hp[i] = (uint8_t)i;
    cudaError_t e = cudaHostRegister(hp, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    printf("cudaHostRegister -> %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) run_all("THP+register", hp, bytes, sink);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
