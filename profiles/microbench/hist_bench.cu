// Microbenchmark (B200): which per-warp 256-bin histogram update is fastest?
//   A  shared atomicAdd on a warp-private u32 histogram (3 bands), random / low-entropy values
//   B  lane-private u8 counters, layout [bin>>2][lane] words (bank == lane: conflict-free), plain LDS/IADD/STS
//   C  __match_any_sync leader + plain RMW on a warp-private u32 histogram
// Each lane processes `PX` pixels x 3 bands per round; values come from registers (xorshift), no global traffic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t xs(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k(int rounds, int lowent, uint32_t *out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t seed = (blockIdx.x * 1315423911u) ^ (threadIdx.x * 2654435761u) ^ 0x9E3779B9u;
    uint32_t acc = 0;
    if (MODE == 0 || MODE == 2) {
        uint32_t *h = reinterpret_cast<uint32_t *>(smem) + warp * 768;
        for (int i = lane; i < 768; i += 32) h[i] = 0;
        __syncwarp();
        for (int r = 0; r < rounds; r++) {
#pragma unroll
            for (int p = 0; p < 8; p++) {
                uint32_t v = xs(seed);
                uint32_t b0 = v & 255, b1 = (v >> 8) & 255, b2 = (v >> 16) & 255;
                if (lowent) { b0 = 100 + (b0 & 15); b1 = 100 + (b1 & 15); b2 = 100 + (b2 & 15); }
                if (MODE == 0) {
                    atomicAdd(&h[b0], 1u); atomicAdd(&h[256 + b1], 1u); atomicAdd(&h[512 + b2], 1u);
                } else {
                    uint32_t bb[3] = {b0, 256 + b1, 512 + b2};
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const unsigned m = __match_any_sync(0xffffffffu, bb[c]);
                        if ((__ffs(m) - 1) == lane) h[bb[c]] += __popc(m);
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < 768; i += 32) acc += h[i];
    } else {
        // lane-private u8 counters: word index = band*2048 + (bin>>2)*32 + lane, byte = bin & 3
        uint8_t *t = smem + warp * 24576;
        uint4 *z = reinterpret_cast<uint4 *>(t);
        for (int i = lane; i < 24576 / 16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        uint32_t tot[3] = {0, 0, 0};
        for (int r = 0; r < rounds; r++) {
#pragma unroll
            for (int p = 0; p < 8; p++) {
                uint32_t v = xs(seed);
                uint32_t b0 = v & 255, b1 = (v >> 8) & 255, b2 = (v >> 16) & 255;
                if (lowent) { b0 = 100 + (b0 & 15); b1 = 100 + (b1 & 15); b2 = 100 + (b2 & 15); }
                uint8_t *a0 = t + (((b0 >> 2) * 32 + lane) << 2) + (b0 & 3);
                uint8_t *a1 = t + 8192 + (((b1 >> 2) * 32 + lane) << 2) + (b1 & 3);
                uint8_t *a2 = t + 16384 + (((b2 >> 2) * 32 + lane) << 2) + (b2 & 3);
                *a0 = *a0 + 1; *a1 = *a1 + 1; *a2 = *a2 + 1;
            }
            if ((r % 31) == 30) {      // flush: lane L sums bins [8L, 8L+8) over the 32 lane copies (rotated: conflict-free)
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    uint32_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
                    const uint32_t *w = reinterpret_cast<const uint32_t *>(t + c * 8192);
                    for (int j = 0; j < 32; j++) {
                        const int jj = (j + lane) & 31;
                        const uint32_t x0 = w[(2 * lane) * 32 + jj], x1 = w[(2 * lane + 1) * 32 + jj];
                        lo0 += x0 & 0x00ff00ffu; hi0 += (x0 >> 8) & 0x00ff00ffu;
                        lo1 += x1 & 0x00ff00ffu; hi1 += (x1 >> 8) & 0x00ff00ffu;
                    }
                    tot[c] += (lo0 & 0xffff) + (lo0 >> 16) + (hi0 & 0xffff) + (hi0 >> 16) + (lo1 & 0xffff) + (lo1 >> 16) + (hi1 & 0xffff) + (hi1 >> 16);
                }
                __syncwarp();
                for (int i = lane; i < 24576 / 16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
                __syncwarp();
            }
        }
        acc = tot[0] + tot[1] + tot[2];
        __syncwarp();
        const uint32_t *w = reinterpret_cast<const uint32_t *>(t);
        for (int i = lane; i < 24576 / 4; i += 32) { uint32_t x = w[i]; acc += (x & 255) + ((x >> 8) & 255) + ((x >> 16) & 255) + (x >> 24); }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
}

template <int MODE, int WARPS>
int run(const char *name, size_t smem_per_warp, int lowent)
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    auto kern = k<MODE, WARPS>;
    const size_t smem = smem_per_warp * WARPS;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    uint32_t *out;
    CK(cudaMalloc(&out, 4));
    CK(cudaMemset(out, 0, 4));
    const int rounds = 31 * 40, grid = sms * per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, WARPS * 32, smem>>>(rounds, lowent, out);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(out, 0, 4));
    cudaEventRecord(e0);
    kern<<<grid, WARPS * 32, smem>>>(rounds, lowent, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    uint32_t h; CK(cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost));
    const double px = (double)grid * WARPS * 32 * rounds * 8;
    printf("%-34s lowent=%d warps/SM=%2d  %8.3f ms  %8.1f Gpx/s (x3 bands = %.1f G updates/s)  check=%u (expect %u)\n", name, lowent,
           per_sm * WARPS, ms, px / ms / 1e6, 3 * px / ms / 1e6, h, (uint32_t)(3 * px));
    cudaFree(out);
    return 0;
}

int main()
{
    for (int le = 0; le < 2; le++) {
        if (run<0, 4>("A shared atomicAdd u32", 3072, le)) return 1;
        if (run<0, 16>("A shared atomicAdd u32 (16w CTA)", 3072, le)) return 1;
        if (run<1, 1>("B lane-private u8 RMW (1w CTA)", 24576, le)) return 1;
        if (run<1, 2>("B lane-private u8 RMW (2w CTA)", 24576, le)) return 1;
        if (run<2, 4>("C match_any + plain RMW", 3072, le)) return 1;
    }
    return 0;
}
