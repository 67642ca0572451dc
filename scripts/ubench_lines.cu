// Microbenchmark: does the granularity of scattered line writes limit a streaming kernel on B200?
// Traffic shape of pack32_kernel<scatter>: per 512 reads a CTA reads 20 KB (streaming), writes 4.5 KB coalesced and 4 KB as
// LINE-byte lines to 256 private streams (pseudo-random stream per line).  LINE = 0: the 4 KB go out coalesced too.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_lines.bin scripts/ubench_lines.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int T = 256, P = 256;

__device__ __forceinline__ uint4 ld_stream(const void *p) {
    uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}

template <int LINE>
__global__ void __launch_bounds__(T, 3) mix(const uint4 *in, uint4 *out_co, uint4 *out_sc, long iters_total, u32 seg_bytes) {
    __shared__ u32 cur[P];
    for (int p = threadIdx.x; p < P; p += T) cur[p] = 0;
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (long it = blockIdx.x; it < iters_total; it += gridDim.x) {
        const uint4 *src = in + it * 1280;                     // 20 KB = 1280 x 16 B
        uint4 v[5];
#pragma unroll
        for (int j = 0; j < 5; j++) v[j] = ld_stream(src + j * T + threadIdx.x);
#pragma unroll
        for (int j = 0; j < 5; j++) { acc.x ^= v[j].x; acc.y += v[j].y; acc.z ^= v[j].z; acc.w += v[j].w; }
        uint4 *co = out_co + it * 288;                          // 4.5 KB coalesced = 288 x 16 B
        co[threadIdx.x] = acc;
        if (threadIdx.x < 32) co[256 + threadIdx.x] = acc;
        if constexpr (LINE == 0) {
            out_sc[it * 256 + threadIdx.x] = acc;               // 4 KB coalesced instead
        } else {
            constexpr int LANES = LINE / 16, LINES = 4096 / LINE;      // lanes per line, lines per iteration
            const int line = threadIdx.x / LANES, sub = threadIdx.x % LANES;
            if (line < LINES) {
                const u32 part = (u32)((it * LINES + line) * 2654435761u) >> 24;   // pseudo-random stream
                u32 pos = 0;
                if (sub == 0) pos = atomicAdd(&cur[part], (u32)LINE);
                pos = __shfl_sync(0xFFFFFFFFu, pos, LANES >= 32 ? 0 : (threadIdx.x & 31) / LANES * LANES);
                if (pos + LINE <= seg_bytes)
                    *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(out_sc) + ((size_t)blockIdx.x * P + part) * seg_bytes + pos + 16 * sub) = acc;
            }
        }
    }
}

int main(int argc, char **argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 1000000;      // x 512 "reads"
    const int grid = 444;
    uint4 *in, *co, *sc;
    const u32 seg_bytes = (u32)(((iters / grid + 1) * 4096 / P) * 3 / 2 + 4096) & ~511u;
    CK(cudaMalloc(&in, iters * 20480)); CK(cudaMalloc(&co, iters * 4608));
    const size_t sc_bytes = (size_t)grid * P * seg_bytes > (size_t)iters * 4096 ? (size_t)grid * P * seg_bytes : (size_t)iters * 4096;
    CK(cudaMalloc(&sc, sc_bytes));
    CK(cudaMemset(in, 1, iters * 20480));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int line) -> float {
        float best = 1e9f;
        for (int r = 0; r < 4; r++) {
            cudaEventRecord(e0);
            if (line == 0) mix<0><<<grid, T>>>(in, co, sc, iters, seg_bytes);
            else if (line == 64) mix<64><<<grid, T>>>(in, co, sc, iters, seg_bytes);
            else if (line == 128) mix<128><<<grid, T>>>(in, co, sc, iters, seg_bytes);
            else if (line == 256) mix<256><<<grid, T>>>(in, co, sc, iters, seg_bytes);
            else mix<512><<<grid, T>>>(in, co, sc, iters, seg_bytes);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
        }
        return best;
    };
    const double bytes = (double)iters * (20480 + 4608 + 4096);
    for (int line : {0, 64, 128, 256, 512}) {
        float ms = run(line);
        if (cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
        printf("scattered line %3d B: %7.3f ms  %7.1f GB/s  (%.2f ms per 1e9 reads)\n", line, ms, bytes / ms / 1e6, ms * (1e9 / 512) / iters);
    }
    return 0;
}
