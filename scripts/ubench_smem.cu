// Standalone microbenchmark (round 2): what the counting kernels are made of on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_smem.bin scripts/ubench_smem.cu
// Measures, per SM (one CTA per SM, W warps), lane-operations per cycle of
//   a  atom.shared.add.u32 (with return)  on 1024 random counters        (ring heads of the scatter)
//   b  red.shared.add.u32                  on 16384 random slots          (count deltas)
//   c  atom.shared.cas.b64                 on 16384 random slots
//   d  ld.shared.u64                       on 16384 random slots          (probe)
//   e  __match_any_sync on a 10-bit value
//   f  st.shared::cluster.u64 to a random CTA of a cluster of 8 (remote push), and remote red.add.u32
//   g  cudaOccupancyMaxActiveClusters for clusters of 2/4/8/16 CTAs with ~200 KB of shared memory
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ u32 xs(u32 &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

constexpr int ITER = 2048;

template <int MODE>
__global__ void __launch_bounds__(1024) k_local(u64 *out, long long *cyc) {
    extern __shared__ __align__(16) u64 sm[];
    u32 *s32 = reinterpret_cast<u32 *>(sm);
    for (int i = threadIdx.x; i < 16384 * 2; i += blockDim.x) s32[i] = 0;
    __syncthreads();
    u32 seed = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    u32 acc = 0;
    const u32 base = (u32)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITER; it++) {
        const u32 r = xs(seed);
        if (MODE == 0) { u32 old; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(base + 4 * (r & 1023)) : "memory"); acc += old; }
        if (MODE == 1) { asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(base + 4 * (r & 16383)) : "memory"); }
        if (MODE == 2) { u64 old; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(base + 8 * (r & 16383)), "l"(0ull), "l"((u64)r) : "memory"); acc += (u32)old; }
        if (MODE == 3) { u64 v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(base + 8 * (r & 16383)) : "memory"); acc += (u32)v; }
        if (MODE == 4) { acc += __match_any_sync(0xFFFFFFFFu, r & 1023); }
        if (MODE == 5) { u64 v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(base + 8 * (r & 16383)) : "memory"); acc += (u32)v;
                         asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(base + 131072 + 4 * (r & 16383)) : "memory"); }
        if (MODE == 6) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(base + 8 * (r & 16383)), "l"((u64)r) : "memory"); }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 1024 + threadIdx.x] = acc + s32[threadIdx.x];
}

// remote push inside a cluster: every lane stores 8 bytes to (MODE 0) / atomically adds to (MODE 1) a random slot of a random CTA
template <int MODE>
__global__ void __launch_bounds__(1024) k_remote(u64 *out, long long *cyc, int csize) {
    extern __shared__ __align__(16) u64 sm[];
    u32 *s32 = reinterpret_cast<u32 *>(sm);
    for (int i = threadIdx.x; i < 4096 * 2; i += blockDim.x) s32[i] = 0;
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    u32 seed = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 999u;
    const u32 base = (u32)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITER; it++) {
        const u32 r = xs(seed);
        const u32 dst = (r >> 20) % (u32)csize;
        u32 ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + 8 * (r & 4095)), "r"(dst));
        if (MODE == 0) asm volatile("st.shared::cluster.u64 [%0], %1;" :: "r"(ra), "l"((u64)r) : "memory");
        if (MODE == 1) asm volatile("red.shared::cluster.add.u32 [%0], 1;" :: "r"(ra) : "memory");
        if (MODE == 2) { u64 v; asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(ra) : "memory"); seed += (u32)v; }
    }
    long long t1 = clock64();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 1024 + threadIdx.x] = seed + s32[threadIdx.x];
}

template <class K>
static int run_local(const char *name, K kern, int warps, u64 *out, long long *cyc, size_t smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<148, warps * 32, smem>>>(out, cyc);
    cudaEventRecord(e0);
    kern<<<148, warps * 32, smem>>>(out, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double ops = (double)warps * 32 * ITER;
    printf("%-28s warps=%2d  %.3f lane-ops/cycle/SM  (%.1f cyc per warp-instr)  %.3f ms  -> 1e9 ops on 148 SMs: %.2f ms\n", name, warps, ops / avg,
           avg / (warps * ITER) , ms, 1e9 / 148 / (ops / avg) / 1.9e6);
    return 0;
}

template <class K>
static int run_remote(const char *name, K kern, int warps, int csize, u64 *out, long long *cyc) {
    const size_t smem = 65536;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (csize > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    const int grid = (148 / csize) * csize;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(warps * 32); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaLaunchKernelEx(&cfg, kern, out, cyc, csize));
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, kern, out, cyc, csize));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    const double ops = (double)warps * 32 * ITER;
    printf("%-28s cluster=%2d warps=%2d  %.3f lane-ops/cycle/SM  %.3f ms (grid %d)\n", name, csize, warps, ops / avg, ms, grid);
    return 0;
}

__global__ void __launch_bounds__(1024) k_dummy(int *p) { extern __shared__ u64 sm[]; if (p) p[0] = (int)sm[threadIdx.x]; }

int main() {
    u64 *out; long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 8)); CK(cudaMalloc(&cyc, 148 * 8 * 2));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("%s  SMs %d  smem/SM %zu  smem/block optin %zu  clock %d kHz  L2 %d MB\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerMultiprocessor,
           pr.sharedMemPerBlockOptin, pr.clockRate, pr.l2CacheSize >> 20);
    const size_t smem = 200 * 1024;
    for (int w : {8, 16, 32}) {
        run_local("a atoms.add.u32 ret (1024)", k_local<0>, w, out, cyc, smem);
        run_local("b red.shared.add.u32 (16K)", k_local<1>, w, out, cyc, smem);
        run_local("c atoms.cas.b64 (16K)", k_local<2>, w, out, cyc, smem);
        run_local("d ld.shared.u64 (16K)", k_local<3>, w, out, cyc, smem);
        run_local("e match_any 10-bit", k_local<4>, w, out, cyc, smem);
        run_local("f lds.u64 + red.u32", k_local<5>, w, out, cyc, smem);
        run_local("g st.shared.u64 (16K)", k_local<6>, w, out, cyc, smem);
    }
    for (int cs : {2, 4, 8}) for (int w : {8, 32}) {
        run_remote("st.shared::cluster.u64", k_remote<0>, w, cs, out, cyc);
        run_remote("red.shared::cluster.add.u32", k_remote<1>, w, cs, out, cyc);
        run_remote("ld.shared::cluster.u64", k_remote<2>, w, cs, out, cyc);
    }
    for (int cs : {1, 2, 4, 8, 16}) for (int threads : {512, 1024}) {
        cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, k_dummy, &cfg);
        printf("max active clusters: cluster=%2d threads=%4d smem=200KB -> %d clusters = %d SMs (%s)\n", cs, threads, nc, nc * cs, cudaGetErrorString(e));
    }
    return 0;
}
