// Emulates the step kernel's structure: a serial "logic" delay per env followed by streaming 100800 B,
// with warp-private streaming (A) or CTA-cooperative streaming (B).  Finds the best structure for HBM writes.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr long long kEnvBytes = 100800;
constexpr int kVecPerEnv = kEnvBytes / 16;
__device__ __forceinline__ void spin(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
// A: each warp: delay, then streams its own env
__global__ void patA(uint4* out, long long B, int warps, long long delay) {
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    spin(delay);
    uint4* p = out + env * kVecPerEnv;
    uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
#pragma unroll 4
    for (int i = lane; i < kVecPerEnv; i += 32) p[i] = v;
}
// B: each warp: delay (its env's logic); __syncthreads; all warps stream the CTA's envs one after the other
__global__ void patB(uint4* out, long long B, int warps, long long delay) {
    extern __shared__ uint4 sm[];
    const long long env0 = (long long)blockIdx.x * warps;
    spin(delay);
    __syncthreads();
    uint4 v = make_uint4(threadIdx.x, 7, 0x3F800000u, 0);
    for (int e = 0; e < warps; ++e) {
        const long long env = env0 + e;
        if (env >= B) break;
        uint4* p = out + env * kVecPerEnv;
#pragma unroll 4
        for (int i = threadIdx.x; i < kVecPerEnv; i += blockDim.x) p[i] = v;
    }
}
// C: like B but the CTA's envs are streamed as one contiguous region (they are adjacent in memory)
__global__ void patC(uint4* out, long long B, int warps, long long delay) {
    extern __shared__ uint4 sm[];
    const long long env0 = (long long)blockIdx.x * warps;
    spin(delay);
    __syncthreads();
    uint4 v = make_uint4(threadIdx.x, 9, 0x3F800000u, 0);
    const long long n_env = (B - env0) < warps ? (B - env0) : warps;
    uint4* p = out + env0 * kVecPerEnv;
    const int n = (int)(n_env * kVecPerEnv);
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = v;
}
template <typename F>
float timeit(F f, int reps = 8) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
int main() {
    const long long B = 65536, n = B * kVecPerEnv;
    uint4* out; cudaMalloc(&out, n * 16);
    const double gb = n * 16 / 1e9;
    cudaFuncSetAttribute(patA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(patB, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(patC, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("%-8s %6s %6s %8s %8s  %9s %8s\n", "pattern", "warps", "ctas/SM", "warps/SM", "delay", "ms", "GB/s");
    for (long long delay : {0LL, 20000LL, 40000LL}) {
        for (int warps : {4, 8, 16}) {
            for (int warps_per_sm : {16, 32, 48, 64}) {
                const int ctas = warps_per_sm / warps;
                if (ctas < 1) continue;
                const size_t smem = (size_t)(220 * 1024) / ctas - 2048;  // forces exactly `ctas` CTAs per SM
                const unsigned grid = (unsigned)((B + warps - 1) / warps);
                float a = timeit([&] { patA<<<grid, warps * 32, smem>>>(out, B, warps, delay); });
                float b = timeit([&] { patB<<<grid, warps * 32, smem>>>(out, B, warps, delay); });
                float c = timeit([&] { patC<<<grid, warps * 32, smem>>>(out, B, warps, delay); });
                printf("A        %6d %6d %8d %8lld  %9.4f %8.0f\n", warps, ctas, warps_per_sm, delay, a, gb / a * 1e3);
                printf("B        %6d %6d %8d %8lld  %9.4f %8.0f\n", warps, ctas, warps_per_sm, delay, b, gb / b * 1e3);
                printf("C        %6d %6d %8d %8lld  %9.4f %8.0f\n", warps, ctas, warps_per_sm, delay, c, gb / c * 1e3);
            }
        }
    }
    return 0;
}
