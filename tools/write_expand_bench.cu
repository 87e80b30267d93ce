// Micro-benchmark: cost of the bit->float expansion inside the warp-per-env store stream.
//   mode 0: store a constant            mode 1: LDS word + (b & 2^k) * K expansion (the kernel's code)
//   mode 2: LDS word + 16-entry float4 LUT in shared memory (LDS.128)   mode 3: like 1 but 8x unrolled
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr long long kEnvBytes = 100800;
constexpr int kVecPerEnv = kEnvBytes / 16;
constexpr int kWordsPerEnv = (kVecPerEnv + 7) / 8 + 4;
__device__ __forceinline__ void spin(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
__device__ __forceinline__ uint4 expand(uint32_t b) {
    uint4 v; v.x = (b & 1u) * 0x3F800000u; v.y = (b & 2u) * 0x1FC00000u; v.z = (b & 4u) * 0x0FE00000u; v.w = (b & 8u) * 0x07F00000u; return v;
}
template <int MODE>
__global__ void k(uint4* out, long long B, int warps, long long delay) {
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* bits = reinterpret_cast<uint32_t*>(sm) + warp * ((kWordsPerEnv + 3) & ~3);
    uint4* lut = sm + (warps * ((kWordsPerEnv + 3) & ~3)) / 4 + 1;
    if (threadIdx.x < 16) lut[threadIdx.x] = expand(threadIdx.x);
    for (int i = lane; i < kWordsPerEnv; i += 32) bits[i] = (i * 2654435761u) & 0x01010101u;  // sparse bits
    __syncthreads();
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    spin(delay);
    uint4* p = out + env * kVecPerEnv;
    const int sh = (lane & 7) * 4;
    const uint32_t* wp = bits + (lane >> 3);
    if (MODE == 0) {
        uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
#pragma unroll 4
        for (int i = lane; i < kVecPerEnv; i += 32) p[i] = v;
    } else if (MODE == 1) {
#pragma unroll 4
        for (int i = lane; i < kVecPerEnv; i += 32, wp += 4) p[i] = expand(*wp >> sh);
    } else if (MODE == 2) {
#pragma unroll 4
        for (int i = lane; i < kVecPerEnv; i += 32, wp += 4) p[i] = lut[(*wp >> sh) & 15u];
    } else {
#pragma unroll 8
        for (int i = lane; i < kVecPerEnv; i += 32, wp += 4) p[i] = expand(*wp >> sh);
    }
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
template <int MODE>
void run(uint4* out, long long B, int ctas, long long delay, double gb) {
    const int warps = 4;
    const size_t smem = (size_t)(220 * 1024) / ctas - 2048;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const unsigned grid = (unsigned)((B + warps - 1) / warps);
    float t = timeit([&] { k<MODE><<<grid, warps * 32, smem>>>(out, B, warps, delay); });
    printf("mode %d  warps/SM=%2d delay=%6lld  %.4f ms %6.0f GB/s\n", MODE, warps * ctas, delay, t, gb / t * 1e3);
}
int main() {
    const long long B = 65536, n = B * kVecPerEnv;
    uint4* out; cudaMalloc(&out, n * 16);
    const double gb = n * 16 / 1e9;
    for (long long delay : {0LL, 30000LL})
        for (int ctas : {6, 9, 12}) {
            run<0>(out, B, ctas, delay, gb); run<1>(out, B, ctas, delay, gb); run<2>(out, B, ctas, delay, gb); run<3>(out, B, ctas, delay, gb);
        }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
