// Write-pattern micro-benchmark: how fast can 65536 x 100800 B be written with different work->address mappings?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
constexpr long long kEnvBytes = 100800;   // 8_arena fp32 obs per env
constexpr int kVecPerEnv = kEnvBytes / 16; // 6300
template <int MODE, int STORE>
__device__ __forceinline__ void st(uint4* p, uint4 v) {
    if (STORE == 0) __stcs(p, v); else *p = v;
}
// MODE 0: one warp streams one env (our kernel's pattern). WARPS warps per CTA.
template <int STORE>
__global__ void warp_per_env(uint4* out, long long B, int warps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    uint4* p = out + env * kVecPerEnv;
    uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
#pragma unroll 4
    for (int i = lane; i < kVecPerEnv; i += 32) st<0, STORE>(p + i, v);
}
// MODE 1: the CTA streams its envs one after the other, all warps on the same env
template <int STORE>
__global__ void cta_per_env(uint4* out, long long B, int envs_per_cta) {
    for (int e = 0; e < envs_per_cta; ++e) {
        const long long env = (long long)blockIdx.x * envs_per_cta + e;
        if (env >= B) return;
        uint4* p = out + env * kVecPerEnv;
        uint4 v = make_uint4(threadIdx.x, e, 0x3F800000u, 0);
#pragma unroll 4
        for (int i = threadIdx.x; i < kVecPerEnv; i += blockDim.x) st<1, STORE>(p + i, v);
    }
}
// MODE 2: grid-stride over the whole buffer (what torch's fill does)
template <int STORE>
__global__ void grid_stride(uint4* out, long long n) {
    uint4 v = make_uint4(threadIdx.x, 1, 0x3F800000u, 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) st<2, STORE>(out + i, v);
}
// MODE 3: persistent warps: each warp takes envs round-robin (env = warp_global + k * total_warps)
template <int STORE>
__global__ void persistent_warp(uint4* out, long long B) {
    const int lane = threadIdx.x & 31;
    const long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, total = ((long long)gridDim.x * blockDim.x) >> 5;
    uint4 v = make_uint4(lane, 3, 0x3F800000u, 0);
    for (long long env = wg; env < B; env += total) {
        uint4* p = out + env * kVecPerEnv;
#pragma unroll 4
        for (int i = lane; i < kVecPerEnv; i += 32) st<3, STORE>(p + i, v);
    }
}
template <typename F>
float timeit(F f, int reps = 10) {
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
    auto report = [&](const char* name, float ms) { printf("%-44s %.4f ms  %.0f GB/s\n", name, ms, gb / ms * 1e3); };
    for (int warps : {1, 2, 4, 8, 16}) {
        char nm[64];
        snprintf(nm, 64, "warp_per_env cs   warps/cta=%d", warps);
        report(nm, timeit([&] { warp_per_env<0><<<(unsigned)((B + warps - 1) / warps), warps * 32>>>(out, B, warps); }));
        snprintf(nm, 64, "warp_per_env wb   warps/cta=%d", warps);
        report(nm, timeit([&] { warp_per_env<1><<<(unsigned)((B + warps - 1) / warps), warps * 32>>>(out, B, warps); }));
    }
    for (int threads : {128, 256, 512}) {
        for (int epc : {1, 4}) {
            char nm[64];
            snprintf(nm, 64, "cta_per_env wb threads=%d envs/cta=%d", threads, epc);
            report(nm, timeit([&] { cta_per_env<1><<<(unsigned)((B + epc - 1) / epc), threads>>>(out, B, epc); }));
        }
    }
    for (int mult : {2, 4, 8, 16}) {
        char nm[64];
        snprintf(nm, 64, "grid_stride wb ctas=148*%d x 256", mult);
        report(nm, timeit([&] { grid_stride<1><<<148 * mult, 256>>>(out, n); }));
        snprintf(nm, 64, "persistent_warp wb ctas=148*%d x 128", mult);
        report(nm, timeit([&] { persistent_warp<1><<<148 * mult, 128>>>(out, B); }));
    }
    report("cudaMemset", timeit([&] { cudaMemsetAsync(out, 0, n * 16); }));
    return 0;
}
