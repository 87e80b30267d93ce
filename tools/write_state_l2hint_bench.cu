// Can L2 eviction-priority hints keep the 22 MB of env state resident across a step's 6.6 GB write stream?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr long long kEnvBytes = 100800;
constexpr int kVecPerEnv = kEnvBytes / 16;
constexpr int kStateVec = 21;
__device__ __forceinline__ void spin(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
__device__ __forceinline__ uint4 ld_evict_last(const uint4* p) {
    uint4 v; uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_evict_last(uint4* p, uint4 v) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_evict_first(uint4* p, uint4 v) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
// MODE 0: plain state r/w + plain stream. 1: state ld/st evict_last + plain stream. 2: state evict_last + stream evict_first. 3: no state
template <int MODE>
__global__ void k(uint4* out, uint4* state, long long B, int warps, long long delay) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
    if (MODE != 3) {
        uint4 s = make_uint4(0, 0, 0, 0);
        if (lane < kStateVec) s = (MODE == 0) ? state[env * kStateVec + lane] : ld_evict_last(state + env * kStateVec + lane);
        v.w = s.x & 1u;
        spin(delay + (s.y & 1u));
        if (lane < kStateVec) { s.x += 1; if (MODE == 0) state[env * kStateVec + lane] = s; else st_evict_last(state + env * kStateVec + lane, s); }
    } else spin(delay);
    uint4* p = out + env * kVecPerEnv;
#pragma unroll 4
    for (int i = lane; i < kVecPerEnv; i += 32) { if (MODE == 2) st_evict_first(p + i, v); else p[i] = v; }
}
template <typename F>
float timeit(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
int main() {
    const long long B = 65536, n = B * kVecPerEnv;
    uint4 *out, *state;
    cudaMalloc(&out, n * 16); cudaMalloc(&state, B * kStateVec * 16); cudaMemset(state, 0, B * kStateVec * 16);
    const double gb = n * 16 / 1e9;
    const int warps = 4; const unsigned grid = (unsigned)((B + warps - 1) / warps);
    const size_t smem = 220 * 1024 / 9 - 2048;  // 9 CTAs/SM like the real kernel
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (long long delay : {0LL, 30000LL}) {
        float t3 = timeit([&] { k<3><<<grid, warps * 32, smem>>>(out, state, B, warps, delay); });
        float t0 = timeit([&] { k<0><<<grid, warps * 32, smem>>>(out, state, B, warps, delay); });
        float t1 = timeit([&] { k<1><<<grid, warps * 32, smem>>>(out, state, B, warps, delay); });
        float t2 = timeit([&] { k<2><<<grid, warps * 32, smem>>>(out, state, B, warps, delay); });
        printf("delay=%6lld  no state %.4f | plain state %.4f | state evict_last %.4f | + stream evict_first %.4f   (GB/s %.0f %.0f %.0f %.0f)\n", delay, t3, t0, t1,
               t2, gb / t3 * 1e3, gb / t0 * 1e3, gb / t1 * 1e3, gb / t2 * 1e3);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
