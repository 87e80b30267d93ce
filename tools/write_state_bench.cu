// Micro-benchmark: does the small per-env state read (336 B) + write-back cost DRAM efficiency of the 100800-B
// store stream, and does an L2 persisting access-policy window on the state buffer recover it?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr long long kEnvBytes = 100800;
constexpr int kVecPerEnv = kEnvBytes / 16;
constexpr int kStateVec = 21;  // 336 B per env
__device__ __forceinline__ void spin(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
template <int MODE>   // 0: no state, 1: read state at start + write state before the stream
__global__ void k(uint4* out, uint4* state, long long B, int warps, long long delay) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
    if (MODE == 1) {
        uint4 s = make_uint4(0, 0, 0, 0);
        if (lane < kStateVec) s = state[env * kStateVec + lane];
        v.w = s.x & 1u;                     // the stream depends on the state (as in the real kernel)
        spin(delay + (s.y & 1u));
        if (lane < kStateVec) { s.x += 1; state[env * kStateVec + lane] = s; }
    } else {
        spin(delay);
    }
    uint4* p = out + env * kVecPerEnv;
#pragma unroll 4
    for (int i = lane; i < kVecPerEnv; i += 32) p[i] = v;
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
    cudaMalloc(&out, n * 16);
    cudaMalloc(&state, B * kStateVec * 16);
    cudaMemset(state, 0, B * kStateVec * 16);
    const double gb = n * 16 / 1e9;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("L2 %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB\n", prop.l2CacheSize >> 20,
           prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
    cudaStream_t s; cudaStreamCreate(&s);
    const int warps = 4;
    const unsigned grid = (unsigned)((B + warps - 1) / warps);
    for (int persist : {0, 1}) {
        if (persist) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 32 << 20);
            cudaStreamAttrValue attr = {};
            attr.accessPolicyWindow.base_ptr = state;
            attr.accessPolicyWindow.num_bytes = B * kStateVec * 16;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            printf("set window: %s\n", cudaGetErrorString(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr)));
        }
        for (long long delay : {0LL, 30000LL}) {
            float t0 = timeit([&] { k<0><<<grid, warps * 32, 0, s>>>(out, state, B, warps, delay); });
            float t1 = timeit([&] { k<1><<<grid, warps * 32, 0, s>>>(out, state, B, warps, delay); });
            printf("persist=%d delay=%6lld  no-state %.4f ms %6.0f GB/s | with state r/w %.4f ms %6.0f GB/s\n", persist, delay, t0, gb / t0 * 1e3, t1,
                   gb / t1 * 1e3);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
