// Micro-benchmark: warp-private streaming of 100800-byte env blocks through shared-memory staging + TMA bulk stores
// (cp.async.bulk.global.shared::cta) versus direct st.global.v4, with an emulated serial logic phase per env.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr long long kEnvBytes = 100800;
constexpr int kVecPerEnv = kEnvBytes / 16;
__device__ __forceinline__ void spin(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// direct stores (the current kernel's pattern)
__global__ void direct(uint4* out, long long B, int warps, long long delay) {
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
// staged: each warp owns 2 x CH bytes of shared memory; fill with 128-bit STS, then one lane issues a bulk store
template <int CH>
__global__ void staged(uint4* out, long long B, int warps, long long delay, int stage_off_vec) {
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    spin(delay);
    constexpr int VEC = CH / 16;
    uint4* buf = sm + stage_off_vec + warp * 2 * VEC;
    uint4* p = out + env * kVecPerEnv;
    uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
    int k = 0;
    for (int base = 0; base < kVecPerEnv; base += VEC, ++k) {
        uint4* b = buf + (k & 1) * VEC;
        const int n = min(VEC, kVecPerEnv - base);
        if (lane == 0) bulk_wait_read<1>();   // the buffer used two chunks ago has been read
        __syncwarp();
        for (int i = lane; i < n; i += 32) b[i] = v;
        fence_async();
        __syncwarp();
        if (lane == 0) { bulk_store(p + base, b, (uint32_t)n * 16); bulk_commit(); }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
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
template <int CH>
void run_staged(uint4* out, long long B, int warps, int ctas, long long delay, double gb) {
    const size_t smem = (size_t)(220 * 1024) / ctas - 2048;
    if ((size_t)warps * 2 * CH > smem) return;
    cudaFuncSetAttribute(staged<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const unsigned grid = (unsigned)((B + warps - 1) / warps);
    float t = timeit([&] { staged<CH><<<grid, warps * 32, smem>>>(out, B, warps, delay, 0); });
    printf("staged CH=%5d  warps/cta=%d ctas/SM=%2d warps/SM=%2d delay=%6lld  %.4f ms %6.0f GB/s\n", CH, warps, ctas, warps * ctas, delay, t, gb / t * 1e3);
}
int main() {
    const long long B = 65536, n = B * kVecPerEnv;
    uint4* out; cudaMalloc(&out, n * 16);
    const double gb = n * 16 / 1e9;
    cudaFuncSetAttribute(direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (long long delay : {0LL, 30000LL}) {
        for (int ctas : {4, 6, 8, 10}) {
            const int warps = 4;
            const size_t smem = (size_t)(220 * 1024) / ctas - 2048;
            const unsigned grid = (unsigned)((B + warps - 1) / warps);
            float t = timeit([&] { direct<<<grid, warps * 32, smem>>>(out, B, warps, delay); });
            printf("direct          warps/cta=%d ctas/SM=%2d warps/SM=%2d delay=%6lld  %.4f ms %6.0f GB/s\n", warps, ctas, warps * ctas, delay, t, gb / t * 1e3);
            run_staged<1024>(out, B, warps, ctas, delay, gb);
            run_staged<2048>(out, B, warps, ctas, delay, gb);
            run_staged<4096>(out, B, warps, ctas, delay, gb);
            run_staged<8192>(out, B, warps, ctas, delay, gb);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
