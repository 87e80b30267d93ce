// Micro-benchmark of a warp-specialised structure for the step kernel: per CTA, L "logic" warps each spin `delay`
// cycles per env (the serial env step) and hand the env to a group of S "stream" warps through a shared-memory
// FIFO; the stream group writes the env's 100800-byte block cooperatively.  Persistent CTAs.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr long long kEnvBytes = 100800;
constexpr int kVecPerEnv = kEnvBytes / 16;
constexpr int kQ = 64;
__device__ __forceinline__ void spin(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
struct Ctl {
    int tail;            // next ticket
    int producers;       // logic warps still running
    int cur_env_lo, cur_env_hi;
    volatile int ready[kQ];
    int env_of[kQ][2];
};
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void ws(uint4* out, long long B, int L, int S, long long delay) {
    extern __shared__ uint4 smraw[];
    Ctl* c = reinterpret_cast<Ctl*>(smraw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        c->tail = 0; c->producers = L;
        for (int i = 0; i < kQ; ++i) c->ready[i] = 0;
    }
    __syncthreads();
    if (warp < L) {
        // logic warp: envs blockIdx.x*L + warp + k * gridDim.x * L
        for (long long env = (long long)blockIdx.x * L + warp; env < B; env += (long long)gridDim.x * L) {
            spin(delay);
            if (lane == 0) {
                const int t = atomicAdd(&c->tail, 1) % kQ;
                while (c->ready[t] != 0) __nanosleep(50);   // queue slot still in use
                c->env_of[t][0] = (int)(env & 0xFFFFFFFF); c->env_of[t][1] = (int)(env >> 32);
                __threadfence_block();
                c->ready[t] = 1;
            }
            __syncwarp();
        }
        if (lane == 0) atomicSub(&c->producers, 1);
    } else {
        // stream group
        const int st = threadIdx.x - L * 32, nst = S * 32;
        int head = 0;
        uint4 v = make_uint4(st, 5, 0x3F800000u, 0);
        for (;;) {
            if (st == 0) {
                int got = 0;
                for (;;) {
                    if (c->ready[head % kQ] == 1) { got = 1; break; }
                    if (*(volatile int*)&c->producers == 0 && *(volatile int*)&c->tail == head) break;
                    __nanosleep(50);
                }
                c->cur_env_lo = got ? c->env_of[head % kQ][0] : -1;
                c->cur_env_hi = got ? c->env_of[head % kQ][1] : -1;
                __threadfence_block();
            }
            named_bar(1, nst);
            const int lo = *(volatile int*)&c->cur_env_lo, hi = *(volatile int*)&c->cur_env_hi;
            if (lo == -1 && hi == -1) break;
            const long long env = ((long long)hi << 32) | (unsigned)lo;
            uint4* p = out + env * kVecPerEnv;
#pragma unroll 4
            for (int i = st; i < kVecPerEnv; i += nst) p[i] = v;
            named_bar(1, nst);
            if (st == 0) { c->ready[head % kQ] = 0; }
            ++head;
        }
    }
}
template <typename F>
float timeit(F f, int reps = 6) {
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
    cudaFuncSetAttribute(ws, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (long long delay : {0LL, 30000LL, 60000LL}) {
        for (int ctas_per_sm : {1, 2}) {
            for (int S : {4, 8, 16}) {
                for (int L : {8, 16, 24}) {
                    if ((S + L) * 32 > 1024) continue;
                    if ((S + L) * ctas_per_sm > 64) continue;
                    const size_t smem = (size_t)(220 * 1024) / ctas_per_sm - 2048;
                    const unsigned grid = 148 * ctas_per_sm;
                    float t = timeit([&] { ws<<<grid, (S + L) * 32, smem>>>(out, B, L, S, delay); });
                    cudaError_t e = cudaGetLastError();
                    printf("ws ctas/SM=%d S=%2d L=%2d delay=%6lld  %.4f ms %6.0f GB/s %s\n", ctas_per_sm, S, L, delay, t, gb / t * 1e3,
                           e == cudaSuccess ? "" : cudaGetErrorString(e));
                    fflush(stdout);
                }
            }
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
