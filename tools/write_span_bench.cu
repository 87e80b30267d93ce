// Round-2 write-stream micro-benchmark: what bounds the 6.6 GB observation write of one 8_arena step?
//
// Round 1 found that the store stream gets faster the fewer env blocks are "open" at the same time
// (60 MB of open blocks: 0.884 ms, 540 MB: 0.94 ms, whole buffer: 1.05 ms).  This bench separates the candidate
// causes and measures the two structures that keep the open set small while env logic runs beside the stream:
//   T0  calibration: cudaMemset, warp-per-env (the round-1 kernel's pattern), 512-thread CTA per env
//   T1  512-thread CTA per env, envs visited in a scattered order (few wide streams, large address span)
//   T2  "token": persistent CTAs, envs fetched in order from a global counter, every warp = spin(delay) logic, then
//       streams its env alone, but only K warps per CTA may stream at a time (ticket semaphore in shared memory)
//   T3  "group": persistent CTAs, L logic warps (fetch env, spin) hand envs through a FIFO to one group of S stream
//       warps that write one env block at a time cooperatively
//   T4  per-SM store bandwidth (are some SMs slower? static work partitioning would then be bound by the slowest)
// Stream loops read a bit string from shared memory (one LDS.128 per four 128-bit stores) and expand bits to floats,
// like the real kernel.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

constexpr long long kEnvBytes = 100800;    // 8_arena fp32 observations per env
constexpr int kVecPerEnv = kEnvBytes / 16;  // 6300
constexpr int kBitsWords = 800;             // 788 used, padded to a multiple of 16

__device__ __forceinline__ void spin(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
__device__ __forceinline__ uint4 expand(uint32_t b) {
    uint4 v;
    v.x = (b & 1u) * 0x3F800000u; v.y = (b & 2u) * 0x1FC00000u; v.z = (b & 4u) * 0x0FE00000u; v.w = (b & 8u) * 0x07F00000u;
    return v;
}
// group j of 128 vectors (2 KB): lane (q = lane >> 3, n = lane & 7) reads words 16j + 4q .. +3 with one LDS.128 and
// stores vector 8 * (16j + 4q + m) + n for m = 0..3 — every store instruction writes four complete 128-byte lines
__device__ __forceinline__ void stream_group(const uint32_t* bits, uint4* __restrict__ p, int j, int lane) {
    const int q = lane >> 3, n = lane & 7, sh = n * 4;
    const uint4 w = *reinterpret_cast<const uint4*>(bits + 16 * j + 4 * q);
    const int v0 = 8 * (16 * j + 4 * q) + n;
    if (v0 < kVecPerEnv) p[v0] = expand(w.x >> sh);
    if (v0 + 8 < kVecPerEnv) p[v0 + 8] = expand(w.y >> sh);
    if (v0 + 16 < kVecPerEnv) p[v0 + 16] = expand(w.z >> sh);
    if (v0 + 24 < kVecPerEnv) p[v0 + 24] = expand(w.w >> sh);
}
constexpr int kGroups = (kVecPerEnv + 127) / 128;  // 50

// ---------------------------------------------------------------- T0 / T1
__global__ void warp_per_env(uint4* out, long long B, int warps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * warps + warp;
    if (env >= B) return;
    uint4* p = out + env * kVecPerEnv;
    uint4 v = make_uint4(lane, warp, 0x3F800000u, 0);
#pragma unroll 4
    for (int i = lane; i < kVecPerEnv; i += 32) p[i] = v;
}
__global__ void cta_per_env(uint4* out, long long B, long long mul) {
    const long long env = mul ? (long long)(((unsigned long long)blockIdx.x * (unsigned long long)mul) % (unsigned long long)B) : blockIdx.x;
    uint4* p = out + env * kVecPerEnv;
    uint4 v = make_uint4(threadIdx.x, 1, 0x3F800000u, 0);
#pragma unroll 4
    for (int i = threadIdx.x; i < kVecPerEnv; i += blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------- T2: token
__global__ void tok(uint4* out, int B, int K, long long delay, int* ctr, int state_rw, uint4* state) {
    extern __shared__ uint4 smraw[];
    __shared__ int tail, head;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* bits = reinterpret_cast<uint32_t*>(smraw) + warp * kBitsWords;
    for (int i = lane; i < kBitsWords; i += 32) bits[i] = 0x01020408u * (i + 1);
    if (threadIdx.x == 0) { tail = 0; head = 0; }
    __syncthreads();
    for (;;) {
        int env = 0;
        if (lane == 0) env = atomicAdd(ctr, 1);
        env = __shfl_sync(0xFFFFFFFFu, env, 0);
        if (env >= B) break;
        uint4 st = make_uint4(0, 0, 0, 0);
        if (state_rw && lane < 21) st = state[(long long)env * 21 + lane];   // 336 B of env state
        spin(delay);
        if (state_rw && lane < 21) { st.x += 1; state[(long long)env * 21 + lane] = st; }
        bits[lane] ^= (uint32_t)env;
        if (lane == 0) {
            const int t = atomicAdd(&tail, 1);
            while (t - *(volatile int*)&head >= K) __nanosleep(64);
        }
        __syncwarp();
        uint4* p = out + (long long)env * kVecPerEnv;
#pragma unroll 2
        for (int j = 0; j < kGroups; ++j) stream_group(bits, p, j, lane);
        __syncwarp();
        if (lane == 0) atomicAdd(&head, 1);
    }
}

// ---------------------------------------------------------------- T3: logic warps -> FIFO -> one stream group
constexpr int kQ = 64;
struct Ctl {
    int tail;
    int producers;
    int cur;                      // FIFO entry the stream group is working on (-1: finished)
    volatile int ready[kQ];       // 0 free, 1 filled
    int env_of[kQ];
    int buf_of[kQ];
    volatile int busy[32][2];     // logic warp's two bit-string buffers: still to be streamed?
};
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void grp(uint4* out, int B, int L, int S, long long delay, int* ctr) {
    extern __shared__ uint4 smraw[];
    Ctl* c = reinterpret_cast<Ctl*>(smraw);
    uint32_t* bits0 = reinterpret_cast<uint32_t*>(smraw) + 1024;  // after the control block (4 KB)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        c->tail = 0; c->producers = L; c->cur = 0;
        for (int i = 0; i < kQ; ++i) c->ready[i] = 0;
        for (int i = 0; i < 32; ++i) { c->busy[i][0] = 0; c->busy[i][1] = 0; }
    }
    for (int i = threadIdx.x; i < L * 2 * kBitsWords; i += blockDim.x) bits0[i] = 0x01020408u * (i + 1);
    __syncthreads();
    if (warp < L) {
        int n = 0;
        for (;;) {
            int env = 0;
            if (lane == 0) env = atomicAdd(ctr, 1);
            env = __shfl_sync(0xFFFFFFFFu, env, 0);
            if (env >= B) break;
            const int par = n & 1;
            if (lane == 0) while (c->busy[warp][par]) __nanosleep(64);   // buffer of two envs ago still being streamed
            __syncwarp();
            spin(delay);
            bits0[(warp * 2 + par) * kBitsWords + lane] ^= (uint32_t)env;
            __syncwarp();
            if (lane == 0) {
                c->busy[warp][par] = 1;
                const int t = atomicAdd(&c->tail, 1) % kQ;
                while (c->ready[t] != 0) __nanosleep(64);
                c->env_of[t] = env; c->buf_of[t] = warp * 2 + par;
                __threadfence_block();
                c->ready[t] = 1;
            }
            __syncwarp();
            ++n;
        }
        if (lane == 0) atomicSub(&c->producers, 1);
    } else {
        const int st = threadIdx.x - L * 32, nst = S * 32, sw = st >> 5;
        int head = 0;
        for (;;) {
            if (st == 0) {
                int got = 0;
                for (;;) {
                    if (c->ready[head % kQ] == 1) { got = 1; break; }
                    if (*(volatile int*)&c->producers == 0 && *(volatile int*)&c->tail == head) break;
                    __nanosleep(64);
                }
                c->cur = got ? head % kQ : -1;
                __threadfence_block();
            }
            named_bar(1, nst);
            const int cur = *(volatile int*)&c->cur;
            if (cur < 0) break;
            const int env = c->env_of[cur], buf = c->buf_of[cur];
            const uint32_t* bits = bits0 + buf * kBitsWords;
            uint4* p = out + (long long)env * kVecPerEnv;
            for (int j = sw; j < kGroups; j += S) stream_group(bits, p, j, lane);
            named_bar(1, nst);
            if (st == 0) { c->busy[buf >> 1][buf & 1] = 0; c->ready[cur] = 0; }
            ++head;
        }
    }
}

// ---------------------------------------------------------------- T4: per-SM store bandwidth
__global__ void per_sm(uint4* out, long long vec_per_cta, long long* cycles, int* smid) {
    uint4* p = out + (long long)blockIdx.x * vec_per_cta;
    uint4 v = make_uint4(threadIdx.x, 2, 0x3F800000u, 0);
    __syncthreads();
    const long long t0 = clock64();
    for (long long i = threadIdx.x; i < vec_per_cta; i += blockDim.x) p[i] = v;
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
        cycles[blockIdx.x] = t1 - t0;
        int id; asm("mov.u32 %0, %%smid;" : "=r"(id));
        smid[blockIdx.x] = id;
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

int main(int argc, char** argv) {
    const int B = 65536;
    const long long n = (long long)B * kVecPerEnv;
    uint4* out; cudaMalloc(&out, n * 16);
    int* ctr; cudaMalloc(&ctr, 4);
    uint4* state; cudaMalloc(&state, (size_t)B * 21 * 16); cudaMemset(state, 0, (size_t)B * 21 * 16);
    const double gb = n * 16 / 1e9;
    auto report = [&](const char* name, float ms) { printf("%-64s %.4f ms  %5.0f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
    char nm[128];

    report("T0 cudaMemset", timeit([&] { cudaMemsetAsync(out, 0, n * 16); }));
    report("T0 warp_per_env 4 warps/CTA (round-1 pattern, no logic)", timeit([&] { warp_per_env<<<B / 4, 128>>>(out, B, 4); }));
    report("T0 cta_per_env 512 threads, in order", timeit([&] { cta_per_env<<<B, 512>>>(out, B, 0); }));
    report("T0 cta_per_env 256 threads, in order", timeit([&] { cta_per_env<<<B, 256>>>(out, B, 0); }));
    report("T1 cta_per_env 512 threads, scattered (env = cta*4099 % B)", timeit([&] { cta_per_env<<<B, 512>>>(out, B, 4099); }));
    report("T1 cta_per_env 512 threads, scattered (env = cta*32771 % B)", timeit([&] { cta_per_env<<<B, 512>>>(out, B, 32771); }));
    report("T1 cta_per_env 256 threads, scattered (env = cta*4099 % B)", timeit([&] { cta_per_env<<<B, 256>>>(out, B, 4099); }));

    cudaFuncSetAttribute(tok, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(grp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (long long delay : {0LL, 20000LL, 40000LL}) {
        for (int ctas_per_sm : {1, 2}) {
            for (int W : {16, 24, 32}) {
                if (W * ctas_per_sm > 64 || W * 32 > 1024) continue;
                if (W * ctas_per_sm < 24 && delay > 0) continue;
                for (int K : {1, 2, 3, 4, 6, 8, 64}) {
                    if (K > W && K != 64) continue;
                    if (delay == 0 && K != 2 && K != 4 && K != 64) continue;
                    const size_t smem = (size_t)W * kBitsWords * 4;
                    if (smem * ctas_per_sm > 200 * 1024) continue;
                    float t = timeit([&] { cudaMemsetAsync(ctr, 0, 4); tok<<<148 * ctas_per_sm, W * 32, smem>>>(out, B, K, delay, ctr, 0, state); });
                    snprintf(nm, sizeof nm, "T2 token  ctas/SM=%d W=%2d K=%2d delay=%5lld", ctas_per_sm, W, K, delay);
                    report(nm, t);
                }
            }
        }
    }
    // with the 336-byte state read + write-back per env, best-looking shapes
    for (int K : {2, 4, 64}) {
        const int W = 32; const size_t smem = (size_t)W * kBitsWords * 4;
        float t = timeit([&] { cudaMemsetAsync(ctr, 0, 4); tok<<<148, W * 32, smem>>>(out, B, K, 20000, ctr, 1, state); });
        snprintf(nm, sizeof nm, "T2 token+state r/w ctas/SM=1 W=32 K=%2d delay=20000", K);
        report(nm, t);
    }
    for (long long delay : {0LL, 20000LL, 40000LL}) {
        for (int ctas_per_sm : {1, 2}) {
            for (int S : {2, 4, 8}) {
                for (int L : {6, 8, 12, 16, 24}) {
                    if ((S + L) * 32 > 1024 || (S + L) * ctas_per_sm > 64) continue;
                    if (delay == 0 && L != 8) continue;
                    const size_t smem = 4096 + (size_t)L * 2 * kBitsWords * 4;
                    if (smem * ctas_per_sm > 200 * 1024) continue;
                    float t = timeit([&] { cudaMemsetAsync(ctr, 0, 4); grp<<<148 * ctas_per_sm, (S + L) * 32, smem>>>(out, B, L, S, delay, ctr); });
                    snprintf(nm, sizeof nm, "T3 group  ctas/SM=%d S=%d L=%2d delay=%5lld", ctas_per_sm, S, L, delay);
                    report(nm, t);
                }
            }
        }
    }
    {   // T4: one 512-thread CTA per SM, 32 MB each
        const int ctas = 148; const long long vec = (32ll << 20) / 16;
        long long* cyc; int* smid; cudaMalloc(&cyc, ctas * 8); cudaMalloc(&smid, ctas * 4);
        for (int rep = 0; rep < 2; ++rep) per_sm<<<ctas, 512>>>(out, vec, cyc, smid);
        cudaDeviceSynchronize();
        std::vector<long long> h(ctas); std::vector<int> s(ctas);
        cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost); cudaMemcpy(s.data(), smid, ctas * 4, cudaMemcpyDeviceToHost);
        std::vector<long long> sorted = h; std::sort(sorted.begin(), sorted.end());
        printf("T4 per-SM cycles for 32 MB: min %lld  p10 %lld  median %lld  p90 %lld  max %lld  (max/min %.3f)\n", sorted[0], sorted[14], sorted[74],
               sorted[133], sorted[147], (double)sorted[147] / sorted[0]);
        printf("T4 slowest SMs:");
        for (int i = 0; i < ctas; ++i) if (h[i] >= sorted[140]) printf(" sm%d=%lld", s[i], h[i]);
        printf("\n");
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
