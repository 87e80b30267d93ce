// ctf_kernels.cu — sm_100a kernels and C ABI of the batched GridworldCtf step path.
//
// One warp owns one environment.  The env's tile map is staged in shared memory
// (16-byte-row layout, 256 B), agent i lives in the registers of lane i, and the
// reference's sequential semantics (gridworld_ctf.py:849-918) are kept by
// iterating the dice-ordered agents serially while using the lanes for the
// parallel inner parts: per-opponent tag tests (ballot), the 3x3 respawn window
// (ballot + nth-set-bit), proximity metrics (ballot + popc), state staging and
// the observation writer.  Observations (standardise_state, :975-1009) are built
// as one bit per output element in shared memory — only non-open cells set bits —
// and streamed to HBM with 128-bit stores, so the kernel's cost is the store
// stream, not the logic.
//
// Randomness is Philox4x32-10 addressed by (seed, global env id, episode, step,
// site); see marl_ctf_development_b200/draws.py for the site map.  Lane l of the
// warp generates site l.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>   // snprintf for error messages
#include <stdlib.h>  // getenv: kernel-shape overrides for A/B runs
#include <string.h>

#include <new>
#include <type_traits>

#include "../../include/ctf_b200.h"

namespace {

#ifndef CTF_WARPS_PER_CTA
#define CTF_WARPS_PER_CTA 4
#endif
#ifndef CTF_STORE_OP
// 0: st.global.cs (streaming), 1: default write-back, 2: st.global.wt.  Measured on B200 (profiles/r01_ab_store_op.log):
// plain write-back stores are 4 % faster than .cs for this stream (L2 merges and schedules the write-backs).
#define CTF_STORE_OP 1
#endif
#ifndef CTF_STATE_HINT
// 1: env state loads / stores carry an L2 evict_last policy (and, with CTF_STORE_OP 3, the observation stream an
// evict_first policy) so that the 22 MB of state written by step t are still in L2 when step t+1 reads them.
#define CTF_STATE_HINT 0
#endif
#ifndef CTF_DIAG
// TIMING DIAGNOSTICS ONLY (results are wrong when non-zero): bit 0 skip the state read, 1 skip the state write-back,
// 2 skip metadata, 3 skip rewards / dones, 4 skip the statistics reductions, 5 skip the actions read
#define CTF_DIAG 0
#endif
constexpr int kWarpsPerCta = CTF_WARPS_PER_CTA;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kRow = 16;       // shared/global grid row stride (cells)
constexpr int kGridBytes = 256;
constexpr int kMaxList = 256;  // non-open cell list entries
constexpr unsigned kFull = 0xFFFFFFFFu;

// Everything the kernels need, derived from ctf_config_t on the host in ctf_create().
struct DevPlan {
    unsigned long long tag_threshold;
    unsigned long long lut64[2];  // 4 bits per tile code -> channel, per observer team
    double reward_step, reward_capture, reward_tag, capture_punish, win_margin, loss_margin;
    int G, N, C, GG, M, E;        // E = N*C*G*G observation elements per env
    int bits_words;               // words of the per-env observation bit string (alignment pad + slack, multiple of 16)
    int wpa;                      // words per agent of the packed observation output: ceil(C*G*G / 32)
    int game_steps, flip_axis;
    int flip_a, flip_b, flip_d;    // flipped cell index = flip_a * r + flip_b * c + flip_d
    int use_adjusted_rewards, home_flag_capture, drop_flag_when_no_hp, reverse_team1_actions;
    int heal_q, vault_cost_q, vault_min_q;
    int zone_distance, guardian_distance, tagging_range, max_agent_blocks, block_pickup_value;
    int hp_max_q[4], damage_q[4], damage_boosted_q[4];
    // hp_float: HP as IEEE doubles, every operation rounded like the reference's Python floats (non-dyadic configs)
    int hp_float;
    double hp_max_f[4], damage_f[4], damage_boosted_f[4], heal_f, vault_cost_f, vault_min_f;
    int warp_smem_bytes, list_off, bits_off;
    unsigned char team[8], type[8], tile[8], start_r[8], start_c[8], obs_rev[8], meta_hp_src[8];
    signed char my_slot[8];        // index of agent i in OPPONENTS[1 - team(i)], -1 if truncated away
    unsigned char n_opp[2];
    unsigned char flag_pos[2][2], capture_pos[2][2], spawn_pos[2][2], flag_tile[2];
    signed char delta[4][9][2];
    unsigned char rev_action[16];
    unsigned int meta_row[8];        // metadata slots of observer a: nibble q -> agent id (0xF = none)
    unsigned int meta_codes[64];     // kMeta* code of every element of the [N][M] metadata block, four per word
    unsigned char grid_template[kGridBytes];  // 16-stride rows
};

struct Launch {
    // state
    uint8_t* grid;
    unsigned long long* agents;
    uint4* envs;
    uint32_t* stats;
    uint8_t* visits;
    double* hp;             // [B][N] when plan.hp_float
    // outputs
    void* obs;
    uint32_t* obs_bits;
    float* meta;
    float* rewards;
    uint8_t* dones;
    const uint8_t* actions;
    uint32_t* faults;
    long long B;
    uint32_t seed_lo, seed_hi;
    uint32_t env_id_base;
    uint32_t rev_override;  // bit 8 set: bits 0..7 replace plan.obs_rev
    int first_reset;
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ int cheb(int r0, int c0, int r1, int c1) { return max(abs(r0 - r1), abs(c0 - c1)); }

// lane-register form of one agent: row | col<<4 | has_flag<<8 | hp_q<<16 (int16)
__device__ __forceinline__ uint32_t pack_agent(int r, int c, int flag, int hp) {
    return (uint32_t)r | ((uint32_t)c << 4) | ((uint32_t)flag << 8) | ((uint32_t)(hp & 0xFFFF) << 16);
}
__device__ __forceinline__ int ag_r(uint32_t m) { return m & 15; }
__device__ __forceinline__ int ag_c(uint32_t m) { return (m >> 4) & 15; }
__device__ __forceinline__ int ag_flag(uint32_t m) { return (m >> 8) & 1; }
__device__ __forceinline__ int ag_hp(uint32_t m) { return (int)(short)(m >> 16); }

struct WarpMem {
    uint8_t* grid;     // [256]
    uint32_t* list;    // [256]
    uint32_t* bits;    // [bits_words]
};

__device__ __forceinline__ WarpMem warp_mem(const DevPlan& P, unsigned char* smem, int warp) {
    unsigned char* base = smem + (size_t)warp * P.warp_smem_bytes;
    WarpMem w;
    w.grid = base;
    w.list = reinterpret_cast<uint32_t*>(base + P.list_off);
    w.bits = reinterpret_cast<uint32_t*>(base + P.bits_off);
    return w;
}

// flipped destination cell of (r, c) for a reversed view (gridworld_ctf.py:1003-1007); all four maps are involutions
// and affine in (r, c): the host folds FLIP_AXIS into three coefficients (build_plan)
__device__ __forceinline__ int flip_cell(const DevPlan& P, int r, int c) { return P.flip_a * r + P.flip_b * c + P.flip_d; }

// ------------------------------------------------------------------------------------------------
// Observation + metadata writer (standardise_state :975-1009, get_env_metadata :1027-1069)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ uint4 expand_bits(uint32_t b);

template <>
__device__ __forceinline__ uint4 expand_bits<float>(uint32_t b) {  // 4 elements
    uint4 v;
    // 1.0f = 0x3F800000 has 23 trailing zero bits, so (b & 2^k) * (0x3F800000 >> k) is exact for k <= 3
    v.x = (b & 1u) * 0x3F800000u;
    v.y = (b & 2u) * 0x1FC00000u;
    v.z = (b & 4u) * 0x0FE00000u;
    v.w = (b & 8u) * 0x07F00000u;
    return v;
}

template <>
__device__ __forceinline__ uint4 expand_bits<uint8_t>(uint32_t b) {  // 16 elements
    uint4 v;
    v.x = ((b & 15u) * 0x00204081u) & 0x01010101u;
    v.y = (((b >> 4) & 15u) * 0x00204081u) & 0x01010101u;
    v.z = (((b >> 8) & 15u) * 0x00204081u) & 0x01010101u;
    v.w = (((b >> 12) & 15u) * 0x00204081u) & 0x01010101u;
    return v;
}

// 8 elements of a 16-bit float type whose 1.0 is ONE: word k packs bits 2k (low half) and 2k+1 (high half)
template <uint32_t ONE>
__device__ __forceinline__ uint4 expand_bits16(uint32_t b) {
    uint4 v;
    v.x = (((b & 3u) * 0x8001u) & 0x10001u) * ONE;
    v.y = ((((b >> 2) & 3u) * 0x8001u) & 0x10001u) * ONE;
    v.z = ((((b >> 4) & 3u) * 0x8001u) & 0x10001u) * ONE;
    v.w = ((((b >> 6) & 3u) * 0x8001u) & 0x10001u) * ONE;
    return v;
}
template <>
__device__ __forceinline__ uint4 expand_bits<__half>(uint32_t b) { return expand_bits16<0x3C00u>(b); }
template <>
__device__ __forceinline__ uint4 expand_bits<__nv_bfloat16>(uint32_t b) { return expand_bits16<0x3F80u>(b); }

template <typename T>
__device__ __forceinline__ T from_bit(uint32_t bit) { return (T)bit; }
template <>
__device__ __forceinline__ __half from_bit<__half>(uint32_t bit) { return __ushort_as_half((unsigned short)(bit ? 0x3C00u : 0u)); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_bit<__nv_bfloat16>(uint32_t bit) {
    return __ushort_as_bfloat16((unsigned short)(bit ? 0x3F80u : 0u));
}

#if CTF_STATE_HINT || CTF_STORE_OP == 3
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#endif

__device__ __forceinline__ void store_vec(uint4* p, uint4 v) {
#if CTF_STORE_OP == 0
    __stcs(p, v);
#elif CTF_STORE_OP == 1
    *p = v;
#elif CTF_STORE_OP == 2
    __stwt(p, v);
#else
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
                 "l"(l2_evict_first()) : "memory");
#endif
}

// env state accesses (tile map, agent records, env record)
__device__ __forceinline__ uint4 ld_state(const uint4* p) {
#if CTF_STATE_HINT
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_evict_last()));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
#if CTF_STATE_HINT
    unsigned long long v;
    asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(l2_evict_last()));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_state(uint4* p, uint4 v) {
#if CTF_STATE_HINT
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
                 "l"(l2_evict_last()) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
#if CTF_STATE_HINT
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(l2_evict_last()) : "memory");
#else
    *p = v;
#endif
}

// Builds the env's observation block as one bit per element in shared memory (w.bits): element
// e = (a*C + c)*G*G + p of the [N][C][G][G] block is bit e.
// `pad` (obs_pad of the output block) shifts the whole string: element e is bit pad + e, so that the store vectors,
// counted from the alignment boundary below the block, cover whole nibbles / bytes / half-words of the string.
__device__ __forceinline__ void build_obs_bits(const DevPlan& P, const WarpMem& w, uint32_t me, uint32_t rev_mask, int pad, int lane) {
    const int N = P.N, GG = P.GG, CGG = P.C * P.GG;
    // 1. clear the bit string
    {
        uint4* b4 = reinterpret_cast<uint4*>(w.bits);
        const int n4 = (P.bits_words + 3) >> 2;
        for (int i = lane; i < n4; i += 32) b4[i] = make_uint4(0, 0, 0, 0);
    }
    // 2. list the non-open cells: channel(team-0 view) | channel(team-1 view)<<4 | dest(normal)<<8 | dest(flipped)<<16
    int count = 0;
#pragma unroll
    for (int j = 0; j < kGridBytes / 32; ++j) {
        const int cell = j * 32 + lane;
        const uint32_t t = w.grid[cell];
        const unsigned nz = __ballot_sync(kFull, t != 0);
        if (t != 0) {
            const int r = cell >> 4, c = cell & 15;
            const int pos = count + __popc(nz & ((1u << lane) - 1u));
            const uint32_t ch0 = (uint32_t)(P.lut64[0] >> (t * 4)) & 15u, ch1 = (uint32_t)(P.lut64[1] >> (t * 4)) & 15u;
            w.list[pos] = ch0 | (ch1 << 4) | ((uint32_t)(r * P.G + c) << 8) | ((uint32_t)flip_cell(P, r, c) << 16);
        }
        count += __popc(nz);
    }
    __syncwarp();
    // 3. scatter: lane l works for agent (l & 7) on list entries (l >> 3), (l >> 3) + 4, ...; lane a < N also sets
    //    agent a's own position plane (channel 0) from its registers
    {
        const int a = lane & 7;
        const bool valid = a < N;
        const bool rev = (rev_mask >> a) & 1u;
        const int ch_shift = P.team[a] ? 4 : 0, p_shift = rev ? 16 : 8;
        const int base = pad + a * CGG;
        for (int k = lane >> 3; k < count; k += 4) {
            const uint32_t ent = w.list[k];
            const int ch = (ent >> ch_shift) & 15;
            if (valid && ch) {
                const int e = base + ch * GG + (int)((ent >> p_shift) & 0xFFu);
                atomicOr(&w.bits[e >> 5], 1u << (e & 31));
            }
        }
        if (lane < N) {
            const int r = ag_r(me), c = ag_c(me);
            const int e = base + (rev ? flip_cell(P, r, c) : r * P.G + c);
            atomicOr(&w.bits[e >> 5], 1u << (e & 31));
        }
    }
    __syncwarp();
}

// elements (= bits of the string) per 128-bit store
template <typename T>
constexpr int kVecElems = 16 / (int)sizeof(T);

#ifndef CTF_PAD_BYTES
// Alignment unit of the store vectors: 16 (vector), 32 (sector) or 128 (cache line).  Measured on one B200
// (profiles/r02_ab_pad.log, ms per step 8_arena B=65536 / 7_gridlocked B=65536 / B=16384): 16 -> 1.003 / 0.618 / 0.173,
// 32 -> 1.003 / 0.551 / 0.151, 128 -> 1.035 / 0.569 / 0.154: whole sectors matter (a sector shared by two store
// instructions is written twice into L2), whole lines do not.
#define CTF_PAD_BYTES 32
#endif
// number of elements between the previous CTF_PAD_BYTES boundary and `out`: 0 .. CTF_PAD_BYTES / sizeof(T) - 1
template <typename T>
__device__ __forceinline__ int obs_pad(const T* out) {
    return (int)((reinterpret_cast<uintptr_t>(out) & (uintptr_t)(CTF_PAD_BYTES - 1)) / sizeof(T));
}

#ifndef CTF_U8_LUT
#define CTF_U8_LUT 1   // uint8 stream: expand 8 bits -> 8 bytes with one 64-bit shared load from a 256-entry table
#endif
// bytes at the start of a kernel's dynamic shared memory taken by the expansion table of element type T
template <typename T>
constexpr int kLutBytes = (CTF_U8_LUT && sizeof(T) == 1) ? 2048 : 0;

// Fills the table (all threads of the CTA, before any thread leaves the kernel); returns it, or nullptr for other types.
template <typename T>
__device__ __forceinline__ const uint2* init_expand_lut(uint4* smem_raw) {
    if constexpr (kLutBytes<T> > 0) {
        uint2* lut = reinterpret_cast<uint2*>(smem_raw);
        for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x)
            lut[i] = make_uint2(((i & 15u) * 0x00204081u) & 0x01010101u, ((i >> 4) * 0x00204081u) & 0x01010101u);
        __syncthreads();
        return lut;
    } else {
        return nullptr;
    }
}

// Stores the full vectors [v_begin, v_end) of a padded bit string (vector v = bits VB*v .. VB*v + VB - 1) at vp[v];
// warp `wi` of `nw` cooperating warps.  `bits` must be 16-byte aligned and readable up to a multiple of 16 words.
template <typename T>
__device__ __forceinline__ void stream_vectors(const uint32_t* __restrict__ bits, uint4* __restrict__ vp, int v_begin,
                                               int v_end, int wi, int nw, int lane, const uint2* __restrict__ lut) {
    if constexpr (sizeof(T) == 4) {
        // group g = 128 vectors = 16 words = 2 KB of output: lane (q, n) = (lane >> 3, lane & 7) reads words
        // 16g + 4q .. +3 with ONE 128-bit shared load and stores vector 8 * (16g + 4q + m) + n for m = 0..3, so every
        // store instruction writes four complete 128-byte runs.  The per-store bounds checks stay on purpose: a loop
        // without them (and with the next group's load issued early) executes 40 % fewer instructions but is 1 - 7 %
        // SLOWER (profiles/r02_ab_stream_loop*.log) — the HBM write stream likes stores spaced out, not bursts of four
        const int q = lane >> 3, n = lane & 7, sh = n * 4;
        const int n_groups = (v_end + 127) >> 7;
#pragma unroll 2
        for (int g = wi; g < n_groups; g += nw) {
            const uint4 wd = *reinterpret_cast<const uint4*>(bits + 16 * g + 4 * q);
            const int v0 = 128 * g + 32 * q + n;
            if (v0 >= v_begin && v0 < v_end) store_vec(vp + v0, expand_bits<T>(wd.x >> sh));
            if (v0 + 8 < v_end) store_vec(vp + v0 + 8, expand_bits<T>(wd.y >> sh));
            if (v0 + 16 < v_end) store_vec(vp + v0 + 16, expand_bits<T>(wd.z >> sh));
            if (v0 + 24 < v_end) store_vec(vp + v0 + 24, expand_bits<T>(wd.w >> sh));
        }
    } else {
        constexpr int VB = kVecElems<T>, VW = 32 / VB;
        if constexpr (kLutBytes<T> > 0) {
            // uint8: the vector's 16 bits are two table lookups (8 bits -> 8 bytes each) instead of sixteen ALU operations
#pragma unroll 4
            for (int v = wi * 32 + lane; v < v_end; v += nw * 32) {
                if (v >= v_begin) {
                    const uint32_t h = bits[v >> 1] >> ((v & 1) * 16);
                    const uint2 lo = lut[h & 0xFFu], hi = lut[(h >> 8) & 0xFFu];
                    store_vec(vp + v, make_uint4(lo.x, lo.y, hi.x, hi.y));
                }
            }
        } else {
#pragma unroll 4
            for (int v = wi * 32 + lane; v < v_end; v += nw * 32)
                if (v >= v_begin) store_vec(vp + v, expand_bits<T>(bits[v / VW] >> ((v % VW) * VB)));
        }
    }
}

// Streams `nbits` elements to `out` (any element alignment) from a bit string built with pad = obs_pad(out), i.e.
// laid out from the CTF_PAD_BYTES boundary below `out`: store vectors are then sector-aligned whatever the block's
// alignment (7_gridlocked's 52 728-byte env blocks start at every multiple of 8 bytes), no 32-byte sector is shared
// by two store instructions.  Vectors that lie completely inside the block go out as 128-bit stores, the
// fewer than VB elements before the first / after the last full vector as scalar stores (their neighbours belong
// to other envs).
template <typename T>
__device__ __forceinline__ void stream_env(const uint32_t* __restrict__ bits, int nbits, T* __restrict__ out, int wi, int nw,
                                           int lane, const uint2* __restrict__ lut) {
    constexpr int VB = kVecElems<T>;
    const int pad = obs_pad(out);
    const int total = pad + nbits;
    const int v_begin = (pad + VB - 1) / VB, v_end = total / VB;   // vectors counted from the alignment boundary
    if (wi == 0) {
        const int head_n = min(v_begin * VB - pad, nbits);          // < VB elements before the first full vector
        if (lane < head_n) out[lane] = from_bit<T>((bits[(pad + lane) >> 5] >> ((pad + lane) & 31)) & 1u);
        const int t = max(v_end * VB, pad + head_n) + lane;         // first bit covered by neither head nor full vectors
        if (t < total) out[t - pad] = from_bit<T>((bits[t >> 5] >> (t & 31)) & 1u);
    }
    stream_vectors<T>(bits, reinterpret_cast<uint4*>(out - pad), v_begin, v_end, wi, nw, lane, lut);
}

// Packed copy of the observation block for rollout storage: agent a's C*G*G bits start at word a*wpa
// (32x smaller than float32; ctf_unpack_obs expands it again).
__device__ __forceinline__ void store_packed(const DevPlan& P, const uint32_t* __restrict__ bits, int pad,
                                             uint32_t* __restrict__ out_env, int wi, int nw, int lane) {
    const int CGG = P.C * P.GG, wpa = P.wpa, total = P.N * wpa;
    for (int i = wi * 32 + lane; i < total; i += nw * 32) {
        const int a = i / wpa, j = i - a * wpa;
        const int o = pad + a * CGG + 32 * j;
        uint32_t v = __funnelshift_r(bits[o >> 5], bits[(o >> 5) + 1], o & 31);
        const int valid = CGG - 32 * j;            // bits of this word that belong to agent a
        if (valid < 32) v &= (1u << valid) - 1u;
        out_env[i] = v;
    }
}

template <typename T>
__device__ __forceinline__ void write_obs(const DevPlan& P, const Launch& L, const WarpMem& w, long long env, uint32_t me,
                                          uint32_t rev_mask, int lane, const uint2* __restrict__ lut) {
    if (!L.obs && !L.obs_bits) return;
    T* out = L.obs ? reinterpret_cast<T*>(L.obs) + env * (long long)P.E : nullptr;
    const int pad = out ? obs_pad(out) : 0;
    build_obs_bits(P, w, me, rev_mask, pad, lane);
    if (L.obs_bits) store_packed(P, w.bits, pad, L.obs_bits + env * (long long)(P.N * P.wpa), 0, 1, lane);
    if (out) stream_env<T>(w.bits, P.E, out, 0, 1, lane, lut);
}

#ifndef CTF_COOP_STREAM
#define CTF_COOP_STREAM 1   // k_step: 1 = the CTA's warps stream its env blocks together (below), 0 = every warp streams its own
#endif
#ifndef CTF_META_VEC
#define CTF_META_VEC 1   // 1: whole [N][M] block as float4 stores; 0: one 4-byte store per element
#endif
// Element codes of the metadata block (host-built, DevPlan::meta_codes): the LANE whose value the element takes.
// write_meta spreads the block's distinct values over the warp — lane i < 8: hp8 of agent i, lane 8 + i: has_flag of
// agent i, lanes 16 / 17 / 18: game progress and the two capture ratios, lane 19: 1.0, lane 20: 0.0 — so that every
// element is one shuffle.
enum : unsigned { kMetaHp = 0, kMetaFlag = 8, kMetaPct = 16, kMetaRatio0 = 17, kMetaRatio1 = 18, kMetaOne = 19, kMetaZero = 20 };

template <bool HPF>
__device__ __forceinline__ void write_meta(const DevPlan& P, uint32_t me, double hpd, int step, int caps0, int caps1,
                                           float* __restrict__ meta_env, int lane) {
    const int N = P.N, M = P.M;
    const int li = lane & 7;
    // hp8[i] = uint8(agent_hp[TYPE_i as agent id] / AGENT_TYPE_HP[TYPE_i])  (:1039-1041)
    const uint32_t src = __shfl_sync(kFull, me, P.meta_hp_src[li]);
    const uint32_t mine = __shfl_sync(kFull, me, li);
    float hp8;
    if (HPF) {   // the same quotient in double, truncated like numpy's float -> uint8 store
        const double srcd = __shfl_sync(kFull, hpd, P.meta_hp_src[li]);
        hp8 = (float)((int)__ddiv_rn(srcd, P.hp_max_f[P.type[li]]) & 0xFF);
    } else {
        hp8 = (float)((ag_hp(src) / P.hp_max_q[P.type[li]]) & 0xFF);
    }
    // the three fp64 quotients that go through float16 (:1035-1036, :1044), one per lane 16..18, in a single pass
    const int num = lane == 16 ? step : (lane == 17 ? caps0 + 1 : caps1 + 1);
    const int den = lane == 16 ? P.game_steps : (lane == 17 ? caps1 + 1 : caps0 + 1);
    const float quot = __half2float(__double2half((double)num / (double)den));
    const float val = lane < 8 ? hp8 : (lane < 16 ? (float)ag_flag(mine) : (lane < 19 ? quot : (lane == 19 ? 1.0f : 0.0f)));
#if CTF_META_VEC
    // The block's layout is static, so the host compiled it into one source lane per element; lane l of pass `it`
    // produces elements 4k .. 4k+3 (k = 32 it + l) and stores them as one float4 (N*M is a multiple of 4 for every N).
    const int n4 = (N * M) >> 2;
    for (int it = 0; it * 32 < n4; ++it) {
        const int k = it * 32 + lane;
        const uint32_t codes = P.meta_codes[k & 63];
        float4 v;
        v.x = __shfl_sync(kFull, val, codes & 31u);
        v.y = __shfl_sync(kFull, val, (codes >> 8) & 31u);
        v.z = __shfl_sync(kFull, val, (codes >> 16) & 31u);
        v.w = __shfl_sync(kFull, val, (codes >> 24) & 31u);
        if (k < n4) reinterpret_cast<float4*>(meta_env)[k] = v;
    }
#else
    const int total = N * M;                  // one 4-byte store per element
    for (int e0 = 0; e0 < total; e0 += 32) {
        const int e = e0 + lane;
        const uint32_t code = (P.meta_codes[(e >> 2) & 63] >> (8 * (e & 3))) & 31u;
        const float x = __shfl_sync(kFull, val, code);
        if (e < total) meta_env[e] = x;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// State staging
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_state(const DevPlan& P, const Launch& L, const WarpMem& w, long long env,
                                            uint32_t me, double hpd, int inv, uint4 ev, int lane) {
    if (P.hp_float && lane < P.N) L.hp[env * P.N + lane] = hpd;
    if (lane < kGridBytes / 16)
        st_state(reinterpret_cast<uint4*>(L.grid + env * kGridBytes) + lane, reinterpret_cast<const uint4*>(w.grid)[lane]);
    if (lane < P.N) {
        const unsigned long long rec = (unsigned long long)ag_r(me) | ((unsigned long long)ag_c(me) << 8) |
                                       ((unsigned long long)ag_flag(me) << 16) |
                                       ((unsigned long long)(uint16_t)ag_hp(me) << 32) |
                                       ((unsigned long long)(uint16_t)inv << 48);
        st_state(L.agents + env * P.N + lane, rec);
    }
    if (lane == 0) st_state(L.envs + env, ev);
}

// Per-step counter deltas of this lane's agent, 4 bits per metric (every per-step increment is <= 15:
// at most 4 tags / respawns / neighbours, distances <= GRID_SIZE - 1); added to the HBM counters once per step.
struct Deltas {
    uint32_t lo = 0, hi = 0;  // metrics 0..7, 8..12
};

// Every bump of one actor's turn is warp-uniform (same condition and amount in all lanes), so the turn's deltas are
// summed in uniform registers and merged into the actor's lane once per turn.
template <bool STATS>
__device__ __forceinline__ void bump(Deltas& turn, int metric, uint32_t by) {
    if (STATS) {
        if (metric < 8) turn.lo += by << (4 * metric);
        else turn.hi += by << (4 * (metric - 8));
    }
}

// ------------------------------------------------------------------------------------------------
// reset (gridworld_ctf.py:383-477)
// ------------------------------------------------------------------------------------------------
template <typename T, bool STATS>
__global__ void __launch_bounds__(kThreads) k_reset(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L) {
    extern __shared__ uint4 smem_raw[];
    const uint2* lut = init_expand_lut<T>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (env >= L.B) return;
    const WarpMem w = warp_mem(P, reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T>, warp);

    if (lane < kGridBytes / 16)
        reinterpret_cast<uint4*>(w.grid)[lane] = reinterpret_cast<const uint4*>(P.grid_template)[lane];
    const int li = lane & 7;
    const uint32_t me = pack_agent(P.start_r[li], P.start_c[li], 0, P.hp_max_q[P.type[li]]);
    const double hpd = P.hp_float ? P.hp_max_f[P.type[li]] : 0.0;
    uint4 ev = make_uint4(0, 0, 0, 0);
    if (!L.first_reset) ev.y = L.envs[env].y + 1;  // episode
    __syncwarp();
    store_state(P, L, w, env, me, hpd, 0, ev, lane);
    if (STATS) {
        const int ns = CTF_N_METRICS * P.N;
        for (int i = lane; i < ns; i += 32) L.stats[env * ns + i] = 0;
        if (L.visits) {
            uint8_t* v = L.visits + env * (long long)P.N * P.GG;
            for (int i = lane; i < P.N * P.GG; i += 32) v[i] = 0;
            __syncwarp();
            if (lane < P.N) v[lane * P.GG + ag_r(me) * P.G + ag_c(me)] = 1;  // update_visitation_map (:473)
        }
    }
    if (L.rewards && lane < P.N) L.rewards[env * P.N + lane] = 0.0f;
    if (L.dones && lane == 0) L.dones[env] = 0;
    const uint32_t rev_mask = (L.rev_override & 0x100u) ? (L.rev_override & 0xFFu)
                                                        : __ballot_sync(kFull, lane < P.N && P.obs_rev[li]);
    if (L.meta) {   // not a hot kernel: the HP representation is a run-time branch here
        if (P.hp_float) write_meta<true>(P, me, hpd, 0, 0, 0, L.meta + env * (long long)P.N * P.M, lane);
        else write_meta<false>(P, me, 0.0, 0, 0, 0, L.meta + env * (long long)P.N * P.M, lane);
    }
    write_obs<T>(P, L, w, env, me, rev_mask, lane, lut);
}

// ------------------------------------------------------------------------------------------------
// observe only
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_observe(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L) {
    extern __shared__ uint4 smem_raw[];
    const uint2* lut = init_expand_lut<T>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (env >= L.B) return;
    const WarpMem w = warp_mem(P, reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T>, warp);
    if (lane < kGridBytes / 16)
        reinterpret_cast<uint4*>(w.grid)[lane] = ld_state(reinterpret_cast<const uint4*>(L.grid + env * kGridBytes) + lane);
    const int li = lane & 7;
    uint32_t me = 0;
    double hpd = 0.0;
    if (lane < P.N) {
        const unsigned long long rec = ld_state(L.agents + env * P.N + lane);
        me = pack_agent((int)(rec & 0xFF), (int)((rec >> 8) & 0xFF), (int)((rec >> 16) & 1), (int)(short)(rec >> 32));
        if (P.hp_float) hpd = L.hp[env * P.N + lane];
    }
    const uint4 ev = ld_state(L.envs + env);
    __syncwarp();
    const uint32_t rev_mask = (L.rev_override & 0x100u) ? (L.rev_override & 0xFFu)
                                                        : __ballot_sync(kFull, lane < P.N && P.obs_rev[li]);
    if (L.meta) {
        if (P.hp_float) write_meta<true>(P, me, hpd, (int)ev.x, (int)ev.z, (int)ev.w, L.meta + env * (long long)P.N * P.M, lane);
        else write_meta<false>(P, me, 0.0, (int)ev.x, (int)ev.z, (int)ev.w, L.meta + env * (long long)P.N * P.M, lane);
    }
    write_obs<T>(P, L, w, env, me, rev_mask, lane, lut);
}

// ------------------------------------------------------------------------------------------------
// step (gridworld_ctf.py:849-918) + observations for every agent
// ------------------------------------------------------------------------------------------------
// One env step by one warp: state in, act / tag / respawn / heal / rewards / statistics / metadata, state out.
// Returns the lane's agent register (position for the observation's self plane) and the per-agent reverse_grid mask;
// the tile map of the new state is left in w.grid for the observation writer.
// HPF: HP is an IEEE double per agent (plan.hp_float) instead of the fixed-point field of the agent register; every
// operation on it is the reference's Python float operation (:648-650 vault gate, :656 vault cost, :818 damage, :824
// lethal test, :785 respawn, :845-846 heal), so non-dyadic HP configurations stay bit-exact.
template <bool STATS, bool HPF>
__device__ __forceinline__ uint32_t step_env(const DevPlan& P, const Launch& L, const WarpMem& w, long long env, int lane,
                                             uint32_t& rev_mask_out) {
    const int N = P.N;
    const int li = lane & 7;

    // ---- stage state: tile map -> shared memory, agent i -> lane i
#if CTF_DIAG & 1
    if (lane < kGridBytes / 16) reinterpret_cast<uint4*>(w.grid)[lane] = reinterpret_cast<const uint4*>(P.grid_template)[lane];
    uint32_t me = pack_agent(P.start_r[lane & 7], P.start_c[lane & 7], 0, P.hp_max_q[P.type[lane & 7]]);
    int inv = 0;
    int action = (CTF_DIAG & 32) ? (int)((env + lane) % 9) : (lane < N ? L.actions[env * N + lane] : 4);
    uint4 ev = make_uint4((uint32_t)(env & 255), 0, 0, 0);
#else
    if (lane < kGridBytes / 16)
        reinterpret_cast<uint4*>(w.grid)[lane] = ld_state(reinterpret_cast<const uint4*>(L.grid + env * kGridBytes) + lane);
    uint32_t me = 0;
    int inv = 0;
    int action = 4;
    if (lane < N) {
        const unsigned long long rec = ld_state(L.agents + env * N + lane);
        me = pack_agent((int)(rec & 0xFF), (int)((rec >> 8) & 0xFF), (int)((rec >> 16) & 1), (int)(short)(rec >> 32));
        inv = (int)((rec >> 48) & 0xFFFF);
        action = (CTF_DIAG & 32) ? (int)((env + lane) % 9) : L.actions[env * N + lane];
    }
    uint4 ev = ld_state(L.envs + env);
#endif
    double hpd = 0.0;   // this lane's agent's HP (HPF only; dead code otherwise)
    if (HPF && lane < N) hpd = L.hp[env * N + lane];
    Deltas dl;
    const int my_team = P.team[li], my_type = P.type[li], my_slot = P.my_slot[li];
    const bool bad_action = lane < N && action >= CTF_N_ACTIONS;
    if (__any_sync(kFull, bad_action)) {
        if (lane == 0) atomicOr(L.faults, CTF_FAULT_BAD_ACTION);
        if (bad_action) action = 4;
    }
    if (P.reverse_team1_actions && my_team == 1) action = P.rev_action[action];

    // ---- step counter and this step's draws: lane l generates site l
    ev.x += 1;                                  // env_step_count += 1 (:857)
    const int step = (int)ev.x;
    uint32_t wd[4];
    philox4x32_10(L.env_id_base + (uint32_t)env, ev.y, ev.x, (uint32_t)lane, L.seed_lo, L.seed_hi, wd);

    // ---- move order: Fisher-Yates from the identity with word 2 of site i (dice_roll :734-742)
    uint32_t order = 0x76543210u;
    for (int i = N - 1; i >= 1; --i) {
        const uint32_t wi = __shfl_sync(kFull, wd[2], i);
        const int j = (int)__umulhi(wi, (uint32_t)(i + 1));
        const uint32_t x = ((order >> (4 * i)) ^ (order >> (4 * j))) & 15u;
        order ^= (x << (4 * i)) | (x << (4 * j));
    }
    __syncwarp();

    uint32_t turn_pos = 0;      // this lane's agent's position at the end of its own turn (zonal metrics)
    bool captured = false;      // this lane's agent captured during its own act (:728-730)
    bool tag_reward = false;    // this lane's agent made a lethal tag (:832)
    uint32_t cap_team = 0;      // _flag_capture_team_current_move as bits (:858)
    int caps0 = (int)ev.z, caps1 = (int)ev.w;

    for (int s = 0; s < N; ++s) {
        const int a = (order >> (4 * s)) & 15;                 // acting agent (warp-uniform)
        const uint32_t am = __shfl_sync(kFull, me, a);
        const int act_code = __shfl_sync(kFull, action, a);
        const int type = P.type[a], team = P.team[a], tile = P.tile[a];
        int ar = ag_r(am), ac = ag_c(am), aflag = ag_flag(am), ahp = ag_hp(am);
        int ainv = (type == 3) ? __shfl_sync(kFull, inv, a) : 0;
        double ahpd = HPF ? __shfl_sync(kFull, hpd, a) : 0.0;
        bool cap_now = false;
        Deltas turn;                                           // this turn's counter increments (warp-uniform)

        // ---------------- act (:700-732)
        const int nr = ar + P.delta[type][act_code][0], nc = ac + P.delta[type][act_code][1];
        if (nr >= 0 && nr < P.G && nc >= 0 && nc < P.G) {     // is_valid_move (:636-641)
            const int target = w.grid[nr * kRow + nc];
            const bool can_vault = HPF ? (__dsub_rn(ahpd, P.vault_cost_f) > P.vault_min_f) : ((ahp - P.vault_cost_q) > P.vault_min_q);
            if (target == 0 && (act_code <= 3 || (act_code >= 5 && type == 2 && can_vault))) {
                // movement_handler (:569-612); every lane stores the same bytes
                __syncwarp();
                w.grid[ar * kRow + ac] = 0;
                w.grid[nr * kRow + nc] = (uint8_t)tile;
                __syncwarp();
                ar = nr; ac = nc;
                const int ofr = P.flag_pos[1 - team][0], ofc = P.flag_pos[1 - team][1];
                const int hfr = P.flag_pos[team][0], hfc = P.flag_pos[team][1];
                if (cheb(nr, nc, ofr, ofc) <= 1 && w.grid[ofr * kRow + ofc] == P.flag_tile[1 - team]) {  // pickup (:583-591)
                    aflag = 1;
                    __syncwarp();
                    w.grid[ofr * kRow + ofc] = 1;
                    __syncwarp();
                    bump<STATS>(turn, CTF_M_FLAG_PICKUPS, 1);
                }
                if (cheb(nr, nc, hfr, hfc) <= 1 && aflag == 1 &&
                    (!P.home_flag_capture || w.grid[hfr * kRow + hfc] == P.flag_tile[team])) {        // capture (:594-610)
                    aflag = 0;
                    __syncwarp();
                    w.grid[ofr * kRow + ofc] = P.flag_tile[1 - team];
                    __syncwarp();
                    if (team == 0) caps0 += 1; else caps1 += 1;
                    cap_team |= 1u << team;
                    cap_now = true;
                    bump<STATS>(turn, CTF_M_FLAG_CAPTURES, 1);
                }
                if (act_code >= 5 && type == 2) {                                                   // (:652-657)
                    if (HPF) ahpd = __dsub_rn(ahpd, P.vault_cost_f);
                    else ahp -= P.vault_cost_q;
                }
            } else if (act_code >= 5 && type == 3 && ainv > 0 && target == 0 &&
                       cheb(nr, nc, P.spawn_pos[team][0], P.spawn_pos[team][1]) > 1 &&
                       cheb(nr, nc, P.spawn_pos[1 - team][0], P.spawn_pos[1 - team][1]) > 1) {
                // add_block (:614-634)
                __syncwarp();
                w.grid[nr * kRow + nc] = 2;
                __syncwarp();
                ainv -= 1;
                if (STATS) {
                    bump<STATS>(turn, CTF_M_BLOCKS_LAID, 1);
                    bump<STATS>(turn, CTF_M_BLOCKS_LAID_DIST_OWN_FLAG, (uint32_t)cheb(ar, ac, P.capture_pos[team][0], P.capture_pos[team][1]));
                    bump<STATS>(turn, CTF_M_BLOCKS_LAID_DIST_OPP_FLAG, (uint32_t)cheb(ar, ac, P.capture_pos[1 - team][0], P.capture_pos[1 - team][1]));
                }
            } else if (act_code < 5 && type == 3 && (target == 2 || target == 3)) {
                // mine_block (:677-690)
                __syncwarp();
                w.grid[nr * kRow + nc] = (target == 2) ? 3 : 0;
                __syncwarp();
                if (target == 3) {
                    if (ainv < P.max_agent_blocks) ainv += P.block_pickup_value;
                    bump<STATS>(turn, CTF_M_BLOCKS_MINED, 1);
                }
            }
        }
        if (lane == a) {
            me = pack_agent(ar, ac, aflag, ahp);
            if (HPF) hpd = ahpd;
            inv = ainv;
            captured = cap_now;
        }

        // distance of this lane's agent to the actor after its move: tag range now, adjacency metrics below
        // (recomputed there only if a respawn moved somebody in between)
        const int d_me = cheb(ar, ac, ag_r(me), ag_c(me));
        bool moved_by_respawn = false;

        // ---------------- tagging_logic (:796-837)
        if (HPF ? (P.damage_f[type] > 0.0) : (P.damage_q[type] > 0)) {
            const bool boosted = type == 1 && cheb(ar, ac, P.flag_pos[team][0], P.flag_pos[team][1]) <= P.guardian_distance;
            const int dmg = boosted ? P.damage_boosted_q[type] : P.damage_q[type];
            const double dmgd = HPF ? (boosted ? P.damage_boosted_f[type] : P.damage_f[type]) : 0.0;
            const bool is_opp = lane < N && my_team != team && my_slot >= 0;
            const int site = 4 * a + (my_slot < 0 ? 0 : my_slot);
            const uint32_t roll = __shfl_sync(kFull, wd[0], site);
            const uint32_t pick_word = __shfl_sync(kFull, wd[1], site);
            const bool hit = is_opp && (unsigned long long)roll < P.tag_threshold && d_me <= P.tagging_range;
            int hp = ag_hp(me);
            if (hit) hp -= dmg;                                                                      // (:818)
            if (HPF && hit) hpd = __dsub_rn(hpd, dmgd);
            const bool lethal = hit && (HPF ? (hpd <= 0.0) : (hp <= 0));                             // (:824)
            if (!HPF && hit) me = (me & 0xFFFFu) | ((uint32_t)(hp & 0xFFFF) << 16);
            const unsigned hits = __ballot_sync(kFull, hit);
            unsigned deaths = __ballot_sync(kFull, lethal);
            moved_by_respawn = deaths != 0;
            if (STATS && hits) bump<STATS>(turn, CTF_M_TAG_COUNT, (uint32_t)__popc(hits));
            // lethal hits respawn one after the other in opponent-id order: each changes the next one's window
            while (deaths) {
                const int opp = __ffs(deaths) - 1;
                deaths &= deaths - 1;
                const uint32_t om = __shfl_sync(kFull, me, opp);
                const uint32_t ow = __shfl_sync(kFull, pick_word, opp);
                const int oteam = 1 - team;
                if (ag_flag(om)) bump<STATS>(turn, CTF_M_FLAG_DISPOSSESSIONS, 1);
                // respawn (:761-794): open cells of the clipped 3x3 window around the victim's spawn, row-major
                const int x = P.spawn_pos[oteam][0], y = P.spawn_pos[oteam][1];
                const int wr = x - 1 + lane / 3, wc = y - 1 + lane % 3;
                const bool open = lane < 9 && wr >= 0 && wr < P.G && wc >= 0 && wc < P.G && w.grid[wr * kRow + wc] == 0;
                const unsigned cand = __ballot_sync(kFull, open);
                const int k = __popc(cand);
                if (k > 0) {
                    const int pick = (int)__umulhi(ow, (uint32_t)k);                                 // randint(k) (:771)
                    const int bit = __fns(cand, 0, pick + 1);
                    const int rr = x - 1 + bit / 3, rc = y - 1 + bit % 3;
                    const int old_r = ag_r(om), old_c = ag_c(om);
                    __syncwarp();
                    w.grid[old_r * kRow + old_c] = 0;
                    w.grid[rr * kRow + rc] = P.tile[opp];
                    if (ag_flag(om)) {                                                                // (:788-794)
                        if (P.drop_flag_when_no_hp) w.grid[old_r * kRow + old_c] = P.flag_tile[1 - oteam];
                        else w.grid[P.flag_pos[1 - oteam][0] * kRow + P.flag_pos[1 - oteam][1]] = P.flag_tile[1 - oteam];
                    }
                    __syncwarp();
                    if (lane == opp) {
                        me = pack_agent(rr, rc, 0, HPF ? 0 : P.hp_max_q[P.type[opp]]);
                        if (HPF) hpd = P.hp_max_f[P.type[opp]];                                       // (:785)
                    }
                } else if (lane == 0) {
                    atomicOr(L.faults, CTF_FAULT_RESPAWN_BLOCKED);   // randint(0) raises in the reference (:771)
                }
                if (lane == a) tag_reward = true;
                bump<STATS>(turn, CTF_M_RESPAWN_TAG_COUNT, 1);
            }
        }

        // ---------------- proximity metrics (:890-902); the zonal ones (:879-889) only need the actor's position at
        // its own turn, which lane a keeps in turn_pos: they are evaluated for all agents at once after the loop
        if (STATS) {
            if (lane == a) turn_pos = (uint32_t)ar | ((uint32_t)ac << 4);
            const int d_now = moved_by_respawn ? cheb(ar, ac, ag_r(me), ag_c(me)) : d_me;
            const bool near = lane < N && my_slot >= 0 && d_now <= 1;
            const unsigned mates = __ballot_sync(kFull, near && my_team == team);   // includes the agent itself
            const unsigned opps = __ballot_sync(kFull, near && my_team != team);
            if (mates) bump<STATS>(turn, CTF_M_STEPS_ADJ_TEAMMATE, (uint32_t)__popc(mates));
            if (opps) bump<STATS>(turn, CTF_M_STEPS_ADJ_OPPONENT, (uint32_t)__popc(opps));
            if (lane == a) { dl.lo += turn.lo; dl.hi += turn.hi; }
        }
    }

    // ---- zonal metrics (:879-889) of every agent, at the position it had at its own turn
    if (STATS) {
        const int tr = (int)(turn_pos & 15u), tc = (int)(turn_pos >> 4);
        const int d_own = cheb(tr, tc, P.capture_pos[my_team][0], P.capture_pos[my_team][1]);
        const int d_opp = cheb(tr, tc, P.capture_pos[1 - my_team][0], P.capture_pos[1 - my_team][1]);
        dl.hi += ((d_own <= P.zone_distance ? 1u : 0u) << (4 * (CTF_M_STEPS_DEFENDING_ZONE - 8))) +
                 ((d_opp <= P.zone_distance ? 1u : 0u) << (4 * (CTF_M_STEPS_ATTACKING_ZONE - 8)));
    }
    // ---- heal_agents (:839-847)
    {
        if (HPF) {
            const double mxd = P.hp_max_f[my_type];
            if (hpd < mxd) hpd = fmin(__dadd_rn(hpd, P.heal_f), mxd);
        } else {
            const int mx = P.hp_max_q[my_type];
            int hp = ag_hp(me);
            if (hp < mx) hp = min(hp + P.heal_q, mx);
            me = (me & 0xFFFFu) | ((uint32_t)(hp & 0xFFFF) << 16);
        }
    }
    // ---- rewards (:727-730, :873, :957-966, :920-940) in fp64, stored as fp32 (ppo.py:108)
    const bool done = step >= P.game_steps;   // set at step == GAME_STEPS and never cleared until reset (:914-915)
    {
        // every operation is a separately rounded IEEE add / multiply (__dadd_rn / __dmul_rn are never contracted
        // into an FMA): -0.5 + 5 * 0.1 must give the reference's exact 0.0, not the fused 2.8e-17
        double r = 0.0;
        r = __dadd_rn(r, P.reward_step);
        if (captured) r = __dadd_rn(r, P.reward_capture);
        if (tag_reward) r = __dadd_rn(r, P.reward_tag);
        if (P.use_adjusted_rewards && ((cap_team >> (1 - my_team)) & 1u)) r = __dadd_rn(r, -P.capture_punish);
        if (step == P.game_steps && caps0 != caps1) {
            const double margin = (double)abs(caps0 - caps1);
            const int winner = caps0 > caps1 ? 0 : 1;
            if (my_team == winner) r = __dadd_rn(r, __dmul_rn(margin, P.win_margin));
            else r = __dadd_rn(r, -__dmul_rn(margin, P.loss_margin));
        }
        if (!(CTF_DIAG & 8) && L.rewards && lane < N) L.rewards[env * N + lane] = (float)r;
    }
    if (!(CTF_DIAG & 8) && L.dones && lane == 0) L.dones[env] = done ? 1 : 0;
    ev.z = (uint32_t)caps0;
    ev.w = (uint32_t)caps1;

    // ---- write state back
    __syncwarp();
    if (!(CTF_DIAG & 2)) store_state(P, L, w, env, me, hpd, inv, ev, lane);
    if (STATS && !(CTF_DIAG & 16)) {
        if (lane < N) {   // fire-and-forget reductions: no load latency, nothing held in registers
            uint32_t* sp = L.stats + env * (long long)(CTF_N_METRICS * N) + lane;
#pragma unroll
            for (int m = 0; m < CTF_N_METRICS; ++m) {
                const uint32_t v = ((m < 8 ? dl.lo : dl.hi) >> (4 * (m & 7))) & 15u;
                if (v) {
#if CTF_STATE_HINT
                    asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(sp + m * N), "r"(v), "l"(l2_evict_last()) : "memory");
#else
                    atomicAdd(sp + m * N, v);
#endif
                }
            }
        }
        if (L.visits && lane < N) {   // update_visitation_map (:911), uint8 wrap
            uint8_t* v = L.visits + (env * N + lane) * (long long)P.GG + ag_r(me) * P.G + ag_c(me);
            *v = (uint8_t)(*v + 1);
        }
    }

    rev_mask_out = (L.rev_override & 0x100u) ? (L.rev_override & 0xFFu) : __ballot_sync(kFull, lane < N && P.obs_rev[li]);
    if (!(CTF_DIAG & 4) && L.meta) write_meta<HPF>(P, me, hpd, step, caps0, caps1, L.meta + env * (long long)N * P.M, lane);
    return me;
}

// Warp-per-env step kernel, the default kernel of ctf_step: every warp steps its env and builds that env's observation
// bit string in its own shared-memory buffer; after one barrier the CTA's warps write the CTA's kWarpsPerCta consecutive
// env blocks one after the other, all warps on the same block (2 KB contiguous per pass) — a quarter as many open
// write streams as when every warp streams its own block, and each CTA's region goes out in address order.
// profiles/r02_coop_stream.log: 8_arena float32 0.980 -> 0.974 ms, uint8 0.283 -> 0.281, 7_gridlocked B = 16384 0.1475 -> 0.1450;
// 4 warps x 5 resident CTAs and 2 x 9 tie, 8 x 3 (0.987) and 16 x 1 (1.18) lose: the barrier idles too much of the SM.
// Per-env ready flags instead of the barrier (a block is written by whichever warps are ready) are SLOWER, 0.9825 vs 0.9786 ms,
// uint8 0.295 vs 0.2835 (profiles/r02_coop_flag.log): the aligned passes are what helps.
template <typename T, bool STATS, bool HPF>
__global__ void __launch_bounds__(kThreads) k_step(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L) {
    extern __shared__ uint4 smem_raw[];
    const uint2* lut = init_expand_lut<T>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * kWarpsPerCta + warp;
#if CTF_COOP_STREAM
    unsigned char* base = reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T>;
    if (env < L.B) {
        const WarpMem w = warp_mem(P, base, warp);
        uint32_t rev_mask;
        const uint32_t me = step_env<STATS, HPF>(P, L, w, env, lane, rev_mask);
        if (L.obs || L.obs_bits) {
            T* out = L.obs ? reinterpret_cast<T*>(L.obs) + env * (long long)P.E : nullptr;
            const int pad = out ? obs_pad(out) : 0;
            build_obs_bits(P, w, me, rev_mask, pad, lane);
            if (L.obs_bits) store_packed(P, w.bits, pad, L.obs_bits + env * (long long)(P.N * P.wpa), 0, 1, lane);
        }
    }
    __syncthreads();
    if (L.obs) {
        for (int e = 0; e < kWarpsPerCta; ++e) {
            const long long env_e = (long long)blockIdx.x * kWarpsPerCta + e;
            if (env_e >= L.B) break;
            stream_env<T>(warp_mem(P, base, e).bits, P.E, reinterpret_cast<T*>(L.obs) + env_e * (long long)P.E, warp, kWarpsPerCta, lane, lut);
        }
    }
#else
    if (env >= L.B) return;
    const WarpMem w = warp_mem(P, reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T>, warp);
    uint32_t rev_mask;
    const uint32_t me = step_env<STATS, HPF>(P, L, w, env, lane, rev_mask);
    write_obs<T>(P, L, w, env, me, rev_mask, lane, lut);   // observations straight into the policy's input buffers
#endif
}

// ------------------------------------------------------------------------------------------------
// Persistent, warp-specialised step kernel (opt-in: CTF_WS=1)
// ------------------------------------------------------------------------------------------------
// What bounds the step is the 100 KB-per-env observation write, and HBM takes that stream fastest when few env
// blocks are open at a time and they are visited in address order (profiles/r02_write_span_microbench.log: 0.88 ms
// for 65536 blocks written one per 512-thread CTA in order, 0.96 ms when every warp streams its own block).  SMs
// also differ by up to 1.46x in store bandwidth, so work has to be handed out dynamically.  Hence:
//   * one persistent CTA per SM (or two): n_logic "logic" warps + n_stream "stream" warps;
//   * a logic warp takes the next env id from a global counter (in order), runs step_env, builds the env's
//     observation bit string in its own shared-memory buffer and queues it in a shared-memory FIFO;
//   * the stream warps take the FIFO entries in order and write ONE env block at a time, together;
//   * a logic warp reuses its buffer when the stream group is done with it — the queue is the back-pressure.
#ifndef CTF_WS_PROFILE
#define CTF_WS_PROFILE 0         // 1: per-phase cycle totals of logic / stream warps (tools/ws_sweep.py --profile build)
#endif
#if CTF_WS_PROFILE
#define CTF_PROF_T(var) const long long var = clock64()
#define CTF_PROF_ADD(acc, t0, t1) acc += (t1) - (t0)
#else
#define CTF_PROF_T(var)
#define CTF_PROF_ADD(acc, t0, t1)
#endif
#ifndef CTF_WS_SLEEP_NS
#define CTF_WS_SLEEP_NS 32       // back-off of the FIFO / buffer polls
#endif
constexpr int kWsQ = 32;         // FIFO slots (>= logic warps per CTA)
constexpr int kWsCtlBytes = 1024;
struct WsCtl {
    int tail;                    // next ticket
    int producers;               // logic warps still running
    volatile int ready[kWsQ];    // ticket + 1 once the entry is filled, 0 when free
    int env[kWsQ];
    int buf[kWsQ];               // logic warp whose buffer holds the bit string
    int done[kWsQ];              // stream warps finished with the entry
    volatile int busy[32];       // logic warp's buffer is queued / being streamed
};
static_assert(sizeof(WsCtl) <= kWsCtlBytes, "control block");

template <typename T, bool STATS, bool HPF>
__global__ void __launch_bounds__(1024, 1) k_step_ws(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L,
                                                     int n_logic, int n_stream, unsigned int* __restrict__ ctr) {
    extern __shared__ uint4 smem_raw[];
    const uint2* lut = init_expand_lut<T>(smem_raw);
    WsCtl* c = reinterpret_cast<WsCtl*>(reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T>);
    unsigned char* warp_base = reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T> + kWsCtlBytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { c->tail = 0; c->producers = n_logic; }
    if (threadIdx.x < kWsQ) { c->ready[threadIdx.x] = 0; c->done[threadIdx.x] = 0; }
    if (threadIdx.x < 32) c->busy[threadIdx.x] = 0;
    __syncthreads();
    const bool any_obs = L.obs || L.obs_bits;
#if CTF_WS_PROFILE
    long long pf[6] = {0, 0, 0, 0, 0, 0};   // logic: fetch, step, wait, build+publish; stream: idle, stream
    unsigned long long* prof = reinterpret_cast<unsigned long long*>(ctr + 4);
#endif
    if (warp < n_logic) {
        const WarpMem w = warp_mem(P, warp_base, warp);
        for (;;) {
            CTF_PROF_T(t0);
            unsigned int e = 0;
            if (lane == 0) e = atomicAdd(ctr, 1u);
            e = __shfl_sync(kFull, e, 0);
            if ((long long)e >= L.B) break;
            CTF_PROF_T(t1);
            uint32_t rev_mask;
            const uint32_t me = step_env<STATS, HPF>(P, L, w, (long long)e, lane, rev_mask);
            if (!any_obs) continue;
            CTF_PROF_T(t2);
            if (lane == 0) while (c->busy[warp]) __nanosleep(CTF_WS_SLEEP_NS);   // previous env of this warp still being streamed
            __syncwarp();
            CTF_PROF_T(t3);
            const int pad = L.obs ? obs_pad(reinterpret_cast<T*>(L.obs) + (long long)e * P.E) : 0;
            build_obs_bits(P, w, me, rev_mask, pad, lane);
            if (lane == 0) {
                c->busy[warp] = 1;
                const int t = atomicAdd(&c->tail, 1);
                const int slot = t % kWsQ;
                while (c->ready[slot] != 0) __nanosleep(CTF_WS_SLEEP_NS);
                c->env[slot] = (int)e;
                c->buf[slot] = warp;
                __threadfence_block();
                c->ready[slot] = t + 1;
            }
            __syncwarp();
            CTF_PROF_T(t4);
            CTF_PROF_ADD(pf[0], t0, t1); CTF_PROF_ADD(pf[1], t1, t2); CTF_PROF_ADD(pf[2], t2, t3); CTF_PROF_ADD(pf[3], t3, t4);
        }
        if (lane == 0) {
#if CTF_WS_PROFILE
            for (int i = 0; i < 4; ++i) atomicAdd(prof + i, (unsigned long long)pf[i]);
#endif
            atomicSub(&c->producers, 1);
            // every logic warp of the grid fails its last fetch exactly once; the last one re-arms the counters
            if (atomicAdd(ctr + 1, 1u) == gridDim.x * (unsigned)n_logic - 1u) { ctr[0] = 0; ctr[1] = 0; }
        }
    } else {
        const int wi = warp - n_logic;
        for (int head = 0;; ++head) {
            const int slot = head % kWsQ;
            int got = 0;
            CTF_PROF_T(t0);
            if (lane == 0) {
                for (;;) {
                    if (c->ready[slot] == head + 1) { got = 1; break; }
                    if (*(volatile int*)&c->producers == 0 && *(volatile int*)&c->tail == head) break;
                    __nanosleep(CTF_WS_SLEEP_NS);
                }
                __threadfence_block();
            }
            got = __shfl_sync(kFull, got, 0);
            if (!got) break;
            CTF_PROF_T(t1);
            const long long env = c->env[slot];
            const int buf = c->buf[slot];
            const uint32_t* bits = warp_mem(P, warp_base, buf).bits;
            T* out = L.obs ? reinterpret_cast<T*>(L.obs) + env * (long long)P.E : nullptr;
            if (L.obs_bits) store_packed(P, bits, out ? obs_pad(out) : 0, L.obs_bits + env * (long long)(P.N * P.wpa), wi, n_stream, lane);
            if (out) stream_env<T>(bits, P.E, out, wi, n_stream, lane, lut);
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(&c->done[slot], 1) == n_stream - 1) {   // last stream warp frees the buffer and the slot
                    c->done[slot] = 0;
                    c->busy[buf] = 0;
                    __threadfence_block();
                    c->ready[slot] = 0;
                }
            }
            CTF_PROF_T(t2);
            CTF_PROF_ADD(pf[4], t0, t1); CTF_PROF_ADD(pf[5], t1, t2);
        }
#if CTF_WS_PROFILE
        if (lane == 0) { atomicAdd(prof + 4, (unsigned long long)pf[4]); atomicAdd(prof + 5, (unsigned long long)pf[5]); }
#endif
    }
}

// ------------------------------------------------------------------------------------------------
// packed observations -> policy input: one CTA per chunk of consecutive agent blocks
// ------------------------------------------------------------------------------------------------
// The chunk's packed words (wpa per block, coalesced) are re-packed in shared memory into ONE contiguous bit string
// (block b's nbits elements follow block b-1's directly, shifted by the output's alignment pad) and all warps stream
// the chunk's output region together — CTAs are scheduled in address order, so few blocks are open at a time.
#ifndef CTF_UNPACK_THREADS
#define CTF_UNPACK_THREADS 256   // 512 measured too: profiles/r02_unpack_threads.log
#endif
constexpr int kUnpackThreads = CTF_UNPACK_THREADS;
template <typename T>
__global__ void __launch_bounds__(kUnpackThreads) k_unpack(const uint32_t* __restrict__ packed, T* __restrict__ out,
                                                          long long n_blocks, int nbits, int wpa, int chunk, int bits_words) {
    extern __shared__ uint4 smem_raw[];
    const uint2* lut = init_expand_lut<T>(smem_raw);
    uint32_t* bits = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(smem_raw) + kLutBytes<T>);   // [bits_words]
    uint32_t* pk = bits + bits_words;                          // [chunk * wpa]
    const long long blk0 = (long long)blockIdx.x * chunk;
    const int nb = (int)min((long long)chunk, n_blocks - blk0);
    if (nb <= 0) return;
    for (int i = threadIdx.x; i < nb * wpa; i += kUnpackThreads) pk[i] = __ldg(packed + blk0 * wpa + i);
    __syncthreads();
    T* dst = out + blk0 * (long long)nbits;
    const int pad = obs_pad(dst), total = nb * nbits;
    for (int wd = threadIdx.x; wd < bits_words; wd += kUnpackThreads) {
        uint32_t val = 0;
        int k = 0, e = 32 * wd - pad;              // bit k of this word is element e + k of the chunk
        if (e < 0) { k = -e; e = 0; }
        while (k < 32 && e < total) {
            const int b = e / nbits, o = e - b * nbits, j = o >> 5;
            const int take = min(32 - k, nbits - o);
            const uint32_t lo = pk[b * wpa + j], hi = j + 1 < wpa ? pk[b * wpa + j + 1] : 0u;
            uint32_t v = __funnelshift_r(lo, hi, o & 31);
            if (take < 32) v &= (1u << take) - 1u;
            val |= v << k;
            k += take; e += take;
        }
        bits[wd] = val;
    }
    __syncthreads();
    stream_env<T>(bits, total, dst, threadIdx.x >> 5, kUnpackThreads / 32, threadIdx.x & 31, lut);
}

// ------------------------------------------------------------------------------------------------
// per-handle sum of the per-env counters
// ------------------------------------------------------------------------------------------------
__global__ void k_stats_sum(const uint32_t* __restrict__ stats, long long B, int ns, unsigned long long* __restrict__ out) {
    // thread t handles counter (t % ns) of envs blockIdx.x*chunk + ...; consecutive threads read consecutive words
    const long long chunk = (B + gridDim.x - 1) / gridDim.x;
    const long long b0 = blockIdx.x * chunk, b1 = min(B, b0 + chunk);
    const int per_pass = blockDim.x / ns;       // envs handled concurrently by this block
    const int sub = threadIdx.x / ns, k = threadIdx.x % ns;
    if (sub >= per_pass) return;
    unsigned long long acc = 0;
    for (long long b = b0 + sub; b < b1; b += per_pass) acc += stats[b * ns + k];
    if (acc) atomicAdd(&out[k], acc);
}

}  // namespace

// ================================================================================================
// Host side: C ABI
// ================================================================================================
struct ctf_env {
    DevPlan plan;
    ctf_config_t cfg;
    long long B;
    int device;
    int stats_level;
    int obs_dtype;
    uint64_t seed;
    uint64_t env_id_base;
    uint32_t* faults;        // device
    uint8_t* actions_stage;  // device [B][N], for ctf_step_host
    float* rewards_stage;    // device [B][N], used by ctf_step_host when out.rewards is NULL
    uint8_t* dones_stage;    // device [B]
    size_t smem_bytes;
    // persistent warp-specialised step kernel (k_step_ws)
    unsigned int* ws_ctr;    // device [2]: next env id, logic warps that have finished; re-armed by the kernel itself
    int ws_logic, ws_stream, ws_ctas;   // logic / stream warps per CTA, CTAs in the grid
    long long ws_min_envs;   // batches below this use the warp-per-env kernel
    int n_sm;
    int k_step_ctas_per_sm;  // resident-CTA cap of the warp-per-env kernels (0: whatever fits)
    size_t ws_smem_bytes;
};

// cudaSetDevice for the duration of an entry point; the caller's current device is restored on return
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define CTF_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) return fail(CTF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* ctf_last_error(void) { return g_err; }
extern "C" int ctf_abi_version(void) { return CTF_ABI_VERSION; }
extern "C" size_t ctf_config_size(void) { return sizeof(ctf_config_t); }

static int build_plan(const ctf_config_t& c, int stats_level, int obs_dtype, DevPlan& P) {
    memset(&P, 0, sizeof(P));
    if (c.grid_size < 2 || c.grid_size > CTF_MAX_GRID) return fail(CTF_ERR_INVALID, "grid_size out of range");
    if (c.n_agents < 1 || c.n_agents > CTF_MAX_AGENTS) return fail(CTF_ERR_INVALID, "n_agents out of range");
    if (c.n_channels < 1 || c.n_channels > CTF_MAX_CHANNELS) return fail(CTF_ERR_INVALID, "n_channels out of range");
    if (c.flip_axis < -1 || c.flip_axis > 2) return fail(CTF_ERR_INVALID, "flip_axis must be -1, 0, 1 or 2");
    if (c.game_steps < 1) return fail(CTF_ERR_INVALID, "game_steps must be positive");
    if (obs_dtype < CTF_OBS_F32 || obs_dtype > CTF_OBS_BF16) return fail(CTF_ERR_INVALID, "unknown obs dtype");
    const int G = c.grid_size, N = c.n_agents;
    P.tag_threshold = c.tag_threshold;
    P.reward_step = c.reward_step; P.reward_capture = c.reward_capture; P.reward_tag = c.reward_tag;
    P.capture_punish = c.capture_punish; P.win_margin = c.win_margin_scalar; P.loss_margin = c.loss_margin_scalar;
    P.G = G; P.N = N; P.C = c.n_channels; P.GG = G * G; P.M = 6 + 2 * N; P.E = N * c.n_channels * G * G;
    P.bits_words = (((P.E + 128 + 31) / 32 + 1) + 15) / 16 * 16;  // pad (<= 127 bits) + 1 slack word, whole 16-word groups
    P.wpa = (c.n_channels * G * G + 31) / 32;
    P.game_steps = c.game_steps; P.flip_axis = c.flip_axis;
    {   // out[i][j] = in[fr][fc] with (fr, fc) = (G1-r, G1-c) both axes, (G1-r, c) rows, (r, G1-c) columns, (G1-c, G1-r) anti-diagonal
        const int G1 = G - 1;
        switch (c.flip_axis) {
            case -1: P.flip_a = -G; P.flip_b = -1; P.flip_d = G1 * G + G1; break;
            case 0: P.flip_a = -G; P.flip_b = 1; P.flip_d = G1 * G; break;
            case 1: P.flip_a = G; P.flip_b = -1; P.flip_d = G1; break;
            default: P.flip_a = -1; P.flip_b = -G; P.flip_d = G1 * G + G1; break;
        }
    }
    P.use_adjusted_rewards = c.use_adjusted_rewards; P.home_flag_capture = c.home_flag_capture;
    P.drop_flag_when_no_hp = c.drop_flag_when_no_hp; P.reverse_team1_actions = c.reverse_team1_actions;
    P.heal_q = c.heal_q; P.vault_cost_q = c.vault_cost_q; P.vault_min_q = c.vault_min_q;
    P.zone_distance = c.zone_distance; P.guardian_distance = c.guardian_distance; P.tagging_range = c.tagging_range;
    P.max_agent_blocks = c.max_agent_blocks; P.block_pickup_value = c.block_pickup_value;
    P.hp_float = c.hp_float ? 1 : 0;
    P.heal_f = c.heal_f; P.vault_cost_f = c.vault_cost_f; P.vault_min_f = c.vault_min_f;
    for (int t = 0; t < 4; ++t) {
        if (P.hp_float) {
            if (!(c.hp_max_f[t] > 0.0)) return fail(CTF_ERR_INVALID, "hp_max_f must be positive");
        } else if (c.hp_max_q[t] <= 0 || c.hp_max_q[t] > 32000) {
            return fail(CTF_ERR_INVALID, "hp_max_q out of int16 range");
        }
        P.hp_max_q[t] = P.hp_float ? 1 : c.hp_max_q[t]; P.damage_q[t] = c.damage_q[t]; P.damage_boosted_q[t] = c.damage_boosted_q[t];
        P.hp_max_f[t] = c.hp_max_f[t]; P.damage_f[t] = c.damage_f[t]; P.damage_boosted_f[t] = c.damage_boosted_f[t];
    }
    for (int i = 0; i < N; ++i) {
        if (c.agent_team[i] > 1 || c.agent_type[i] > 3) return fail(CTF_ERR_INVALID, "agent team/type out of range");
        if (c.start_row[i] >= G || c.start_col[i] >= G) return fail(CTF_ERR_INVALID, "agent start outside the grid");
        if (c.meta_hp_src[i] >= N) return fail(CTF_ERR_INVALID, "meta_hp_src outside 0..N-1");
        P.team[i] = c.agent_team[i]; P.type[i] = c.agent_type[i]; P.tile[i] = c.agent_tile[i];
        P.start_r[i] = c.start_row[i]; P.start_c[i] = c.start_col[i];
        P.obs_rev[i] = c.obs_reverse[i]; P.meta_hp_src[i] = c.meta_hp_src[i];
        P.my_slot[i] = -1;
    }
    for (int t = 0; t < 2; ++t) {
        if (c.n_opponents[t] > 4) return fail(CTF_ERR_INVALID, "more than 4 opponents per team");
        P.n_opp[t] = c.n_opponents[t];
        for (int j = 0; j < c.n_opponents[t]; ++j) {
            const int k = c.opponents[t][j];
            if (k >= N || c.agent_team[k] != 1 - t) return fail(CTF_ERR_INVALID, "opponents list inconsistent with teams");
            if (j > 0 && c.opponents[t][j - 1] >= k) return fail(CTF_ERR_INVALID, "opponents list must be in id order");
            P.my_slot[k] = (signed char)j;
        }
        for (int d = 0; d < 2; ++d) {
            if (c.flag_pos[t][d] >= G || c.capture_pos[t][d] >= G || c.spawn_pos[t][d] >= G)
                return fail(CTF_ERR_INVALID, "flag/capture/spawn position outside the grid");
            if (c.spawn_pos[t][d] < 1) return fail(CTF_ERR_INVALID, "spawn positions need row, col >= 1");
            P.flag_pos[t][d] = c.flag_pos[t][d]; P.capture_pos[t][d] = c.capture_pos[t][d]; P.spawn_pos[t][d] = c.spawn_pos[t][d];
        }
        P.flag_tile[t] = c.flag_tile[t];
        unsigned long long lut = 0;
        for (int tile = 0; tile < CTF_N_TILE_CODES; ++tile) {
            if (c.chan_lut[t][tile] >= c.n_channels) return fail(CTF_ERR_INVALID, "chan_lut entry >= n_channels");
            lut |= (unsigned long long)(c.chan_lut[t][tile] & 15) << (4 * tile);
        }
        P.lut64[t] = lut;
    }
    for (int t = 0; t < 4; ++t)
        for (int a = 0; a < CTF_N_ACTIONS; ++a) { P.delta[t][a][0] = c.action_delta[t][a][0]; P.delta[t][a][1] = c.action_delta[t][a][1]; }
    for (int a = 0; a < 16; ++a) P.rev_action[a] = a < CTF_N_ACTIONS ? c.reversed_action[a] : 4;
    for (int a = 0; a < CTF_N_ACTIONS; ++a)
        if (c.reversed_action[a] >= CTF_N_ACTIONS) return fail(CTF_ERR_INVALID, "reversed_action out of range");
    // metadata slot order (:1050-1067): self, team-mates in id order without self, opponents in id order
    for (int a = 0; a < N; ++a) {
        int q = 0;
        unsigned int row = 0xFFFFFFFFu;
        auto put = [&](int agent) { row = (row & ~(0xFu << (4 * q))) | ((unsigned)agent << (4 * q)); ++q; };
        put(a);
        const int team = c.agent_team[a];
        for (int j = 0; j < c.n_opponents[1 - team]; ++j)
            if (c.opponents[1 - team][j] != a && q < N) put(c.opponents[1 - team][j]);
        for (int j = 0; j < c.n_opponents[team]; ++j)
            if (q < N) put(c.opponents[team][j]);
        P.meta_row[a] = row;
    }
    // the same layout as one code per element (write_meta's float4 path)
    {
        unsigned char codes[256];
        memset(codes, kMetaZero, sizeof(codes));
        const int M = 6 + 2 * N;
        for (int a = 0; a < N; ++a) {
            unsigned char* row = codes + a * M;
            row[0] = kMetaPct;
            row[1] = c.agent_team[a] == 0 ? kMetaRatio0 : kMetaRatio1;
            for (int t = 0; t < 4; ++t) row[2 + t] = (t == c.agent_type[a]) ? kMetaOne : kMetaZero;
            for (int q = 0; q < N; ++q) {
                const unsigned srcAgent = (P.meta_row[a] >> (4 * q)) & 15u;
                row[6 + 2 * q] = srcAgent == 15u ? kMetaZero : (unsigned char)(kMetaHp + srcAgent);
                row[7 + 2 * q] = srcAgent == 15u ? kMetaZero : (unsigned char)(kMetaFlag + srcAgent);
            }
        }
        for (int k = 0; k < 64; ++k)
            P.meta_codes[k] = codes[4 * k] | (codes[4 * k + 1] << 8) | (codes[4 * k + 2] << 16) | ((unsigned)codes[4 * k + 3] << 24);
    }
    for (int r = 0; r < G; ++r)
        for (int cc = 0; cc < G; ++cc) {
            const uint8_t t = c.grid_template[r * G + cc];
            if (t >= CTF_N_TILE_CODES) return fail(CTF_ERR_INVALID, "grid_template holds an unknown tile code");
            P.grid_template[r * kRow + cc] = t;
        }
    // shared-memory carve-up per warp
    int off = kGridBytes;
    (void)stats_level;
    P.list_off = off;
    off += kMaxList * 4;
    P.bits_off = off;
    off += ((P.bits_words * 4 + 15) / 16) * 16 + 16;
    P.warp_smem_bytes = off;
    return CTF_OK;
}

extern "C" int ctf_create(const ctf_config_t* cfg, int64_t num_envs, int device, uint64_t seed, uint64_t env_id_base,
                          int stats_level, int obs_dtype, ctf_handle_t* out) {
    if (!cfg || !out) return fail(CTF_ERR_INVALID, "null argument");
    if (num_envs < 1) return fail(CTF_ERR_INVALID, "num_envs must be >= 1");
    if (stats_level < 0 || stats_level > 2) return fail(CTF_ERR_INVALID, "stats_level must be 0, 1 or 2");
    if (env_id_base + (uint64_t)num_envs > 0x100000000ull) return fail(CTF_ERR_INVALID, "global env ids must fit 32 bits");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(CTF_ERR_NO_DEVICE, "no CUDA device: the step path has no CPU fallback");
    if (device < 0 || device >= n_dev) return fail(CTF_ERR_INVALID, "device index out of range");
    ctf_env* h = new (std::nothrow) ctf_env();
    if (!h) return fail(CTF_ERR_INVALID, "out of host memory");
    int rc = build_plan(*cfg, stats_level, obs_dtype, h->plan);
    if (rc != CTF_OK) { delete h; return rc; }
    h->cfg = *cfg; h->B = num_envs; h->device = device; h->stats_level = stats_level; h->obs_dtype = obs_dtype;
    h->seed = seed; h->env_id_base = env_id_base;
    const size_t lut_bytes = (CTF_U8_LUT && obs_dtype == CTF_OBS_U8) ? 2048 : 0;   // kLutBytes<uint8_t>
    h->smem_bytes = lut_bytes + (size_t)h->plan.warp_smem_bytes * kWarpsPerCta;
    h->k_step_ctas_per_sm = 0;   // decided below, once the SM count is known
    DeviceGuard guard(device);
    cudaError_t e = guard.err;
    if (e == cudaSuccess) e = cudaMalloc(&h->faults, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(h->faults, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&h->actions_stage, (size_t)num_envs * cfg->n_agents);
    if (e == cudaSuccess) e = cudaMalloc(&h->rewards_stage, (size_t)num_envs * cfg->n_agents * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->dones_stage, (size_t)num_envs);
    h->ws_ctr = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&h->ws_ctr, 128);   // two 32-bit counters (+ the profiling build's cycle totals)
    if (e == cudaSuccess) e = cudaMemset(h->ws_ctr, 0, 128);
    // The persistent kernel is opt-in (CTF_WS=1; shape: one CTA per SM with 8 logic + 16 stream warps unless overridden,
    // tools/ws_sweep.py).  It only ever beat the warp-per-env kernel where the observation stream dwarfs the env logic —
    // float32 8_arena-sized blocks: 0.989 vs 1.002 ms at B = 65536 when it was written, 0.996 vs 1.001 ms after the env
    // logic was trimmed (profiles/r02_ws_final_sweep.jsonl) — and loses everywhere else (7_gridlocked 0.73 vs 0.55 ms,
    // uint8 / bf16 / packed-only outputs: profiles/r02_ws_matrix.log).  Half a per cent does not pay for a second hot
    // kernel, so ctf_step launches k_step unless asked otherwise; both pass the whole parity suite.
    int n_sm = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    h->ws_logic = env_int("CTF_WS_LOGIC", 8);
    h->ws_stream = env_int("CTF_WS_STREAM", 16);
    const int ctas_per_sm = env_int("CTF_WS_CTAS_PER_SM", 1);
    if (h->ws_logic < 1 || h->ws_logic > kWsQ || h->ws_stream < 1 || (h->ws_logic + h->ws_stream) > 32 || ctas_per_sm < 1 ||
        ctas_per_sm > 4) {
        cudaFree(h->faults); cudaFree(h->actions_stage); cudaFree(h->rewards_stage); cudaFree(h->dones_stage); cudaFree(h->ws_ctr);
        delete h;
        return fail(CTF_ERR_INVALID, "CTF_WS_LOGIC / CTF_WS_STREAM / CTF_WS_CTAS_PER_SM out of range");
    }
    h->n_sm = n_sm;
    // Resident CTAs per SM of the warp-per-env kernels.  Registers allow 9 (36 warps); when the step is bound by the
    // observation stream, FEWER concurrent streams write faster (8_arena float32 B = 65536: 9 -> 1.001 ms, 7 -> 0.989,
    // 5 -> 0.980, 4 -> 0.982; 7_gridlocked 0.551 -> 0.536; bf16 0.526 -> 0.516), when it is bound by the env logic they
    // are slower (uint8 0.283 -> 0.347, 0_the_split 0.174 -> 0.229): profiles/r02_ab_residency*.log.  So blocks of
    // >= 6 KB per agent run with 5 CTAs per SM when there are many waves of them, and with the cap in 5..9 that fills
    // the last wave best when there are few (B = 16384: 7).  The cap is applied by padding dynamic shared memory.
    {
        int cap = env_int("CTF_K_STEP_CTAS_PER_SM", -1);
        if (cap < 0) {
            cap = 0;
            const size_t obs_bytes_per_agent = (size_t)h->plan.C * h->plan.GG * (obs_dtype == CTF_OBS_F32 ? 4 : (obs_dtype == CTF_OBS_U8 ? 1 : 2));
            if (obs_bytes_per_agent >= 6000 && n_sm > 0) {
                const double ctas_per_sm_total = (double)((num_envs + kWarpsPerCta - 1) / kWarpsPerCta) / n_sm;
                if (ctas_per_sm_total >= 64.0) {
                    cap = 5;
                } else {
                    double best = -1.0;
                    for (int c = 5; c <= 9; ++c) {   // fill of the last wave; ties go to the smaller cap
                        const double waves = ctas_per_sm_total / c, full = (double)(long long)(waves + 0.999999);
                        const double fill = full > 0 ? waves / full : 0.0;
                        if (fill > best + 1e-9) { best = fill; cap = c; }
                    }
                }
            }
        }
        h->k_step_ctas_per_sm = cap;
        if (cap > 0) {
            const size_t want = (size_t)(227 * 1024) / (size_t)cap - 1024;
            if (want > h->smem_bytes) h->smem_bytes = want;
        }
    }
    h->ws_ctas = n_sm * ctas_per_sm;
    h->ws_smem_bytes = lut_bytes + (size_t)kWsCtlBytes + (size_t)h->plan.warp_smem_bytes * h->ws_logic;
    // below ~4 envs per logic warp the persistent kernel is all ramp-up and tail: use the warp-per-env kernel
    h->ws_min_envs = (long long)env_int("CTF_WS_MIN_ENVS", 4 * h->ws_ctas * h->ws_logic);
    if (env_int("CTF_WS", 0) != 1 || num_envs > 0x7FFFFFFFll) h->ws_min_envs = 0x7FFFFFFFFFFFFFFFll;
    // The attribute is per kernel, not per handle: give every kernel a ceiling that covers any handle's launch (another
    // handle with a smaller footprint must not lower it under this one's), not this handle's own size.
    const int smem = h->smem_bytes > 100 * 1024 ? (int)h->smem_bytes : 100 * 1024;
    const int ws_smem = h->ws_smem_bytes > 200 * 1024 ? (int)h->ws_smem_bytes : 200 * 1024;
#define CTF_SET_SMEM(K, BYTES) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES)
#define CTF_SET_SMEM_T(T)                                                                                       \
    CTF_SET_SMEM((k_step<T, false, false>), smem); CTF_SET_SMEM((k_step<T, true, false>), smem);               \
    CTF_SET_SMEM((k_step<T, false, true>), smem); CTF_SET_SMEM((k_step<T, true, true>), smem);                 \
    CTF_SET_SMEM((k_reset<T, false>), smem); CTF_SET_SMEM((k_reset<T, true>), smem); CTF_SET_SMEM((k_observe<T>), smem); \
    CTF_SET_SMEM((k_step_ws<T, false, false>), ws_smem); CTF_SET_SMEM((k_step_ws<T, true, false>), ws_smem);   \
    CTF_SET_SMEM((k_step_ws<T, false, true>), ws_smem); CTF_SET_SMEM((k_step_ws<T, true, true>), ws_smem);     \
    CTF_SET_SMEM((k_unpack<T>), smem)
    CTF_SET_SMEM_T(float); CTF_SET_SMEM_T(uint8_t); CTF_SET_SMEM_T(__half); CTF_SET_SMEM_T(__nv_bfloat16);
#undef CTF_SET_SMEM_T
#undef CTF_SET_SMEM
    if (e != cudaSuccess) {
        fail(CTF_ERR_CUDA, "ctf_create: %s", cudaGetErrorString(e));
        cudaFree(h->faults); cudaFree(h->actions_stage); cudaFree(h->rewards_stage); cudaFree(h->dones_stage); cudaFree(h->ws_ctr);
        delete h;
        return CTF_ERR_CUDA;
    }
    *out = h;
    return CTF_OK;
}

extern "C" int ctf_destroy(ctf_handle_t h) {
    if (!h) return CTF_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->faults); cudaFree(h->actions_stage); cudaFree(h->rewards_stage); cudaFree(h->dones_stage); cudaFree(h->ws_ctr);
    delete h;
    return CTF_OK;
}

extern "C" int ctf_get_sizes(ctf_handle_t h, ctf_sizes_t* s) {
    if (!h || !s) return fail(CTF_ERR_INVALID, "null argument");
    const DevPlan& P = h->plan;
    const size_t B = (size_t)h->B, elem = h->obs_dtype == CTF_OBS_F32 ? 4 : (h->obs_dtype == CTF_OBS_U8 ? 1 : 2);
    s->grid_stride = kGridBytes;
    s->grid_bytes = B * kGridBytes;
    s->agents_bytes = B * P.N * 8;
    s->envs_bytes = B * 16;
    s->hp_bytes = h->plan.hp_float ? B * P.N * sizeof(double) : 0;
    s->stats_bytes = h->stats_level > 0 ? B * CTF_N_METRICS * P.N * 4 : 0;
    s->visits_bytes = h->stats_level > 1 ? B * P.N * P.GG : 0;
    s->obs_bytes = B * P.E * elem;
    s->obs_bits_bytes = B * P.N * P.wpa * 4;
    s->bits_words_per_agent = (size_t)P.wpa;
    s->meta_bytes = B * P.N * P.M * 4;
    s->rewards_bytes = B * P.N * 4;
    s->dones_bytes = B;
    s->obs_elems_per_env = (size_t)P.E;
    s->meta_elems_per_env = (size_t)P.N * P.M;
    return CTF_OK;
}

static int make_launch(ctf_handle_t h, const ctf_state_t& st, const ctf_outputs_t& out, Launch& L) {
    if (!st.grid || !st.agents || !st.envs) return fail(CTF_ERR_INVALID, "state.grid/agents/envs must not be null");
    if (h->stats_level > 0 && !st.stats) return fail(CTF_ERR_INVALID, "state.stats is null but the handle was created with stats");
    if (h->stats_level > 1 && !st.visits) return fail(CTF_ERR_INVALID, "state.visits is null but the handle was created with visitation maps");
    if ((reinterpret_cast<uintptr_t>(st.grid) & 15) || (reinterpret_cast<uintptr_t>(st.envs) & 15) ||
        (reinterpret_cast<uintptr_t>(st.agents) & 7))
        return fail(CTF_ERR_INVALID, "state buffers must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(out.meta) & 15) return fail(CTF_ERR_INVALID, "outputs.meta must be 16-byte aligned");
    memset(&L, 0, sizeof(L));
    L.grid = st.grid; L.agents = reinterpret_cast<unsigned long long*>(st.agents); L.envs = reinterpret_cast<uint4*>(st.envs);
    if (h->plan.hp_float && (!st.hp || (reinterpret_cast<uintptr_t>(st.hp) & 7)))
        return fail(CTF_ERR_INVALID, "state.hp (8-byte aligned [B][N] doubles) is required when cfg.hp_float is set");
    L.hp = h->plan.hp_float ? st.hp : nullptr;
    L.stats = h->stats_level > 0 ? st.stats : nullptr;
    L.visits = h->stats_level > 1 ? st.visits : nullptr;
    L.obs = out.obs; L.obs_bits = out.obs_bits; L.meta = out.meta; L.rewards = out.rewards; L.dones = out.dones;
    L.faults = h->faults;
    L.B = h->B;
    L.seed_lo = (uint32_t)(h->seed & 0xFFFFFFFFu); L.seed_hi = (uint32_t)(h->seed >> 32);
    L.env_id_base = (uint32_t)h->env_id_base;
    return CTF_OK;
}

static unsigned grid_dim(long long B) { return (unsigned)((B + kWarpsPerCta - 1) / kWarpsPerCta); }

// calls f(T{}) with the observation element type of the handle
template <typename F>
static void with_obs_type(int obs_dtype, F&& f) {
    switch (obs_dtype) {
        case CTF_OBS_F32: f(float{}); break;
        case CTF_OBS_U8: f(uint8_t{}); break;
        case CTF_OBS_F16: f(__half{}); break;
        default: f(__nv_bfloat16{}); break;
    }
}

extern "C" int ctf_reset(ctf_handle_t h, ctf_state_t st, ctf_outputs_t out, int first, void* stream) {
    if (!h) return fail(CTF_ERR_INVALID, "null handle");
    Launch L;
    int rc = make_launch(h, st, out, L);
    if (rc != CTF_OK) return rc;
    L.first_reset = first ? 1 : 0;
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool stats = h->stats_level > 0;
    with_obs_type(h->obs_dtype, [&](auto tag) {
        using T = decltype(tag);
        if (stats) k_reset<T, true><<<grid_dim(h->B), kThreads, h->smem_bytes, s>>>(h->plan, L);
        else k_reset<T, false><<<grid_dim(h->B), kThreads, h->smem_bytes, s>>>(h->plan, L);
    });
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

static int launch_step(ctf_handle_t h, const Launch& L, cudaStream_t s) {
    const bool stats = h->stats_level > 0;
    const bool ws = h->B >= h->ws_min_envs;
    const bool hpf = h->plan.hp_float != 0;
    with_obs_type(h->obs_dtype, [&](auto tag) {
        using T = decltype(tag);
        auto launch = [&](auto stats_c, auto hpf_c) {
            constexpr bool S = decltype(stats_c)::value, F = decltype(hpf_c)::value;
            if (ws) {
                const unsigned threads = (unsigned)(h->ws_logic + h->ws_stream) * 32u;
                k_step_ws<T, S, F><<<h->ws_ctas, threads, h->ws_smem_bytes, s>>>(h->plan, L, h->ws_logic, h->ws_stream, h->ws_ctr);
            } else {
                k_step<T, S, F><<<grid_dim(h->B), kThreads, h->smem_bytes, s>>>(h->plan, L);
            }
        };
        if (stats) { if (hpf) launch(std::true_type{}, std::true_type{}); else launch(std::true_type{}, std::false_type{}); }
        else { if (hpf) launch(std::false_type{}, std::true_type{}); else launch(std::false_type{}, std::false_type{}); }
    });
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

extern "C" int ctf_step(ctf_handle_t h, ctf_state_t st, const uint8_t* actions, ctf_outputs_t out, void* stream) {
    if (!h) return fail(CTF_ERR_INVALID, "null handle");
    if (!actions) return fail(CTF_ERR_INVALID, "actions must not be null");
    Launch L;
    int rc = make_launch(h, st, out, L);
    if (rc != CTF_OK) return rc;
    L.actions = actions;
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    return launch_step(h, L, static_cast<cudaStream_t>(stream));
}

extern "C" int ctf_observe(ctf_handle_t h, ctf_state_t st, const uint8_t* reverse_flags, ctf_outputs_t out, void* stream) {
    if (!h) return fail(CTF_ERR_INVALID, "null handle");
    Launch L;
    int rc = make_launch(h, st, out, L);
    if (rc != CTF_OK) return rc;
    if (reverse_flags) {
        uint32_t m = 0x100u;
        for (int i = 0; i < h->plan.N; ++i) m |= (reverse_flags[i] ? 1u : 0u) << i;
        L.rev_override = m;
    }
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    with_obs_type(h->obs_dtype, [&](auto tag) {
        k_observe<decltype(tag)><<<grid_dim(h->B), kThreads, h->smem_bytes, s>>>(h->plan, L);
    });
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

extern "C" int ctf_unpack_obs(ctf_handle_t h, const uint32_t* packed, void* out, int out_dtype, int64_t n_agent_blocks,
                              void* stream) {
    if (!h || !packed || !out) return fail(CTF_ERR_INVALID, "null argument");
    if (out_dtype < CTF_OBS_F32 || out_dtype > CTF_OBS_BF16) return fail(CTF_ERR_INVALID, "unknown obs dtype");
    if (n_agent_blocks < 0) return fail(CTF_ERR_INVALID, "negative block count");
    if (n_agent_blocks == 0) return CTF_OK;
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int nbits = h->plan.C * h->plan.GG, wpa = h->plan.wpa;
    const int elem = out_dtype == CTF_OBS_F32 ? 4 : (out_dtype == CTF_OBS_U8 ? 1 : 2);
    static const int chunk_kb = env_int("CTF_UNPACK_CHUNK_KB", 48);   // 12 .. 384 KB swept: profiles/r02_unpack_chunk_sweep.log
    int chunk = (chunk_kb * 1024) / (nbits * elem);    // ~48 KB of output per CTA
    chunk = chunk < 1 ? 1 : (chunk > 64 ? 64 : chunk);
    while (chunk > 1 && ((size_t)(chunk * nbits) / 8 + (size_t)chunk * wpa * 4) > 40 * 1024) --chunk;   // static shared-memory budget
    const int bits_words = (((chunk * nbits + 128 + 31) / 32 + 1) + 15) / 16 * 16;
    const unsigned grid = (unsigned)((n_agent_blocks + chunk - 1) / chunk);
    size_t smem = ((size_t)bits_words + (size_t)chunk * wpa) * sizeof(uint32_t) + ((CTF_U8_LUT && out_dtype == CTF_OBS_U8) ? 2048 : 0);
    // like k_step: fewer resident CTAs = fewer concurrent write streams (profiles/r02_unpack_residency.log)
    // float32 6 777 -> 6 838 GB/s and bf16 6 600 -> 6 703 at 5 CTAs per SM, uint8 (issue-bound) 6 365 -> 5 239: capped for >= 2-byte elements
    static const int unpack_ctas_env = env_int("CTF_UNPACK_CTAS_PER_SM", -1);
    const int unpack_ctas = unpack_ctas_env >= 0 ? unpack_ctas_env : (elem >= 2 ? 5 : 0);
    if (unpack_ctas > 0) {
        const size_t want = (size_t)(227 * 1024) / (size_t)unpack_ctas - 1024;
        if (want > smem) smem = want;
    }
    with_obs_type(out_dtype, [&](auto tag) {
        using T = decltype(tag);
        k_unpack<T><<<grid, kUnpackThreads, smem, s>>>(packed, static_cast<T*>(out), n_agent_blocks, nbits, wpa, chunk, bits_words);
    });
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

extern "C" int ctf_stats_sum(ctf_handle_t h, ctf_state_t st, int64_t* stats_sum, void* stream) {
    if (!h || !stats_sum) return fail(CTF_ERR_INVALID, "null argument");
    if (h->stats_level == 0 || !st.stats) return fail(CTF_ERR_INVALID, "handle was created without statistics");
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ns = CTF_N_METRICS * h->plan.N;
    CTF_CUDA(cudaMemsetAsync(stats_sum, 0, sizeof(int64_t) * ns, s));
    const int threads = 832;  // 8 * 104: a multiple of ns for N = 8 (and >= ns for every N)
    const unsigned blocks = (unsigned)((h->B + 255) / 256 < 592 ? (h->B + 255) / 256 : 592);
    k_stats_sum<<<blocks ? blocks : 1, threads, 0, s>>>(st.stats, h->B, ns, reinterpret_cast<unsigned long long*>(stats_sum));
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

extern "C" int ctf_get_kernel_info(ctf_handle_t h, ctf_kernel_info_t* out) {
    if (!h || !out) return fail(CTF_ERR_INVALID, "null argument");
    out->persistent = h->B >= h->ws_min_envs ? 1 : 0;
    out->logic_warps = h->ws_logic; out->stream_warps = h->ws_stream; out->ctas = h->ws_ctas;
    out->min_envs_for_persistent = h->ws_min_envs;
    out->warp_per_env_ctas_per_sm = h->k_step_ctas_per_sm;
    return CTF_OK;
}

extern "C" int ctf_take_faults(ctf_handle_t h, void* stream, uint32_t* faults) {
    if (!h || !faults) return fail(CTF_ERR_INVALID, "null argument");
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CTF_CUDA(cudaMemcpyAsync(faults, h->faults, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CTF_CUDA(cudaMemsetAsync(h->faults, 0, sizeof(uint32_t), s));
    CTF_CUDA(cudaStreamSynchronize(s));
    return CTF_OK;
}

// device-visible alias of a pinned (page-locked, UVA-mapped) host buffer, or nullptr for pageable memory
static void* mapped_alias(const void* host_ptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (attr.type == cudaMemoryTypeHost && attr.devicePointer) return attr.devicePointer;
    return nullptr;
}

extern "C" int ctf_step_host(ctf_handle_t h, ctf_state_t st, const uint8_t* actions_host, ctf_outputs_t out,
                             float* rewards_host, uint8_t* dones_host, void* stream) {
    if (!h) return fail(CTF_ERR_INVALID, "null handle");
    if (!actions_host) return fail(CTF_ERR_INVALID, "actions_host must not be null");
    DeviceGuard guard(h->device);
    CTF_CUDA(guard.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)h->B * h->plan.N;
    // Pinned host buffers are read / written by the kernel itself over PCIe (zero-copy): 8 B of actions in and
    // 33 B of rewards + done out per env ride along with the step instead of three serialized copies.
    // Pageable buffers go through the handle's staging buffers.
    uint8_t* a_dev = static_cast<uint8_t*>(mapped_alias(actions_host));
    float* r_dev = rewards_host ? static_cast<float*>(mapped_alias(rewards_host)) : nullptr;
    uint8_t* d_dev = dones_host ? static_cast<uint8_t*>(mapped_alias(dones_host)) : nullptr;
    const bool r_copy = rewards_host && !r_dev, d_copy = dones_host && !d_dev;
    if (r_dev) out.rewards = r_dev;
    else if (!out.rewards) out.rewards = h->rewards_stage;
    if (d_dev) out.dones = d_dev;
    else if (!out.dones) out.dones = h->dones_stage;
    Launch L;
    int rc = make_launch(h, st, out, L);
    if (rc != CTF_OK) return rc;
    if (a_dev) {
        L.actions = a_dev;
    } else {
        CTF_CUDA(cudaMemcpyAsync(h->actions_stage, actions_host, n, cudaMemcpyHostToDevice, s));
        L.actions = h->actions_stage;
    }
    rc = launch_step(h, L, s);
    if (rc != CTF_OK) return rc;
    if (r_copy) CTF_CUDA(cudaMemcpyAsync(rewards_host, out.rewards, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (d_copy) CTF_CUDA(cudaMemcpyAsync(dones_host, out.dones, (size_t)h->B, cudaMemcpyDeviceToHost, s));
    CTF_CUDA(cudaStreamSynchronize(s));
    return CTF_OK;
}

#if CTF_WS_PROFILE
// profiling builds only (tools/ws_sweep.py --profile): cycle totals {fetch, step, wait, build, stream-idle, stream}, then cleared
extern "C" int ctf_debug_ws_profile(ctf_handle_t h, unsigned long long* out6) {
    DeviceGuard guard(h->device);
    CTF_CUDA(cudaDeviceSynchronize());
    CTF_CUDA(cudaMemcpy(out6, reinterpret_cast<unsigned char*>(h->ws_ctr) + 16, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CTF_CUDA(cudaMemset(reinterpret_cast<unsigned char*>(h->ws_ctr) + 16, 0, 6 * sizeof(unsigned long long)));
    return CTF_OK;
}
#endif
