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
#include <string.h>

#include <new>

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
    int bits_words;               // words of the per-env observation bit string (incl. 1 pad word)
    int wpa;                      // words per agent of the packed observation output: ceil(C*G*G / 32)
    int game_steps, flip_axis;
    int use_adjusted_rewards, home_flag_capture, drop_flag_when_no_hp, reverse_team1_actions;
    int heal_q, vault_cost_q, vault_min_q;
    int zone_distance, guardian_distance, tagging_range, max_agent_blocks, block_pickup_value;
    int hp_max_q[4], damage_q[4], damage_boosted_q[4];
    int warp_smem_bytes, list_off, bits_off;
    unsigned char team[8], type[8], tile[8], start_r[8], start_c[8], obs_rev[8], meta_hp_src[8];
    signed char my_slot[8];        // index of agent i in OPPONENTS[1 - team(i)], -1 if truncated away
    unsigned char n_opp[2];
    unsigned char flag_pos[2][2], capture_pos[2][2], spawn_pos[2][2], flag_tile[2];
    signed char delta[4][9][2];
    unsigned char rev_action[16];
    unsigned int meta_row[8];        // metadata slots of observer a: nibble q -> agent id (0xF = none)
    unsigned char grid_template[kGridBytes];  // 16-stride rows
};

struct Launch {
    // state
    uint8_t* grid;
    unsigned long long* agents;
    uint4* envs;
    uint32_t* stats;
    uint8_t* visits;
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
__device__ __forceinline__ int flip_cell(const DevPlan& P, int r, int c) {
    const int G1 = P.G - 1;
    int fr, fc;
    switch (P.flip_axis) {
        case -1: fr = G1 - r; fc = G1 - c; break;
        case 0: fr = G1 - r; fc = c; break;
        case 1: fr = r; fc = G1 - c; break;
        default: fr = G1 - c; fc = G1 - r; break;
    }
    return fr * P.G + fc;
}

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
__device__ __forceinline__ void build_obs_bits(const DevPlan& P, const WarpMem& w, uint32_t me, uint32_t rev_mask, int lane) {
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
        const int base = a * CGG;
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

// Streams `nbits` elements (bit e of `bits` -> element e) to `out` with 128-bit stores; `bits` needs one readable
// word past the last one.  Handles any alignment of `out` (scalar head/tail, funnel-shifted body).
template <typename T>
__device__ __forceinline__ void stream_bits(const uint32_t* bits, int nbits, T* __restrict__ out, int lane) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(out) / sizeof(T)) % VEC);
    const int head = mis ? min(VEC - (int)mis, nbits) : 0;
    const int nvec = (nbits - head) / VEC;
    const int tail = nbits - head - nvec * VEC;
    if (lane < head) out[lane] = from_bit<T>((bits[lane >> 5] >> (lane & 31)) & 1u);
    if (lane < tail) {
        const int e = head + nvec * VEC + lane;
        out[e] = from_bit<T>((bits[e >> 5] >> (e & 31)) & 1u);
    }
    uint4* __restrict__ vp = reinterpret_cast<uint4*>(out + head);
    if (head == 0 && sizeof(T) == 4) {
        // aligned float path: vector i is nibble (i & 7) of word (i >> 3)
        const int sh = (lane & 7) * 4;
        const uint32_t* wp = bits + (lane >> 3);
        int i = lane;
#pragma unroll 4
        for (; i < nvec; i += 32, wp += 4) {
            store_vec(vp + i, expand_bits<T>(*wp >> sh));
        }
    } else {
#pragma unroll 2
        for (int i = lane; i < nvec; i += 32) {
            const int o = head + i * VEC;
            const uint32_t lo = bits[o >> 5], hi = bits[(o >> 5) + 1];
            store_vec(vp + i, expand_bits<T>(__funnelshift_r(lo, hi, o & 31)));
        }
    }
}

// Packed copy of the observation block for rollout storage: agent a's C*G*G bits start at word a*wpa
// (32x smaller than float32; ctf_unpack_obs expands it again).
__device__ __forceinline__ void store_packed(const DevPlan& P, const WarpMem& w, uint32_t* __restrict__ out_env, int lane) {
    const int CGG = P.C * P.GG, wpa = P.wpa, total = P.N * wpa;
    for (int i = lane; i < total; i += 32) {
        const int a = i / wpa, j = i - a * wpa;
        const int o = a * CGG + 32 * j;
        uint32_t v = __funnelshift_r(w.bits[o >> 5], w.bits[(o >> 5) + 1], o & 31);
        const int valid = CGG - 32 * j;            // bits of this word that belong to agent a
        if (valid < 32) v &= (1u << valid) - 1u;
        out_env[i] = v;
    }
}

template <typename T>
__device__ __forceinline__ void write_obs(const DevPlan& P, const Launch& L, const WarpMem& w, long long env, uint32_t me,
                                          uint32_t rev_mask, int lane) {
    if (!L.obs && !L.obs_bits) return;
    build_obs_bits(P, w, me, rev_mask, lane);
    if (L.obs_bits) store_packed(P, w, L.obs_bits + env * (long long)(P.N * P.wpa), lane);
    if (L.obs) stream_bits<T>(w.bits, P.E, reinterpret_cast<T*>(L.obs) + env * (long long)P.E, lane);
}

__device__ __forceinline__ void write_meta(const DevPlan& P, uint32_t me, int step, int caps0, int caps1,
                                           float* __restrict__ meta_env, int lane) {
    const int N = P.N, M = P.M;
    const int li = lane & 7;
    // hp8[i] = uint8(agent_hp[TYPE_i as agent id] / AGENT_TYPE_HP[TYPE_i])  (:1039-1041); lane i keeps hp8 | has_flag<<8
    const uint32_t src = __shfl_sync(kFull, me, P.meta_hp_src[li]);
    const uint32_t pair = (uint32_t)((ag_hp(src) / P.hp_max_q[P.type[li]]) & 0xFF) | ((uint32_t)ag_flag(me) << 8);
    // the three fp64 quotients that go through float16 (:1035-1036, :1044), one per lane, in a single pass
    const int num = lane == 0 ? step : (lane == 1 ? caps0 + 1 : caps1 + 1);
    const int den = lane == 0 ? P.game_steps : (lane == 1 ? caps1 + 1 : caps0 + 1);
    const float quot = __half2float(__double2half((double)num / (double)den));
    const float pct = __shfl_sync(kFull, quot, 0);
    const float ratio0 = __shfl_sync(kFull, quot, 1), ratio1 = __shfl_sync(kFull, quot, 2);
    const int m = lane;                       // lane m produces element m of each agent's vector
    const int q4 = m >= 6 ? ((m - 6) >> 1) * 4 : 0;
    for (int a = 0; a < N; ++a) {
        const uint32_t srcAgent = (P.meta_row[a] >> q4) & 15u;
        const uint32_t sp = __shfl_sync(kFull, pair, srcAgent & 7u);
        float v;
        if (m >= 6) v = srcAgent == 15u ? 0.0f : (float)((m & 1) ? (sp >> 8) : (sp & 0xFFu));
        else if (m >= 2) v = (m - 2 == P.type[a]) ? 1.0f : 0.0f;
        else v = m == 0 ? pct : (P.team[a] == 0 ? ratio0 : ratio1);
        if (m < M) meta_env[a * M + m] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// State staging
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_state(const DevPlan& P, const Launch& L, const WarpMem& w, long long env,
                                            uint32_t me, int inv, uint4 ev, int lane) {
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

template <bool STATS>
__device__ __forceinline__ void bump(Deltas& d, int metric, int agent, uint32_t by, int lane) {
    if (STATS) {
        if (lane == agent) {
            if (metric < 8) d.lo += by << (4 * metric);
            else d.hi += by << (4 * (metric - 8));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// reset (gridworld_ctf.py:383-477)
// ------------------------------------------------------------------------------------------------
template <typename T, bool STATS>
__global__ void __launch_bounds__(kThreads) k_reset(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (env >= L.B) return;
    const WarpMem w = warp_mem(P, reinterpret_cast<unsigned char*>(smem_raw), warp);

    if (lane < kGridBytes / 16)
        reinterpret_cast<uint4*>(w.grid)[lane] = reinterpret_cast<const uint4*>(P.grid_template)[lane];
    const int li = lane & 7;
    const uint32_t me = pack_agent(P.start_r[li], P.start_c[li], 0, P.hp_max_q[P.type[li]]);
    uint4 ev = make_uint4(0, 0, 0, 0);
    if (!L.first_reset) ev.y = L.envs[env].y + 1;  // episode
    __syncwarp();
    store_state(P, L, w, env, me, 0, ev, lane);
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
    if (L.meta) write_meta(P, me, 0, 0, 0, L.meta + env * (long long)P.N * P.M, lane);
    write_obs<T>(P, L, w, env, me, rev_mask, lane);
}

// ------------------------------------------------------------------------------------------------
// observe only
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_observe(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (env >= L.B) return;
    const WarpMem w = warp_mem(P, reinterpret_cast<unsigned char*>(smem_raw), warp);
    if (lane < kGridBytes / 16)
        reinterpret_cast<uint4*>(w.grid)[lane] = ld_state(reinterpret_cast<const uint4*>(L.grid + env * kGridBytes) + lane);
    const int li = lane & 7;
    uint32_t me = 0;
    if (lane < P.N) {
        const unsigned long long rec = ld_state(L.agents + env * P.N + lane);
        me = pack_agent((int)(rec & 0xFF), (int)((rec >> 8) & 0xFF), (int)((rec >> 16) & 1), (int)(short)(rec >> 32));
    }
    const uint4 ev = ld_state(L.envs + env);
    __syncwarp();
    const uint32_t rev_mask = (L.rev_override & 0x100u) ? (L.rev_override & 0xFFu)
                                                        : __ballot_sync(kFull, lane < P.N && P.obs_rev[li]);
    if (L.meta) write_meta(P, me, (int)ev.x, (int)ev.z, (int)ev.w, L.meta + env * (long long)P.N * P.M, lane);
    write_obs<T>(P, L, w, env, me, rev_mask, lane);
}

// ------------------------------------------------------------------------------------------------
// step (gridworld_ctf.py:849-918) + observations for every agent
// ------------------------------------------------------------------------------------------------
template <typename T, bool STATS>
__global__ void __launch_bounds__(kThreads) k_step(const __grid_constant__ DevPlan P, const __grid_constant__ Launch L) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long env = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (env >= L.B) return;
    const WarpMem w = warp_mem(P, reinterpret_cast<unsigned char*>(smem_raw), warp);
    const int N = P.N;
    const int li = lane & 7;

    // ---- stage state: tile map -> shared memory, agent i -> lane i
    if (lane < kGridBytes / 16)
        reinterpret_cast<uint4*>(w.grid)[lane] = ld_state(reinterpret_cast<const uint4*>(L.grid + env * kGridBytes) + lane);
    uint32_t me = 0;
    int inv = 0;
    int action = 4;
    if (lane < N) {
        const unsigned long long rec = ld_state(L.agents + env * N + lane);
        me = pack_agent((int)(rec & 0xFF), (int)((rec >> 8) & 0xFF), (int)((rec >> 16) & 1), (int)(short)(rec >> 32));
        inv = (int)((rec >> 48) & 0xFFFF);
        action = L.actions[env * N + lane];
    }
    uint4 ev = ld_state(L.envs + env);
    Deltas dl;
    const int my_team = P.team[li], my_type = P.type[li], my_slot = P.my_slot[li];
    const bool bad_action = lane < N && action >= CTF_N_ACTIONS;
    if (__any_sync(kFull, bad_action)) {
        if (lane == 0) atomicOr(L.faults, 1u);
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
        bool cap_now = false;

        // ---------------- act (:700-732)
        const int nr = ar + P.delta[type][act_code][0], nc = ac + P.delta[type][act_code][1];
        if (nr >= 0 && nr < P.G && nc >= 0 && nc < P.G) {     // is_valid_move (:636-641)
            const int target = w.grid[nr * kRow + nc];
            if (target == 0 && (act_code <= 3 || (act_code >= 5 && type == 2 && (ahp - P.vault_cost_q) > P.vault_min_q))) {
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
                    bump<STATS>(dl, CTF_M_FLAG_PICKUPS, a, 1, lane);
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
                    bump<STATS>(dl, CTF_M_FLAG_CAPTURES, a, 1, lane);
                }
                if (act_code >= 5 && type == 2) ahp -= P.vault_cost_q;                               // (:652-657)
            } else if (act_code >= 5 && type == 3 && ainv > 0 && target == 0 &&
                       cheb(nr, nc, P.spawn_pos[team][0], P.spawn_pos[team][1]) > 1 &&
                       cheb(nr, nc, P.spawn_pos[1 - team][0], P.spawn_pos[1 - team][1]) > 1) {
                // add_block (:614-634)
                __syncwarp();
                w.grid[nr * kRow + nc] = 2;
                __syncwarp();
                ainv -= 1;
                if (STATS) {
                    bump<STATS>(dl, CTF_M_BLOCKS_LAID, a, 1, lane);
                    bump<STATS>(dl, CTF_M_BLOCKS_LAID_DIST_OWN_FLAG, a,
                                (uint32_t)cheb(ar, ac, P.capture_pos[team][0], P.capture_pos[team][1]), lane);
                    bump<STATS>(dl, CTF_M_BLOCKS_LAID_DIST_OPP_FLAG, a,
                                (uint32_t)cheb(ar, ac, P.capture_pos[1 - team][0], P.capture_pos[1 - team][1]), lane);
                }
            } else if (act_code < 5 && type == 3 && (target == 2 || target == 3)) {
                // mine_block (:677-690)
                __syncwarp();
                w.grid[nr * kRow + nc] = (target == 2) ? 3 : 0;
                __syncwarp();
                if (target == 3) {
                    if (ainv < P.max_agent_blocks) ainv += P.block_pickup_value;
                    bump<STATS>(dl, CTF_M_BLOCKS_MINED, a, 1, lane);
                }
            }
        }
        if (lane == a) {
            me = pack_agent(ar, ac, aflag, ahp);
            inv = ainv;
            captured = cap_now;
        }

        // ---------------- tagging_logic (:796-837)
        if (P.damage_q[type] > 0) {
            const int dmg = (type == 1 && cheb(ar, ac, P.flag_pos[team][0], P.flag_pos[team][1]) <= P.guardian_distance)
                                ? P.damage_boosted_q[type] : P.damage_q[type];
            const bool is_opp = lane < N && my_team != team && my_slot >= 0;
            const int site = 4 * a + (my_slot < 0 ? 0 : my_slot);
            const uint32_t roll = __shfl_sync(kFull, wd[0], site);
            const uint32_t pick_word = __shfl_sync(kFull, wd[1], site);
            const bool hit = is_opp && (unsigned long long)roll < P.tag_threshold &&
                             cheb(ar, ac, ag_r(me), ag_c(me)) <= P.tagging_range;
            int hp = ag_hp(me);
            if (hit) hp -= dmg;                                                                      // (:818)
            const bool lethal = hit && hp <= 0;                                                      // (:824)
            if (hit) me = (me & 0xFFFFu) | ((uint32_t)(hp & 0xFFFF) << 16);
            const unsigned hits = __ballot_sync(kFull, hit);
            unsigned deaths = __ballot_sync(kFull, lethal);
            if (STATS && hits) bump<STATS>(dl, CTF_M_TAG_COUNT, a, (uint32_t)__popc(hits), lane);
            // lethal hits respawn one after the other in opponent-id order: each changes the next one's window
            while (deaths) {
                const int opp = __ffs(deaths) - 1;
                deaths &= deaths - 1;
                const uint32_t om = __shfl_sync(kFull, me, opp);
                const uint32_t ow = __shfl_sync(kFull, pick_word, opp);
                const int oteam = 1 - team;
                if (ag_flag(om)) bump<STATS>(dl, CTF_M_FLAG_DISPOSSESSIONS, a, 1, lane);
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
                    if (lane == opp) me = pack_agent(rr, rc, 0, P.hp_max_q[P.type[opp]]);
                }
                if (lane == a) tag_reward = true;
                bump<STATS>(dl, CTF_M_RESPAWN_TAG_COUNT, a, 1, lane);
            }
        }

        // ---------------- zonal / proximity metrics (:879-902)
        if (STATS) {
            const int d_own = cheb(ar, ac, P.capture_pos[team][0], P.capture_pos[team][1]);
            const int d_opp = cheb(ar, ac, P.capture_pos[1 - team][0], P.capture_pos[1 - team][1]);
            if (d_own <= P.zone_distance) bump<STATS>(dl, CTF_M_STEPS_DEFENDING_ZONE, a, 1, lane);
            if (d_opp <= P.zone_distance) bump<STATS>(dl, CTF_M_STEPS_ATTACKING_ZONE, a, 1, lane);
            const bool near = lane < N && my_slot >= 0 && cheb(ar, ac, ag_r(me), ag_c(me)) <= 1;
            const unsigned mates = __ballot_sync(kFull, near && my_team == team);   // includes the agent itself
            const unsigned opps = __ballot_sync(kFull, near && my_team != team);
            if (mates) bump<STATS>(dl, CTF_M_STEPS_ADJ_TEAMMATE, a, (uint32_t)__popc(mates), lane);
            if (opps) bump<STATS>(dl, CTF_M_STEPS_ADJ_OPPONENT, a, (uint32_t)__popc(opps), lane);
        }
    }

    // ---- heal_agents (:839-847)
    {
        const int mx = P.hp_max_q[my_type];
        int hp = ag_hp(me);
        if (hp < mx) hp = min(hp + P.heal_q, mx);
        me = (me & 0xFFFFu) | ((uint32_t)(hp & 0xFFFF) << 16);
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
        if (L.rewards && lane < N) L.rewards[env * N + lane] = (float)r;
    }
    if (L.dones && lane == 0) L.dones[env] = done ? 1 : 0;
    ev.z = (uint32_t)caps0;
    ev.w = (uint32_t)caps1;

    // ---- write state back
    __syncwarp();
    store_state(P, L, w, env, me, inv, ev, lane);
    if (STATS) {
        if (lane < N) {   // fire-and-forget reductions: no load latency, nothing held in registers
            uint32_t* sp = L.stats + env * (long long)(CTF_N_METRICS * N) + lane;
#pragma unroll
            for (int m = 0; m < CTF_N_METRICS; ++m) {
                const uint32_t v = ((m < 8 ? dl.lo : dl.hi) >> (4 * (m & 7))) & 15u;
                if (v) atomicAdd(sp + m * N, v);
            }
        }
        if (L.visits && lane < N) {   // update_visitation_map (:911), uint8 wrap
            uint8_t* v = L.visits + (env * N + lane) * (long long)P.GG + ag_r(me) * P.G + ag_c(me);
            *v = (uint8_t)(*v + 1);
        }
    }

    // ---- observations straight into the policy's input buffers
    const uint32_t rev_mask = (L.rev_override & 0x100u) ? (L.rev_override & 0xFFu)
                                                        : __ballot_sync(kFull, lane < N && P.obs_rev[li]);
    if (L.meta) write_meta(P, me, step, caps0, caps1, L.meta + env * (long long)N * P.M, lane);
    write_obs<T>(P, L, w, env, me, rev_mask, lane);
}

// ------------------------------------------------------------------------------------------------
// packed observations -> policy input: one warp per agent block (wpa words -> C*G*G elements)
// ------------------------------------------------------------------------------------------------
constexpr int kUnpackWarps = 8;
template <typename T>
__global__ void __launch_bounds__(kUnpackWarps * 32) k_unpack(const uint32_t* __restrict__ packed, T* __restrict__ out,
                                                             long long n_blocks, int nbits, int wpa) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long blk = (long long)blockIdx.x * kUnpackWarps + warp;
    if (blk >= n_blocks) return;
    uint32_t* bits = reinterpret_cast<uint32_t*>(smem_raw) + warp * (wpa + 4);
    for (int i = lane; i < wpa + 1; i += 32) bits[i] = i < wpa ? __ldg(packed + blk * wpa + i) : 0u;
    __syncwarp();
    stream_bits<T>(bits, nbits, out + blk * (long long)nbits, lane);
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
};

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
    P.bits_words = (P.E + 31) / 32 + 1;
    P.wpa = (c.n_channels * G * G + 31) / 32;
    P.game_steps = c.game_steps; P.flip_axis = c.flip_axis;
    P.use_adjusted_rewards = c.use_adjusted_rewards; P.home_flag_capture = c.home_flag_capture;
    P.drop_flag_when_no_hp = c.drop_flag_when_no_hp; P.reverse_team1_actions = c.reverse_team1_actions;
    P.heal_q = c.heal_q; P.vault_cost_q = c.vault_cost_q; P.vault_min_q = c.vault_min_q;
    P.zone_distance = c.zone_distance; P.guardian_distance = c.guardian_distance; P.tagging_range = c.tagging_range;
    P.max_agent_blocks = c.max_agent_blocks; P.block_pickup_value = c.block_pickup_value;
    for (int t = 0; t < 4; ++t) {
        if (c.hp_max_q[t] <= 0 || c.hp_max_q[t] > 32000) return fail(CTF_ERR_INVALID, "hp_max_q out of int16 range");
        P.hp_max_q[t] = c.hp_max_q[t]; P.damage_q[t] = c.damage_q[t]; P.damage_boosted_q[t] = c.damage_boosted_q[t];
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
    h->smem_bytes = (size_t)h->plan.warp_smem_bytes * kWarpsPerCta;
    h->faults = nullptr; h->actions_stage = nullptr; h->rewards_stage = nullptr; h->dones_stage = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&h->faults, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(h->faults, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&h->actions_stage, (size_t)num_envs * cfg->n_agents);
    if (e == cudaSuccess) e = cudaMalloc(&h->rewards_stage, (size_t)num_envs * cfg->n_agents * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->dones_stage, (size_t)num_envs);
    const int smem = (int)h->smem_bytes;
#define CTF_SET_SMEM(K) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
#define CTF_SET_SMEM_T(T)                                                                                    \
    CTF_SET_SMEM((k_step<T, false>)); CTF_SET_SMEM((k_step<T, true>)); CTF_SET_SMEM((k_reset<T, false>)); \
    CTF_SET_SMEM((k_reset<T, true>)); CTF_SET_SMEM((k_observe<T>))
    CTF_SET_SMEM_T(float); CTF_SET_SMEM_T(uint8_t); CTF_SET_SMEM_T(__half); CTF_SET_SMEM_T(__nv_bfloat16);
#undef CTF_SET_SMEM_T
#undef CTF_SET_SMEM
    if (e != cudaSuccess) {
        fail(CTF_ERR_CUDA, "ctf_create: %s", cudaGetErrorString(e));
        cudaFree(h->faults); cudaFree(h->actions_stage); cudaFree(h->rewards_stage); cudaFree(h->dones_stage);
        delete h;
        return CTF_ERR_CUDA;
    }
    *out = h;
    return CTF_OK;
}

extern "C" int ctf_destroy(ctf_handle_t h) {
    if (!h) return CTF_OK;
    cudaSetDevice(h->device);
    cudaFree(h->faults); cudaFree(h->actions_stage); cudaFree(h->rewards_stage); cudaFree(h->dones_stage);
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
    memset(&L, 0, sizeof(L));
    L.grid = st.grid; L.agents = reinterpret_cast<unsigned long long*>(st.agents); L.envs = reinterpret_cast<uint4*>(st.envs);
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
    CTF_CUDA(cudaSetDevice(h->device));
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
    with_obs_type(h->obs_dtype, [&](auto tag) {
        using T = decltype(tag);
        if (stats) k_step<T, true><<<grid_dim(h->B), kThreads, h->smem_bytes, s>>>(h->plan, L);
        else k_step<T, false><<<grid_dim(h->B), kThreads, h->smem_bytes, s>>>(h->plan, L);
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
    CTF_CUDA(cudaSetDevice(h->device));
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
    CTF_CUDA(cudaSetDevice(h->device));
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
    CTF_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int nbits = h->plan.C * h->plan.GG, wpa = h->plan.wpa;
    const unsigned grid = (unsigned)((n_agent_blocks + kUnpackWarps - 1) / kUnpackWarps);
    const size_t smem = (size_t)kUnpackWarps * (wpa + 4) * sizeof(uint32_t);
    with_obs_type(out_dtype, [&](auto tag) {
        using T = decltype(tag);
        k_unpack<T><<<grid, kUnpackWarps * 32, smem, s>>>(packed, static_cast<T*>(out), n_agent_blocks, nbits, wpa);
    });
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

extern "C" int ctf_stats_sum(ctf_handle_t h, ctf_state_t st, int64_t* stats_sum, void* stream) {
    if (!h || !stats_sum) return fail(CTF_ERR_INVALID, "null argument");
    if (h->stats_level == 0 || !st.stats) return fail(CTF_ERR_INVALID, "handle was created without statistics");
    CTF_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int ns = CTF_N_METRICS * h->plan.N;
    CTF_CUDA(cudaMemsetAsync(stats_sum, 0, sizeof(int64_t) * ns, s));
    const int threads = 832;  // 8 * 104: a multiple of ns for N = 8 (and >= ns for every N)
    const unsigned blocks = (unsigned)((h->B + 255) / 256 < 592 ? (h->B + 255) / 256 : 592);
    k_stats_sum<<<blocks ? blocks : 1, threads, 0, s>>>(st.stats, h->B, ns, reinterpret_cast<unsigned long long*>(stats_sum));
    CTF_CUDA(cudaGetLastError());
    return CTF_OK;
}

extern "C" int ctf_take_faults(ctf_handle_t h, void* stream, uint32_t* faults) {
    if (!h || !faults) return fail(CTF_ERR_INVALID, "null argument");
    CTF_CUDA(cudaSetDevice(h->device));
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
    CTF_CUDA(cudaSetDevice(h->device));
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
