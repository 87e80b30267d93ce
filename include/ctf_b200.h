/*
 * ctf_b200.h — C ABI of the B200-native batched GridworldCtf step path.
 *
 * The reference (g-nightingale/marl-ctf-development) has no FFI layer: its
 * boundary is the duck-typed Python surface of `GridworldCtf`
 * (gridworld_ctf.py:19-52 ctor, :383 reset, :849-918 step, :975-1009
 * standardise_state, :1027-1069 get_env_metadata).  Each entry point below
 * names the reference member(s) it replaces.  The Python host
 * (marl_ctf_development_b200/env.py) binds these with ctypes; nothing in the
 * signatures is a torch type — plain pointers, sizes and a CUDA stream handle
 * passed as void*.
 *
 * Ownership: the caller owns every state and output buffer (device memory,
 * e.g. torch CUDA tensors).  The library owns only the opaque handle.
 * Threading: all calls are asynchronous on the given stream, never
 * synchronise the host (except the *_host entry points, which return when the
 * host buffers are valid) and never fall back to the CPU.
 * Errors: 0 = ok, negative = error; ctf_last_error() returns the message of
 * the last failing call on this thread.
 */
#ifndef CTF_B200_H
#define CTF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTF_ABI_VERSION 4

#define CTF_MAX_AGENTS 8    /* largest AGENT_STARTING_POSITIONS in scenarios.py has 8 entries */
#define CTF_MAX_GRID 16     /* GRID_SIZE <= 16 (shipped maps: 11, 13, 15) */
#define CTF_MAX_CELLS 256
#define CTF_N_TILE_CODES 14 /* gridworld_ctf.py:250-270: 0 open .. 13 team-1 flag */
#define CTF_MAX_CHANNELS 14 /* 1 self plane + at most 13 tile planes */
#define CTF_N_ACTIONS 9     /* gridworld_ctf.py:100-145 */
#define CTF_N_TYPES 4       /* 0 scout, 1 guardian, 2 vaulter, 3 miner (by behaviour) */
#define CTF_N_METRICS 13    /* agent-level counter families, gridworld_ctf.py:456-468 */

/* order of the per-agent counters in the stats buffer ([B][CTF_N_METRICS][N] u32) */
enum ctf_metric {
    CTF_M_TAG_COUNT = 0,
    CTF_M_RESPAWN_TAG_COUNT = 1,
    CTF_M_FLAG_PICKUPS = 2,
    CTF_M_FLAG_CAPTURES = 3,
    CTF_M_FLAG_DISPOSSESSIONS = 4,
    CTF_M_BLOCKS_LAID = 5,
    CTF_M_BLOCKS_MINED = 6,
    CTF_M_BLOCKS_LAID_DIST_OWN_FLAG = 7,
    CTF_M_BLOCKS_LAID_DIST_OPP_FLAG = 8,
    CTF_M_STEPS_DEFENDING_ZONE = 9,
    CTF_M_STEPS_ATTACKING_ZONE = 10,
    CTF_M_STEPS_ADJ_TEAMMATE = 11,
    CTF_M_STEPS_ADJ_OPPONENT = 12
};

/* bits of the device fault word (ctf_take_faults) */
#define CTF_FAULT_BAD_ACTION 1u      /* an action > 8 was passed (KeyError in ACTION_DELTAS, gridworld_ctf.py:710) */
#define CTF_FAULT_RESPAWN_BLOCKED 2u /* a lethal tag found no open cell in the victim's 3x3 spawn window: the
                                        reference raises ValueError from np.random.randint(0) (gridworld_ctf.py:771);
                                        here the victim stays where it is with HP <= 0 and the bit is set */

enum ctf_error {
    CTF_OK = 0,
    CTF_ERR_INVALID = -1, /* bad argument / config */
    CTF_ERR_CUDA = -2,    /* a CUDA runtime call failed */
    CTF_ERR_NO_DEVICE = -3
};

/* element type of the observation buffer; the values are {0, 1}, exact in every type */
enum ctf_obs_dtype { CTF_OBS_F32 = 0, CTF_OBS_U8 = 1, CTF_OBS_F16 = 2, CTF_OBS_BF16 = 3 };

/*
 * Compiled environment description: everything GridworldCtf.__init__ +
 * load_scenario + reset derive from (AGENT_CONFIG, SCENARIO, kwargs), flattened
 * to a POD.  Built on the host by marl_ctf_development_b200/config.py.
 * HP quantities are fixed point: value_q = value * hp_scale (exact for every
 * shipped config; the compiler refuses configs where it is not).
 */
typedef struct ctf_config {
    /* 8-byte members first so that C, CUDA and ctypes agree on the layout */
    uint64_t tag_threshold;     /* #32-bit words w with w/2^32 < TAG_PROBABILITY  (:815) */
    double reward_step;         /* REWARD_STEP (:78) */
    double reward_capture;      /* REWARD_CAPTURE (:77) */
    double reward_tag;          /* REWARD_TAG (:79) */
    double capture_punish;      /* 1.0 * REWARD_CAPTURE * OPP_FLAG_CAPTURE_PUNISHMENT_SCALAR (:964) */
    double win_margin_scalar;   /* WIN_MARGIN_SCALAR (:75) */
    double loss_margin_scalar;  /* LOSS_MARGIN_SCALAR (:76) */

    int32_t grid_size;          /* G */
    int32_t n_agents;           /* N */
    int32_t n_channels;         /* C = 1 + len(TILES_USED)  (:1016) */
    int32_t game_steps;         /* GAME_STEPS */
    int32_t flip_axis;          /* FLIP_AXIS: -1 = None (both axes), 0 rows, 1 cols, 2 anti-diagonal (:1003-1007) */
    int32_t use_adjusted_rewards;
    int32_t home_flag_capture;
    int32_t drop_flag_when_no_hp;
    int32_t hp_scale;
    int32_t heal_q;             /* AGENT_HP_HEALING_PER_STEP */
    int32_t vault_cost_q;       /* VAULT_HP_COST */
    int32_t vault_min_q;        /* VAULT_MIN_HP */
    int32_t zone_distance;      /* DEFENSIVE_ZONE_DISTANCE (:87) */
    int32_t guardian_distance;  /* GUARDIAN_DEFENSE_DISTANCE (:226) */
    int32_t tagging_range;      /* GUARDIAN_TAGGING_RANGE (:227) */
    int32_t max_agent_blocks;   /* MAX_AGENT_BLOCKS (:241) */
    int32_t block_pickup_value; /* BLOCK_PICKUP_VALUE (:85) */
    int32_t reverse_team1_actions; /* 1: step() maps team-1 actions through reversed_action first (:968-973) */
    int32_t hp_max_q[CTF_N_TYPES];         /* AGENT_TYPE_HP */
    int32_t damage_q[CTF_N_TYPES];         /* AGENT_TYPE_DAMAGE */
    int32_t damage_boosted_q[CTF_N_TYPES]; /* AGENT_TYPE_DAMAGE * GUARDIAN_DAMAGE_MULTIPLIER */

    uint8_t agent_team[CTF_MAX_AGENTS];
    uint8_t agent_type[CTF_MAX_AGENTS];
    uint8_t agent_tile[CTF_MAX_AGENTS];    /* AGENT_TILE_MAP (:264) */
    uint8_t start_row[CTF_MAX_AGENTS];
    uint8_t start_col[CTF_MAX_AGENTS];
    uint8_t obs_reverse[CTF_MAX_AGENTS];   /* reverse_grid flag used for agent i's observation (callers: team != 0) */
    uint8_t meta_hp_src[CTF_MAX_AGENTS];   /* agent id whose HP feeds metadata slot i: = type of agent i (:1040-1041 quirk) */
    uint8_t n_opponents[2];                /* len(OPPONENTS[t]) (:391-395) */
    uint8_t opponents[2][CTF_MAX_AGENTS];  /* OPPONENTS[t] in id order */
    uint8_t flag_pos[2][2];                /* FLAG_POSITIONS */
    uint8_t capture_pos[2][2];             /* CAPTURE_POSITIONS */
    uint8_t spawn_pos[2][2];               /* SPAWN_POSITIONS */
    uint8_t flag_tile[2];                  /* FLAG_TILE_MAP: 12, 13 */
    int8_t action_delta[CTF_N_TYPES][CTF_N_ACTIONS][2]; /* ACTION_DELTAS (:100-145) */
    uint8_t reversed_action[CTF_N_ACTIONS + 7];          /* REVERSED_ACTION_MAP[FLIP_AXIS], padded to 16 */
    uint8_t type_action_mask[CTF_N_TYPES];               /* AGENT_TYPE_ACTION_MASK (:218-223) */
    uint8_t chan_lut[2][16];               /* [observer team][tile code] -> channel 1..C-1, 0 = no channel (O1) */
    uint8_t grid_template[CTF_MAX_CELLS];  /* load_scenario() result, row-major G x G */

    /* HP representation.  0: exact fixed point (the *_q fields, hp_scale) — every shipped configuration.  1: IEEE doubles
       (the *_f fields below), for HP / damage / heal / vault quantities that are not dyadic rationals (e.g. heal 0.1):
       the device then performs the reference's Python float operations one by one (gridworld_ctf.py:648-657, 818-824,
       785, 845-846, 1040) and keeps HP in state.hp. */
    int32_t hp_float;
    int32_t reserved0;
    double hp_max_f[CTF_N_TYPES];          /* AGENT_TYPE_HP */
    double damage_f[CTF_N_TYPES];          /* AGENT_TYPE_DAMAGE */
    double damage_boosted_f[CTF_N_TYPES];  /* AGENT_TYPE_DAMAGE * GUARDIAN_DAMAGE_MULTIPLIER (one float multiplication) */
    double heal_f, vault_cost_f, vault_min_f;
} ctf_config_t;

/* Device buffers of the B resident environments (SoA of field groups, env-major). */
typedef struct ctf_state {
    uint8_t* grid;    /* [B][grid_stride] tile codes, grid_stride = ctf_sizes.grid_stride */
    uint64_t* agents; /* [B][N] packed: row | col<<8 | has_flag<<16 | hp_q(int16)<<32 | inventory(uint16)<<48 */
    uint32_t* envs;   /* [B][4]: env_step_count, episode, team_flag_captures[0], team_flag_captures[1] */
    uint32_t* stats;  /* [B][13][N] counters, or NULL when created with stats_level 0 */
    uint8_t* visits;  /* [B][N][G*G] uint8 visitation maps (wrap at 256), or NULL unless stats_level 2 */
    double* hp;       /* [B][N] agent HP as doubles when cfg.hp_float is set (the hp_q field of `agents` is then unused), else NULL */
} ctf_state_t;

/* Device output buffers of one reset()/step(). Any pointer may be NULL to skip that output. */
typedef struct ctf_outputs {
    void* obs;        /* [B][N][C][G][G] float32 (or uint8 / float16 / bfloat16): standardise_state(i, obs_reverse[i]) for every agent */
    uint32_t* obs_bits; /* [B][N][bits_words_per_agent] packed copy of the same observations, 1 bit per element:
                         element c*G*G + p of agent a is bit (c*G*G + p) % 32 of word a*wpa + (c*G*G + p) / 32.
                         For rollout storage (32x smaller than float32); ctf_unpack_obs expands it. */
    float* meta;      /* [B][N][6+2N] float32: get_env_metadata(i) (fp16-rounded values) */
    float* rewards;   /* [B][N] float32 */
    uint8_t* dones;   /* [B] 0/1 */
} ctf_outputs_t;

typedef struct ctf_sizes {
    size_t grid_stride;  /* bytes per env in state.grid (G*G padded to 16) */
    size_t grid_bytes, agents_bytes, envs_bytes, stats_bytes, visits_bytes;
    size_t obs_bytes, meta_bytes, rewards_bytes, dones_bytes;
    size_t obs_elems_per_env, meta_elems_per_env;
    size_t obs_bits_bytes;       /* bytes of outputs.obs_bits */
    size_t bits_words_per_agent; /* ceil(C*G*G / 32) */
    size_t hp_bytes;             /* bytes of state.hp (0 unless cfg.hp_float) */
} ctf_sizes_t;

typedef struct ctf_env* ctf_handle_t;

const char* ctf_last_error(void);
int ctf_abi_version(void);

/* Number of bytes of ctf_config_t this library was built with (layout check for bindings). */
size_t ctf_config_size(void);

/*
 * Replaces GridworldCtf.__init__ (gridworld_ctf.py:19-350) for a batch:
 * validates cfg, selects `device`, keeps (seed, env_id_base) as the counter-RNG
 * key/offset (env b draws with id env_id_base + b, so results do not depend on
 * how envs are sharded over GPUs).  stats_level: 0 none, 1 counters, 2 counters
 * + visitation maps.  obs_dtype: ctf_obs_dtype.
 */
int ctf_create(const ctf_config_t* cfg, int64_t num_envs, int device, uint64_t seed,
               uint64_t env_id_base, int stats_level, int obs_dtype, ctf_handle_t* out);
int ctf_destroy(ctf_handle_t h);

/* Buffer sizes the caller must allocate (replaces get_env_dims, gridworld_ctf.py:1011-1025). */
int ctf_get_sizes(ctf_handle_t h, ctf_sizes_t* out);

/*
 * Replaces GridworldCtf.reset (gridworld_ctf.py:383-477): every env back to the
 * scenario's initial state, episode counter + 1 (first reset -> episode 0 when
 * `first` is non-zero), statistics cleared; writes observations/metadata of
 * the initial state, zero rewards and dones.
 */
int ctf_reset(ctf_handle_t h, ctf_state_t state, ctf_outputs_t out, int first, void* stream);

/*
 * Replaces GridworldCtf.step (gridworld_ctf.py:849-918) followed by
 * standardise_state + get_env_metadata for every agent (ppo.py:66-95,
 * utils.py:528-551).  actions: device uint8 [B][N], values 0..8 (values > 8,
 * which raise KeyError in the reference, set bit 0 of the handle's fault word
 * and are treated as no-op 4).
 */
int ctf_step(ctf_handle_t h, ctf_state_t state, const uint8_t* actions, ctf_outputs_t out, void* stream);

/*
 * Replaces standardise_state(i, reverse_grid) / get_env_metadata(i) called on
 * their own: observations of the current state without stepping.
 * reverse_flags: host uint8 [N] (NULL = cfg.obs_reverse).
 */
int ctf_observe(ctf_handle_t h, ctf_state_t state, const uint8_t* reverse_flags, ctf_outputs_t out, void* stream);

/*
 * Expands packed observations (outputs.obs_bits layout, any number of agent blocks, e.g. a minibatch gathered
 * from a rollout buffer) into out[n_agent_blocks][C][G][G] of ctf_obs_dtype out_dtype — what the policy
 * consumes at PPO update time (ppo.py:298, 440) without ever holding the float32 rollout in memory.
 */
int ctf_unpack_obs(ctf_handle_t h, const uint32_t* packed, void* out, int out_dtype, int64_t n_agent_blocks, void* stream);

/*
 * Sum of the per-env counters over this handle's envs: device int64
 * [13][N] written to stats_sum (replaces the per-episode harvest of
 * env.metrics, metrics_logger.py:137-159); the caller all-reduces it over ranks.
 */
int ctf_stats_sum(ctf_handle_t h, ctf_state_t state, int64_t* stats_sum, void* stream);

/* Reads and clears the device fault word (CTF_FAULT_* bits). Synchronises the stream. */
int ctf_take_faults(ctf_handle_t h, void* stream, uint32_t* faults);

/*
 * Host-buffer variant of ctf_step for callers that keep actions and rewards on
 * the host: actions_host ([B][N] uint8) in, rewards ([B][N] float32) and dones
 * ([B] uint8) out, valid when the call returns.  Pinned (page-locked) buffers
 * are accessed by the kernel directly over PCIe; pageable ones are staged with
 * copies.  Observations and metadata stay on the device in `out`.
 */
int ctf_step_host(ctf_handle_t h, ctf_state_t state, const uint8_t* actions_host, ctf_outputs_t out,
                  float* rewards_host, uint8_t* dones_host, void* stream);

/*
 * Which kernel ctf_step launches for this handle (no reference counterpart; for benchmarks and profiles).
 * persistent = 1: the persistent warp-specialised kernel k_step_ws (ctas CTAs of logic_warps + stream_warps warps,
 * envs handed out in order from a device counter); 0: the warp-per-env kernel k_step.  Both give identical results.
 * One handle must not have two steps in flight on different streams (the persistent kernel's env counter and the
 * host-step staging buffers are per handle).
 */
typedef struct ctf_kernel_info {
    int persistent, logic_warps, stream_warps, ctas;
    int64_t min_envs_for_persistent;
    int warp_per_env_ctas_per_sm; /* resident-CTA cap of k_step / k_reset / k_observe (0: as many as fit) */
    int reserved;
} ctf_kernel_info_t;
int ctf_get_kernel_info(ctf_handle_t h, ctf_kernel_info_t* out);

#ifdef __cplusplus
}
#endif
#endif /* CTF_B200_H */
