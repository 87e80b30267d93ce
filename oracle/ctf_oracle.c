/*
 * ctf_oracle.c — TEST INFRASTRUCTURE: sequential CPU restatement of the
 * reference's GridworldCtf step path (g-nightingale/marl-ctf-development,
 * gridworld_ctf.py), one function per reference method, each citing the lines it
 * follows.  It is the checker for the CUDA path and the "port" CPU baseline of
 * bench.py; the product (marl_ctf_development_b200/) never links or calls it.
 *
 * Parity pin: tests/test_oracle_vs_reference.py runs this against the imported,
 * unmodified reference with the same injected draws (oracle/ref_shim.py) where
 * /root/reference exists, and tests/test_oracle_golden.py checks it against the
 * committed traces in tests/golden/ that the reference generated.
 *
 * Randomness: Philox4x32-10 site draws, see marl_ctf_development_b200/draws.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ctf_b200.h"

#include <pthread.h>

typedef struct ctf_oracle_env {
    const ctf_config_t* cfg;
    uint64_t seed;
    uint32_t env_id;
    int32_t episode;         /* -1 before the first reset */
    int32_t step;            /* env_step_count */
    int32_t done;
    int32_t caps[2];         /* metrics['team_flag_captures'] */
    uint8_t grid[CTF_MAX_CELLS];
    int32_t row[CTF_MAX_AGENTS], col[CTF_MAX_AGENTS];
    int32_t hp_q[CTF_MAX_AGENTS];
    double hp_f[CTF_MAX_AGENTS]; /* agent_hp as Python floats when cfg->hp_float (non-dyadic HP quantities) */
    uint8_t has_flag[CTF_MAX_AGENTS];
    int32_t inventory[CTF_MAX_AGENTS];
    uint32_t stats[CTF_N_METRICS][CTF_MAX_AGENTS];
    uint8_t visits[CTF_MAX_AGENTS][CTF_MAX_CELLS];
    /* per-step scratch */
    uint32_t words[32][4];
    int32_t capture_current_move;
    double capture_team_current_move[2];
    uint32_t faults;         /* CTF_FAULT_* bits, like the device fault word */
} ctf_oracle_env_t;

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void ctf_oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

static void step_words(ctf_oracle_env_t* e) {
    uint32_t key[2] = {(uint32_t)(e->seed & 0xFFFFFFFFu), (uint32_t)(e->seed >> 32)};
    for (uint32_t site = 0; site < 32; ++site) {
        uint32_t ctr[4] = {e->env_id, (uint32_t)e->episode, (uint32_t)e->step, site};
        philox4x32_10(ctr, key, e->words[site]);
    }
}

/* ------------------------------------------------------------------ helpers */
static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* max_dim_distance_to_xy / agent_distance_to_xy (gridworld_ctf.py:744-759): Chebyshev */
static inline int cheb(int r0, int c0, int r1, int c1) { return imax(iabs(r0 - r1), iabs(c0 - c1)); }

#define CELL(e, r, c) ((e)->grid[(r) * (e)->cfg->grid_size + (c)])

static inline void bump(ctf_oracle_env_t* e, int metric, int agent, uint32_t by) { e->stats[metric][agent] += by; }

/* update_visitation_map (gridworld_ctf.py:479-486); uint8 wraps at 256 (:469) */
static void update_visitation_map(ctf_oracle_env_t* e) {
    const ctf_config_t* c = e->cfg;
    for (int i = 0; i < c->n_agents; ++i) e->visits[i][e->row[i] * c->grid_size + e->col[i]] += 1;
}

/* ------------------------------------------------------------------ reset (gridworld_ctf.py:383-477) */
void ctf_oracle_reset(ctf_oracle_env_t* e) {
    const ctf_config_t* c = e->cfg;
    e->episode += 1;
    e->step = 0;                                   /* :388 */
    e->done = 0;                                   /* :389 */
    memcpy(e->grid, c->grid_template, CTF_MAX_CELLS); /* load_scenario :352-381 */
    for (int i = 0; i < c->n_agents; ++i) {
        e->row[i] = c->start_row[i];               /* :407 */
        e->col[i] = c->start_col[i];
        e->has_flag[i] = 0;                        /* :410 */
        e->hp_q[i] = c->hp_max_q[c->agent_type[i]]; /* :415 */
        e->hp_f[i] = c->hp_max_f[c->agent_type[i]];
        e->inventory[i] = 0;                       /* :418 */
    }
    e->caps[0] = e->caps[1] = 0;
    e->capture_current_move = 0;                   /* :421 */
    e->capture_team_current_move[0] = e->capture_team_current_move[1] = 0.0;
    memset(e->stats, 0, sizeof(e->stats));         /* :425-470 */
    memset(e->visits, 0, sizeof(e->visits));
    update_visitation_map(e);                      /* :473 */
}

size_t ctf_oracle_env_size(void) { return sizeof(ctf_oracle_env_t); }

void ctf_oracle_init(ctf_oracle_env_t* e, const ctf_config_t* cfg, uint64_t seed, uint32_t env_id) {
    memset(e, 0, sizeof(*e));
    e->cfg = cfg;
    e->seed = seed;
    e->env_id = env_id;
    e->episode = -1;
    ctf_oracle_reset(e); /* the reference ctor ends with self.reset() (:350) */
}

/* ------------------------------------------------------------------ respawn (gridworld_ctf.py:761-794) */
static void respawn(ctf_oracle_env_t* e, int agent, uint32_t word) {
    const ctf_config_t* c = e->cfg;
    const int G = c->grid_size;
    const int team = c->agent_team[agent];
    const int x = c->spawn_pos[team][0], y = c->spawn_pos[team][1];
    /* np.where(grid[max(x-1,0):x+2, max(y-1,0):y+2] == OPEN): row-major open cells of the clipped window (:768) */
    int cand_r[9], cand_c[9], k = 0;
    for (int r = imax(x - 1, 0); r < imin(x + 2, G); ++r)
        for (int cc = imax(y - 1, 0); cc < imin(y + 2, G); ++cc)
            if (CELL(e, r, cc) == 0) { cand_r[k] = r; cand_c[k] = cc; ++k; }
    /* np.random.randint(k) (:771): injected pick = (word * k) >> 32. k == 0 raises ValueError in the
       reference; it cannot happen on the shipped maps.  Defined behaviour of this backend (CUDA path and
       this restatement alike): the victim stays where it is with HP <= 0 and CTF_FAULT_RESPAWN_BLOCKED is raised. */
    if (k > 0) {
        int pick = (int)(((uint64_t)word * (uint64_t)k) >> 32);
        int nr = cand_r[pick], nc = cand_c[pick]; /* == (x + di - 1, y + dj - 1) for x, y >= 1 (:775) */
        CELL(e, e->row[agent], e->col[agent]) = 0;          /* :778 */
        CELL(e, nr, nc) = c->agent_tile[agent];             /* :779 */
        int old_r = e->row[agent], old_c = e->col[agent];
        e->row[agent] = nr; e->col[agent] = nc;             /* :782 */
        e->hp_q[agent] = c->hp_max_q[c->agent_type[agent]]; /* :785 */
        e->hp_f[agent] = c->hp_max_f[c->agent_type[agent]];
        if (e->has_flag[agent] == 1) {                      /* :788-794 */
            e->has_flag[agent] = 0;
            if (c->drop_flag_when_no_hp)
                CELL(e, old_r, old_c) = c->flag_tile[1 - team];
            else
                CELL(e, c->flag_pos[1 - team][0], c->flag_pos[1 - team][1]) = c->flag_tile[1 - team];
        }
    } else {
        e->faults |= CTF_FAULT_RESPAWN_BLOCKED;
    }
}

/* ------------------------------------------------------------------ movement_handler (gridworld_ctf.py:569-612) */
static void movement_handler(ctf_oracle_env_t* e, int agent, int nr, int nc) {
    const ctf_config_t* c = e->cfg;
    const int team = c->agent_team[agent];
    if (CELL(e, nr, nc) != 0) return;                       /* :577 */
    CELL(e, e->row[agent], e->col[agent]) = 0;              /* :578 */
    CELL(e, nr, nc) = c->agent_tile[agent];                 /* :579 */
    e->row[agent] = nr; e->col[agent] = nc;
    const int ofr = c->flag_pos[1 - team][0], ofc = c->flag_pos[1 - team][1];
    /* flag pickup (:583-591); AGENT_FLAG_CAPTURE_TYPES is all four types (:238) */
    if (cheb(nr, nc, ofr, ofc) <= 1 && CELL(e, ofr, ofc) == c->flag_tile[1 - team]) {
        e->has_flag[agent] = 1;
        CELL(e, ofr, ofc) = 1; /* BLOCK_TILE left at the flag's home cell (:587) */
        bump(e, CTF_M_FLAG_PICKUPS, agent, 1);
    }
    /* flag capture (:594-610) */
    const int hfr = c->flag_pos[team][0], hfc = c->flag_pos[team][1];
    if (cheb(nr, nc, hfr, hfc) <= 1 && e->has_flag[agent] == 1) {
        if (!c->home_flag_capture || CELL(e, hfr, hfc) == c->flag_tile[team]) {
            e->has_flag[agent] = 0;
            CELL(e, ofr, ofc) = c->flag_tile[1 - team];
            e->caps[team] += 1;
            bump(e, CTF_M_FLAG_CAPTURES, agent, 1);
            e->capture_current_move = 1;
            e->capture_team_current_move[team] = 1.0;
        }
    }
}

/* ------------------------------------------------------------------ act (gridworld_ctf.py:700-732) */
static double act(ctf_oracle_env_t* e, int agent, int action) {
    const ctf_config_t* c = e->cfg;
    const int G = c->grid_size;
    const int type = c->agent_type[agent], team = c->agent_team[agent];
    double reward = 0;
    const int nr = e->row[agent] + c->action_delta[type][action][0]; /* :709-711 */
    const int nc = e->col[agent] + c->action_delta[type][action][1];
    if (nr >= 0 && nr < G && nc >= 0 && nc < G) {                    /* is_valid_move :636-641 */
        const int target = CELL(e, nr, nc);
        /* move_to_open_tile (:643-650) */
        const int can_vault = c->hp_float ? ((e->hp_f[agent] - c->vault_cost_f) > c->vault_min_f)   /* :648-650, Python floats */
                                          : ((e->hp_q[agent] - c->vault_cost_q) > c->vault_min_q);
        if (target == 0 && (action <= 3 || (action >= 5 && type == 2 && can_vault))) {
            movement_handler(e, agent, nr, nc);
            if (action >= 5 && type == 2) {                                   /* update_vaulter_hp :652-657 */
                e->hp_q[agent] -= c->vault_cost_q;
                e->hp_f[agent] -= c->vault_cost_f;
            }
        }
        /* can_add_blocks (:659-667) -> add_block (:614-634) */
        else if (action >= 5 && type == 3 && e->inventory[agent] > 0 && target == 0 &&
                 cheb(nr, nc, c->spawn_pos[team][0], c->spawn_pos[team][1]) > 1 &&
                 cheb(nr, nc, c->spawn_pos[1 - team][0], c->spawn_pos[1 - team][1]) > 1) {
            CELL(e, nr, nc) = 2;
            e->inventory[agent] -= 1;
            int d_own = cheb(e->row[agent], e->col[agent], c->capture_pos[team][0], c->capture_pos[team][1]);
            int d_opp = cheb(e->row[agent], e->col[agent], c->capture_pos[1 - team][0], c->capture_pos[1 - team][1]);
            bump(e, CTF_M_BLOCKS_LAID, agent, 1);
            bump(e, CTF_M_BLOCKS_LAID_DIST_OWN_FLAG, agent, (uint32_t)d_own);
            bump(e, CTF_M_BLOCKS_LAID_DIST_OPP_FLAG, agent, (uint32_t)d_opp);
        }
        /* can_mine_blocks (:669-675) -> mine_block (:677-690) */
        else if (action < 5 && type == 3 && (target == 2 || target == 3)) {
            if (target == 2) {
                CELL(e, nr, nc) = 3;
            } else {
                CELL(e, nr, nc) = 0;
                if (e->inventory[agent] < c->max_agent_blocks) e->inventory[agent] += c->block_pickup_value;
                bump(e, CTF_M_BLOCKS_MINED, agent, 1);
            }
        }
    }
    reward += c->reward_step;                        /* :727 */
    if (e->capture_current_move) {                   /* :728-730 */
        reward += c->reward_capture;
        e->capture_current_move = 0;
    }
    return reward;
}

/* ------------------------------------------------------------------ tagging_logic (gridworld_ctf.py:796-837) */
static double tagging_logic(ctf_oracle_env_t* e, int agent) {
    const ctf_config_t* c = e->cfg;
    const int type = c->agent_type[agent], team = c->agent_team[agent];
    double tagging_reward = 0;
    if (c->hp_float ? (c->damage_f[type] > 0.0) : (c->damage_q[type] > 0)) {   /* :804 */
        int dmg = c->damage_q[type];
        double dmg_f = c->damage_f[type];
        if (cheb(e->row[agent], e->col[agent], c->flag_pos[team][0], c->flag_pos[team][1]) <= c->guardian_distance && type == 1) {
            dmg = c->damage_boosted_q[type];         /* :808-810, :818 */
            dmg_f = c->damage_boosted_f[type];
        }
        for (int j = 0; j < c->n_opponents[team]; ++j) { /* :813, OPPONENTS in id order */
            const int opp = c->opponents[team][j];
            const uint32_t* w = e->words[4 * agent + j];
            /* np.random.rand() < TAG_PROBABILITY is evaluated first, always (:815) */
            if ((uint64_t)w[0] < c->tag_threshold &&
                cheb(e->row[agent], e->col[agent], e->row[opp], e->col[opp]) <= c->tagging_range) {
                e->hp_q[opp] -= dmg;                 /* :818 */
                e->hp_f[opp] -= dmg_f;
                bump(e, CTF_M_TAG_COUNT, agent, 1);
                if (c->hp_float ? (e->hp_f[opp] <= 0.0) : (e->hp_q[opp] <= 0)) {   /* :824 */
                    if (e->has_flag[opp] == 1) bump(e, CTF_M_FLAG_DISPOSSESSIONS, agent, 1);
                    respawn(e, opp, w[1]);           /* :831 */
                    tagging_reward = c->reward_tag;  /* :832 */
                    bump(e, CTF_M_RESPAWN_TAG_COUNT, agent, 1);
                }
            }
        }
    }
    return tagging_reward;
}

/* ------------------------------------------------------------------ step (gridworld_ctf.py:849-918) */
void ctf_oracle_step(ctf_oracle_env_t* e, const uint8_t* actions, float* rewards_out, uint8_t* done_out) {
    const ctf_config_t* c = e->cfg;
    const int N = c->n_agents;
    double rewards[CTF_MAX_AGENTS];
    int order[CTF_MAX_AGENTS];

    e->step += 1;                                                        /* :857 */
    e->capture_team_current_move[0] = e->capture_team_current_move[1] = 0; /* :858 */
    step_words(e);
    for (int i = 0; i < N; ++i) { rewards[i] = 0; order[i] = i; }
    /* dice_roll (:734-742), injected: Fisher-Yates from the identity with word 2 of site i */
    for (int i = N - 1; i >= 1; --i) {
        int j = (int)(((uint64_t)e->words[i][2] * (uint64_t)(i + 1)) >> 32);
        int t = order[i]; order[i] = order[j]; order[j] = t;
    }
    for (int s = 0; s < N; ++s) {                                        /* :861 */
        const int agent = order[s];
        const int team = c->agent_team[agent];
        int action = actions[agent];
        if (action >= CTF_N_ACTIONS) action = 4; /* KeyError in the reference; see ctf_step() in the header */
        if (c->reverse_team1_actions && team == 1) action = c->reversed_action[action]; /* :968-973 done by callers */
        rewards[agent] = act(e, agent, action);                          /* :870 */
        rewards[agent] += tagging_logic(e, agent);                       /* :873 */

        /* zonal metrics (:879-889) */
        int d_own = cheb(e->row[agent], e->col[agent], c->capture_pos[team][0], c->capture_pos[team][1]);
        int d_opp = cheb(e->row[agent], e->col[agent], c->capture_pos[1 - team][0], c->capture_pos[1 - team][1]);
        if (d_own <= c->zone_distance) bump(e, CTF_M_STEPS_DEFENDING_ZONE, agent, 1);
        if (d_opp <= c->zone_distance) bump(e, CTF_M_STEPS_ATTACKING_ZONE, agent, 1);
        /* proximity metrics (:892-902); OPPONENTS[1-team] is the own team and includes the agent itself */
        for (int j = 0; j < c->n_opponents[1 - team]; ++j) {
            int mate = c->opponents[1 - team][j];
            if (cheb(e->row[agent], e->col[agent], e->row[mate], e->col[mate]) <= 1) bump(e, CTF_M_STEPS_ADJ_TEAMMATE, agent, 1);
        }
        for (int j = 0; j < c->n_opponents[team]; ++j) {
            int opp = c->opponents[team][j];
            if (cheb(e->row[agent], e->col[agent], e->row[opp], e->col[opp]) <= 1) bump(e, CTF_M_STEPS_ADJ_OPPONENT, agent, 1);
        }
    }
    /* heal_agents (:839-847); its shuffle does not affect the result */
    for (int i = 0; i < N; ++i) {
        int mx = c->hp_max_q[c->agent_type[i]];
        if (e->hp_q[i] < mx) e->hp_q[i] = imin(e->hp_q[i] + c->heal_q, mx);
        const double mx_f = c->hp_max_f[c->agent_type[i]];                  /* :845-846 with Python floats */
        if (e->hp_f[i] < mx_f) { e->hp_f[i] += c->heal_f; if (!(e->hp_f[i] <= mx_f)) e->hp_f[i] = mx_f; }
    }
    /* get_adjusted_rewards (:957-966) */
    if (c->use_adjusted_rewards)
        for (int i = 0; i < N; ++i)
            rewards[i] -= e->capture_team_current_move[1 - c->agent_team[i]] * c->capture_punish;
    update_visitation_map(e);                                            /* :911 */
    if (e->step == c->game_steps) {                                      /* :914-916 */
        e->done = 1;
        /* get_terminal_rewards (:920-940) */
        int margin = iabs(e->caps[0] - e->caps[1]);
        int winner = e->caps[0] > e->caps[1] ? 0 : (e->caps[0] < e->caps[1] ? 1 : -1);
        if (winner >= 0)
            for (int i = 0; i < N; ++i) {
                if (c->agent_team[i] == winner) rewards[i] += margin * c->win_margin_scalar;
                else rewards[i] -= margin * c->loss_margin_scalar;
            }
    }
    for (int i = 0; i < N; ++i) rewards_out[i] = (float)rewards[i];      /* ppo.py:108 fp32 store */
    *done_out = (uint8_t)e->done;
}

/* ------------------------------------------------------------------ fp64 -> fp16 (numpy float16 store, :1044) */
static uint16_t double_to_half_bits(double d) {
    uint64_t b; memcpy(&b, &d, 8);
    uint16_t sign = (uint16_t)((b >> 48) & 0x8000u);
    int exp = (int)((b >> 52) & 0x7FF);
    uint64_t man = b & 0xFFFFFFFFFFFFFull;
    if (exp == 0x7FF) return (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0));
    if (exp == 0 && man == 0) return sign;
    int e = exp - 1023;
    if (e > 15) return (uint16_t)(sign | 0x7C00u);
    uint64_t sig = man | (1ull << 52);  /* 53-bit significand (denormal doubles underflow to 0 below) */
    int shift;                          /* bits dropped from the 53-bit significand */
    int hexp;
    if (e >= -14) { shift = 42; hexp = e + 15; }
    else { shift = 42 + (-14 - e); hexp = 0; }
    if (shift > 63) return sign;
    uint64_t kept = sig >> shift;
    uint64_t rem = sig & ((1ull << shift) - 1);
    uint64_t half = 1ull << (shift - 1);
    if (rem > half || (rem == half && (kept & 1))) kept += 1;
    /* kept carries the implicit bit for normals: adding hexp-1 to it handles mantissa overflow too */
    uint32_t h = hexp > 0 ? (uint32_t)(((uint32_t)(hexp - 1) << 10) + kept) : (uint32_t)kept;
    if (h >= 0x7C00u) h = 0x7C00u;
    return (uint16_t)(sign | h);
}

static float half_bits_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    int exp = (h >> 10) & 0x1F;
    uint32_t man = h & 0x3FFu;
    float f;
    if (exp == 0) {
        f = ldexpf((float)man, -24);
    } else if (exp == 31) {
        f = man ? NAN : INFINITY;
    } else {
        f = ldexpf((float)(man | 0x400u), exp - 25);
    }
    uint32_t bits; memcpy(&bits, &f, 4);
    bits |= sign;
    memcpy(&f, &bits, 4);
    return f;
}

float ctf_oracle_f64_to_f16_to_f32(double d) { return half_bits_to_float(double_to_half_bits(d)); }

/* ------------------------------------------------------------------ standardise_state (gridworld_ctf.py:975-1009) */
void ctf_oracle_standardise_state(const ctf_oracle_env_t* e, int agent, int reverse_grid, uint8_t* out /* [C][G][G] */) {
    const ctf_config_t* c = e->cfg;
    const int G = c->grid_size, C = c->n_channels, GG = G * G;
    const int team = c->agent_team[agent];
    uint8_t plane[CTF_MAX_CHANNELS][CTF_MAX_CELLS];
    memset(plane, 0, sizeof(plane));
    plane[0][e->row[agent] * G + e->col[agent]] = 1;           /* :998 */
    /* :987-1001 folded: channel of each cell as seen by this team (chan_lut built in config.py) */
    for (int p = 0; p < GG; ++p) {
        int ch = c->chan_lut[team][e->grid[p] & 15];
        if (ch) plane[ch][p] = 1;
    }
    for (int ch = 0; ch < C; ++ch)
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < G; ++j) {
                int si = i, sj = j;
                if (reverse_grid) {                            /* :1003-1007 */
                    switch (c->flip_axis) {
                        case -1: si = G - 1 - i; sj = G - 1 - j; break; /* np.flip(x, None) */
                        case 0:  si = G - 1 - i; break;
                        case 1:  sj = G - 1 - j; break;
                        default: si = G - 1 - j; sj = G - 1 - i; break; /* rot90(x.T, 2) */
                    }
                }
                out[ch * GG + i * G + j] = plane[ch][si * G + sj];
            }
}

/* ------------------------------------------------------------------ get_env_metadata (gridworld_ctf.py:1027-1069) */
void ctf_oracle_metadata(const ctf_oracle_env_t* e, int agent, float* out /* [6+2N] */) {
    const ctf_config_t* c = e->cfg;
    const int N = c->n_agents, M = 6 + 2 * N;
    const int team = c->agent_team[agent];
    double m[6 + 2 * CTF_MAX_AGENTS];
    uint8_t hp8[CTF_MAX_AGENTS];
    for (int i = 0; i < M; ++i) m[i] = 0.0;
    m[0] = (double)e->step / (double)c->game_steps;                         /* :1035 */
    m[1] = (double)(e->caps[team] + 1) / (double)(e->caps[1 - team] + 1);   /* :1036 */
    /* :1039-1041 — HP of the agent whose *id* equals agent i's type, over agent i's max HP, as uint8 */
    for (int i = 0; i < N; ++i)
        hp8[i] = c->hp_float ? (uint8_t)((int)(e->hp_f[c->meta_hp_src[i]] / c->hp_max_f[c->agent_type[i]]) & 0xFF)
                             : (uint8_t)(e->hp_q[c->meta_hp_src[i]] / c->hp_max_q[c->agent_type[i]]);
    m[2 + c->agent_type[agent]] = 1.0;                                      /* :1047 */
    m[6] = hp8[agent];                                                      /* :1050 */
    m[7] = e->has_flag[agent];                                              /* :1051 */
    int idx = 8;
    for (int j = 0; j < c->n_opponents[1 - team]; ++j) {                    /* :1055-1060 */
        int mate = c->opponents[1 - team][j];
        if (mate != agent && idx + 1 < M) { m[idx++] = hp8[mate]; m[idx++] = e->has_flag[mate]; }
    }
    for (int j = 0; j < c->n_opponents[team]; ++j) {                        /* :1063-1067 */
        int opp = c->opponents[team][j];
        if (idx + 1 < M) { m[idx++] = hp8[opp]; m[idx++] = e->has_flag[opp]; }
    }
    for (int i = 0; i < M; ++i) out[i] = ctf_oracle_f64_to_f16_to_f32(m[i]); /* :1044 float16 store, ppo.py:70 */
}

/* Same result as ctf_oracle_standardise_state + float cast, written the way an optimised CPU port would
 * (zero the block, then set one element per non-open cell and the self plane).  Used only by the timed CPU
 * baseline legs so that the reported baseline is not handicapped; tests/test_oracle_golden.py checks it
 * against the literal restatement above. */
void ctf_oracle_standardise_state_f32_fast(const ctf_oracle_env_t* e, int agent, int reverse_grid, float* out) {
    const ctf_config_t* c = e->cfg;
    const int G = c->grid_size, GG = G * G;
    const int team = c->agent_team[agent];
    memset(out, 0, sizeof(float) * (size_t)c->n_channels * GG);
    for (int r = 0; r < G; ++r)
        for (int cc = 0; cc < G; ++cc) {
            const int t = e->grid[r * G + cc];
            const int self = (r == e->row[agent] && cc == e->col[agent]);
            if (t == 0 && !self) continue;
            int dr = r, dc = cc;
            if (reverse_grid) {
                switch (c->flip_axis) {
                    case -1: dr = G - 1 - r; dc = G - 1 - cc; break;
                    case 0:  dr = G - 1 - r; break;
                    case 1:  dc = G - 1 - cc; break;
                    default: dr = G - 1 - cc; dc = G - 1 - r; break;
                }
            }
            const int ch = c->chan_lut[team][t & 15];
            if (ch) out[ch * GG + dr * G + dc] = 1.0f;
            if (self) out[dr * G + dc] = 1.0f;
        }
}

void ctf_oracle_observe_fast(const ctf_oracle_env_t* e, float* obs, float* meta) {
    const ctf_config_t* c = e->cfg;
    const int per_agent = c->n_channels * c->grid_size * c->grid_size, M = 6 + 2 * c->n_agents;
    for (int a = 0; a < c->n_agents; ++a) {
        ctf_oracle_standardise_state_f32_fast(e, a, c->obs_reverse[a], obs + (size_t)a * per_agent);
        ctf_oracle_metadata(e, a, meta + a * M);
    }
}

/* observations for all agents as the callers build them (ppo.py:66-95, utils.py:528-551) */
void ctf_oracle_observe(const ctf_oracle_env_t* e, const uint8_t* reverse_flags, float* obs, uint8_t* obs_u8, float* meta) {
    const ctf_config_t* c = e->cfg;
    const int per_agent = c->n_channels * c->grid_size * c->grid_size, M = 6 + 2 * c->n_agents;
    uint8_t tmp[CTF_MAX_CHANNELS * CTF_MAX_CELLS];
    for (int a = 0; a < c->n_agents; ++a) {
        int rev = reverse_flags ? reverse_flags[a] : c->obs_reverse[a];
        if (obs || obs_u8) {
            ctf_oracle_standardise_state(e, a, rev, tmp);
            if (obs) for (int i = 0; i < per_agent; ++i) obs[a * per_agent + i] = (float)tmp[i];
            if (obs_u8) memcpy(obs_u8 + a * per_agent, tmp, (size_t)per_agent);
        }
        if (meta) ctf_oracle_metadata(e, a, meta + a * M);
    }
}

/* ------------------------------------------------------------------ state access for tests */
void ctf_oracle_get_state(const ctf_oracle_env_t* e, uint8_t* grid, int32_t* pos, int32_t* hp_q, uint8_t* has_flag,
                          int32_t* inventory, int32_t* scalars /* step, episode, caps0, caps1, done */,
                          uint32_t* stats, uint8_t* visits) {
    const ctf_config_t* c = e->cfg;
    const int GG = c->grid_size * c->grid_size, N = c->n_agents;
    if (grid) memcpy(grid, e->grid, (size_t)GG);
    for (int i = 0; i < N; ++i) {
        if (pos) { pos[2 * i] = e->row[i]; pos[2 * i + 1] = e->col[i]; }
        if (hp_q) hp_q[i] = e->hp_q[i];
        if (has_flag) has_flag[i] = e->has_flag[i];
        if (inventory) inventory[i] = e->inventory[i];
    }
    if (scalars) { scalars[0] = e->step; scalars[1] = e->episode; scalars[2] = e->caps[0]; scalars[3] = e->caps[1]; scalars[4] = e->done; }
    if (stats) for (int m = 0; m < CTF_N_METRICS; ++m) for (int i = 0; i < N; ++i) stats[m * N + i] = e->stats[m][i];
    if (visits) for (int i = 0; i < N; ++i) memcpy(visits + i * GG, e->visits[i], (size_t)GG);
}

void ctf_oracle_set_state(ctf_oracle_env_t* e, const uint8_t* grid, const int32_t* pos, const int32_t* hp_q,
                          const uint8_t* has_flag, const int32_t* inventory, const int32_t* scalars) {
    const ctf_config_t* c = e->cfg;
    const int GG = c->grid_size * c->grid_size, N = c->n_agents;
    memcpy(e->grid, grid, (size_t)GG);
    for (int i = 0; i < N; ++i) {
        e->row[i] = pos[2 * i]; e->col[i] = pos[2 * i + 1];
        e->hp_q[i] = hp_q[i]; e->has_flag[i] = has_flag[i]; e->inventory[i] = inventory[i];
    }
    e->step = scalars[0]; e->episode = scalars[1]; e->caps[0] = scalars[2]; e->caps[1] = scalars[3]; e->done = scalars[4];
}

/* ------------------------------------------------------------------ CPU baseline driver (bench.py)
 * Runs n_envs independent envs for `steps` steps each with uniform random actions (splitmix64 per env),
 * doing per step what the reference's callers do: observations + metadata for every agent, then step().
 * with_obs = 0 times step() only.  Returns a checksum so the work cannot be optimised away. */
static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct baseline_job {
    const ctf_config_t* cfg;
    int n_envs, steps, with_obs;
    uint64_t seed;
    int next_env;          /* shared work counter, guarded by lock */
    pthread_mutex_t lock;
    double checksum;
} baseline_job_t;

static void* baseline_worker(void* arg) {
    baseline_job_t* job = (baseline_job_t*)arg;
    const ctf_config_t* cfg = job->cfg;
    const int N = cfg->n_agents;
    const size_t per_env_obs = (size_t)N * cfg->n_channels * cfg->grid_size * cfg->grid_size;
    const size_t per_env_meta = (size_t)N * (6 + 2 * N);
    ctf_oracle_env_t* e = (ctf_oracle_env_t*)malloc(sizeof(ctf_oracle_env_t));
    float* obs = (float*)malloc(per_env_obs * sizeof(float));
    float* meta = (float*)malloc(per_env_meta * sizeof(float));
    double checksum = 0.0;
    for (;;) {
        pthread_mutex_lock(&job->lock);
        int b = job->next_env++;
        pthread_mutex_unlock(&job->lock);
        if (b >= job->n_envs) break;
        uint64_t s = job->seed * 0x9E3779B97F4A7C15ull + (uint64_t)b;
        uint8_t actions[CTF_MAX_AGENTS];
        float rewards[CTF_MAX_AGENTS];
        uint8_t done = 0;
        ctf_oracle_init(e, cfg, job->seed, (uint32_t)b);
        for (int t = 0; t < job->steps; ++t) {
            if (done) ctf_oracle_reset(e);
            if (job->with_obs) {
                ctf_oracle_observe(e, NULL, obs, NULL, meta);
                checksum += obs[(size_t)(t * 7919) % per_env_obs] + meta[0];
            }
            uint64_t r = splitmix64(&s);
            for (int i = 0; i < N; ++i) { actions[i] = (uint8_t)(((r & 0xFF) * 9) >> 8); r >>= 8; }
            ctf_oracle_step(e, actions, rewards, &done);
            checksum += rewards[0];
        }
    }
    free(e); free(obs); free(meta);
    pthread_mutex_lock(&job->lock);
    job->checksum += checksum;
    pthread_mutex_unlock(&job->lock);
    return NULL;
}

double ctf_oracle_run_baseline(const ctf_config_t* cfg, int n_envs, int steps, uint64_t seed, int with_obs, int n_threads) {
    baseline_job_t job;
    job.cfg = cfg; job.n_envs = n_envs; job.steps = steps; job.with_obs = with_obs; job.seed = seed;
    job.next_env = 0; job.checksum = 0.0;
    pthread_mutex_init(&job.lock, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, baseline_worker, &job);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&job.lock);
    return job.checksum;
}

/* ------------------------------------------------------------------ batch API for the parity tests
 * B sequential envs with global ids env_id_base + b; the same entry points as the single-env API,
 * looped (optionally over threads) so that Python does one call per step. */
typedef struct ctf_oracle_batch {
    ctf_config_t cfg;
    int B;
    ctf_oracle_env_t* envs;
} ctf_oracle_batch_t;

ctf_oracle_batch_t* ctf_oracle_batch_create(const ctf_config_t* cfg, int B, uint64_t seed, uint32_t env_id_base) {
    ctf_oracle_batch_t* bt = (ctf_oracle_batch_t*)malloc(sizeof(ctf_oracle_batch_t));
    bt->cfg = *cfg;
    bt->B = B;
    bt->envs = (ctf_oracle_env_t*)malloc(sizeof(ctf_oracle_env_t) * (size_t)B);
    for (int b = 0; b < B; ++b) ctf_oracle_init(&bt->envs[b], &bt->cfg, seed, env_id_base + (uint32_t)b);
    return bt;
}

void ctf_oracle_batch_destroy(ctf_oracle_batch_t* bt) {
    if (!bt) return;
    free(bt->envs);
    free(bt);
}

void ctf_oracle_batch_get_hp_f(const ctf_oracle_batch_t* bt, double* out) {
    for (int b = 0; b < bt->B; ++b)
        for (int i = 0; i < bt->cfg.n_agents; ++i) out[b * bt->cfg.n_agents + i] = bt->envs[b].hp_f[i];
}

void ctf_oracle_batch_set_hp_f(ctf_oracle_batch_t* bt, const double* in) {
    for (int b = 0; b < bt->B; ++b)
        for (int i = 0; i < bt->cfg.n_agents; ++i) bt->envs[b].hp_f[i] = in[b * bt->cfg.n_agents + i];
}

void ctf_oracle_get_hp_f(const ctf_oracle_env_t* e, double* out) { for (int i = 0; i < e->cfg->n_agents; ++i) out[i] = e->hp_f[i]; }

uint32_t ctf_oracle_take_faults(ctf_oracle_env_t* e) { uint32_t f = e->faults; e->faults = 0; return f; }

uint32_t ctf_oracle_batch_take_faults(ctf_oracle_batch_t* bt) {
    uint32_t f = 0;
    for (int b = 0; b < bt->B; ++b) f |= ctf_oracle_take_faults(&bt->envs[b]);
    return f;
}

void ctf_oracle_batch_reset(ctf_oracle_batch_t* bt) {
    for (int b = 0; b < bt->B; ++b) ctf_oracle_reset(&bt->envs[b]);
}

void ctf_oracle_batch_step(ctf_oracle_batch_t* bt, const uint8_t* actions, float* rewards, uint8_t* dones) {
    const int N = bt->cfg.n_agents;
    for (int b = 0; b < bt->B; ++b) ctf_oracle_step(&bt->envs[b], actions + (size_t)b * N, rewards + (size_t)b * N, dones + b);
}

void ctf_oracle_batch_observe(const ctf_oracle_batch_t* bt, const uint8_t* reverse_flags, float* obs, uint8_t* obs_u8, float* meta) {
    const ctf_config_t* c = &bt->cfg;
    const size_t E = (size_t)c->n_agents * c->n_channels * c->grid_size * c->grid_size;
    const size_t NM = (size_t)c->n_agents * (6 + 2 * c->n_agents);
    for (int b = 0; b < bt->B; ++b)
        ctf_oracle_observe(&bt->envs[b], reverse_flags, obs ? obs + b * E : NULL, obs_u8 ? obs_u8 + b * E : NULL,
                           meta ? meta + b * NM : NULL);
}

void ctf_oracle_batch_get_state(const ctf_oracle_batch_t* bt, uint8_t* grid, int32_t* pos, int32_t* hp_q, uint8_t* has_flag,
                                int32_t* inventory, int32_t* scalars, uint32_t* stats, uint8_t* visits) {
    const ctf_config_t* c = &bt->cfg;
    const size_t GG = (size_t)c->grid_size * c->grid_size, N = (size_t)c->n_agents;
    for (int b = 0; b < bt->B; ++b)
        ctf_oracle_get_state(&bt->envs[b], grid ? grid + b * GG : NULL, pos ? pos + b * N * 2 : NULL, hp_q ? hp_q + b * N : NULL,
                             has_flag ? has_flag + b * N : NULL, inventory ? inventory + b * N : NULL,
                             scalars ? scalars + b * 5 : NULL, stats ? stats + b * CTF_N_METRICS * N : NULL,
                             visits ? visits + b * N * GG : NULL);
}

void ctf_oracle_batch_set_state(ctf_oracle_batch_t* bt, const uint8_t* grid, const int32_t* pos, const int32_t* hp_q,
                                const uint8_t* has_flag, const int32_t* inventory, const int32_t* scalars) {
    const ctf_config_t* c = &bt->cfg;
    const size_t GG = (size_t)c->grid_size * c->grid_size, N = (size_t)c->n_agents;
    for (int b = 0; b < bt->B; ++b)
        ctf_oracle_set_state(&bt->envs[b], grid + b * GG, pos + b * N * 2, hp_q + b * N, has_flag + b * N, inventory + b * N,
                             scalars + b * 5);
}

/* Threaded continuation of a persistent batch (bench.py CPU legs): every env advances `steps` steps from
 * its current state with uniform random actions, the callers' per-step work included when with_obs
 * (observations + metadata for every agent, then step), auto-reset after done.  Thread i owns a
 * contiguous slice of envs. */
typedef struct batch_run_job {
    ctf_oracle_batch_t* bt;
    int lo, hi, steps, with_obs;
    uint64_t seed;
    double checksum;
} batch_run_job_t;

static void* batch_run_worker(void* arg) {
    batch_run_job_t* job = (batch_run_job_t*)arg;
    const ctf_config_t* cfg = &job->bt->cfg;
    const int N = cfg->n_agents;
    const size_t per_env_obs = (size_t)N * cfg->n_channels * cfg->grid_size * cfg->grid_size;
    float* obs = (float*)malloc(per_env_obs * sizeof(float));
    float* meta = (float*)malloc((size_t)N * (6 + 2 * N) * sizeof(float));
    double checksum = 0.0;
    for (int b = job->lo; b < job->hi; ++b) {
        ctf_oracle_env_t* e = &job->bt->envs[b];
        uint64_t s = job->seed * 0x9E3779B97F4A7C15ull + ((uint64_t)b << 20) + (uint64_t)e->step + ((uint64_t)e->episode << 40);
        uint8_t actions[CTF_MAX_AGENTS];
        float rewards[CTF_MAX_AGENTS];
        uint8_t done = (uint8_t)e->done;
        for (int t = 0; t < job->steps; ++t) {
            if (done) ctf_oracle_reset(e);
            if (job->with_obs) {
                /* 1: the literal restatement of standardise_state (compare planes, then flip);
                   2: the tuned writer (same bytes, scatter of the non-open cells) */
                if (job->with_obs == 2) ctf_oracle_observe_fast(e, obs, meta);
                else ctf_oracle_observe(e, NULL, obs, NULL, meta);
                checksum += obs[(size_t)(t * 7919) % per_env_obs] + meta[0];
            }
            uint64_t r = splitmix64(&s);
            for (int i = 0; i < N; ++i) { actions[i] = (uint8_t)(((r & 0xFF) * 9) >> 8); r >>= 8; }
            ctf_oracle_step(e, actions, rewards, &done);
            checksum += rewards[0];
        }
    }
    free(obs); free(meta);
    job->checksum = checksum;
    return NULL;
}

double ctf_oracle_batch_run(ctf_oracle_batch_t* bt, int steps, uint64_t seed, int with_obs, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > bt->B) n_threads = bt->B;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    batch_run_job_t* jobs = (batch_run_job_t*)malloc(sizeof(batch_run_job_t) * (size_t)n_threads);
    double checksum = 0.0;
    for (int i = 0; i < n_threads; ++i) {
        jobs[i].bt = bt; jobs[i].steps = steps; jobs[i].with_obs = with_obs; jobs[i].seed = seed; jobs[i].checksum = 0.0;
        jobs[i].lo = (int)((long long)bt->B * i / n_threads);
        jobs[i].hi = (int)((long long)bt->B * (i + 1) / n_threads);
        pthread_create(&th[i], NULL, batch_run_worker, &jobs[i]);
    }
    for (int i = 0; i < n_threads; ++i) { pthread_join(th[i], NULL); checksum += jobs[i].checksum; }
    free(th); free(jobs);
    return checksum;
}
