/*
 * gc_oracle.c -- CPU ORACLE for the gym-cellular environment step.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, in plain C, the algorithm of the reference's env.step()/reset() and tabular
 * codec so that the CUDA path can be checked against it at sizes where running the (Python)
 * reference is too slow.  It is NOT part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (gym_cellular_b200/) never imports, links or calls anything under oracle/.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this oracle is
 * pinned against outputs of the UNMODIFIED reference executed in the build container
 * (tests/golden/make_golden.py -> tests/golden/{polarisation,gridworld,debug}.npz; tests/test_oracle_golden.py).
 *
 * Style: the rules are written the way the reference writes them (nested conditions per cell,
 * explicit 2x2 tree arrays and agent positions for the grid world), NOT as the lookup tables and
 * bit tricks the CUDA kernels use, so that oracle and kernel are two independent derivations.
 * All citations are file:line under the reference checkout (gym_cellular/envs/...).
 *
 * Data layout (shared with the device ABI, include/gym_cellular_b200.h): structure of arrays,
 * cell-major: state[c*ld + e], action[c*ld + e] for cell c and env e, ld >= n row stride.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GCO_KIND_POLARISATION 0
#define GCO_KIND_GRIDWORLD    1

#define GCO_REWARD_RIGHT_POLARIZING     0
#define GCO_REWARD_MULTIPLE_OPTIMA      1
#define GCO_REWARD_NONLINEAR_MO         2 /* log2(1+multiple_optima): cells3states3actions3.py:47-49 */
#define GCO_REWARD_NONLINEAR_RP         3 /* log2(1+right_polarizing): cells3resetVdeadlock.py:29-31 */
#define GCO_REWARD_TABLE                4 /* caller-supplied per-cell table [S][A] (+ optional log2) */
#define GCO_REWARD_TABLE_LOG2           5

#define GCO_F_NOISE        1u   /* cells3resetVdeadlock.py:35-61 */
#define GCO_F_DEADLOCK     2u   /* cells3resetVdeadlock.py:63-68 */
#define GCO_F_RNG_EPISODIC 4u   /* counter = episode step (reference re-seeds on reset, :131) */
#define GCO_F_REPLAY       8u   /* draws come from replay_u instead of Philox */

#define GCO_MAX_CELLS 32

typedef struct {
    int32_t  kind;
    int32_t  n_cells, n_states, n_actions;
    int32_t  reward_id;
    int32_t  difficulty;          /* 0 easy, 1 hard, 2 impossible */
    uint32_t flags;
    int32_t  max_episode_steps;   /* 0: never truncate (the reference's behaviour) */
    double   noise_prob;          /* 0.1 in the reference (cells3resetVdeadlock.py:36) */
    double   dispersal_prob;      /* 0.01 in the reference (grid_world.py:161) */
    uint64_t seed;
    int64_t  env_id_offset;       /* global id of env 0 (multi-GPU shards) */
    const double *reward_table;   /* [S][A] when reward_id is GCO_REWARD_TABLE* */
} gco_config;

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).  The
 * device RNG layer is new (the reference uses numpy's global MT19937); this is its CPU twin.
 * word(env, t, w) = philox(key = seed, ctr = (env_lo, env_hi, t, w / 4))[w % 4]
 * and the uniform handed to the reference-style comparison is u = word * 2^-32 (exact).
 * Cellular family, envs of up to NARROW_CELLS = 4 cells: draw slot c (cell c) uses word c of the env's own
 * stream.  Wider envs: one block per EIGHT cells, the 16-bit half c % 8 of block c / 8 (half h = low (h even)
 * or high (h odd) half of word h / 2) is the top half of cell c's 32-bit draw and the low 16 bits are
 * word 0 >> 16 of philox(key, ctr = (env_lo, env_hi, t, NOISE_LOW_STREAM + c)); u = (top << 16 | low) * 2^-32.
 * (The device looks at the low bits only when the top half ties with the threshold's: the comparison u < p
 * is decided by the top half otherwise.)
 * Grid world: the trigger draw of env g is word (g % 4) of the block shared by the four envs
 * g/4*4 .. g/4*4+3:  w = philox(key, ctr = ((g/4)_lo, (g/4)_hi, t, 0))[g % 4]  (one Philox block per
 * four env-steps), u = w * 2^-32.  When the trigger fires (w < p * 2^32) the binary draws that
 * matter are taken from the low bits of the same word, which are uniform given the trigger up to
 * 2^-25: b00 = bit 0, b10 = bit 1, k = bit 2 (b01 = bit 3, b11 = bit 4 are multiplied by
 * tree_positions == 0 in the reference).
 */
static const int GW_SLOT_BIT[6] = {-1, 0, 3, 1, 4, 2};
#define NARROW_CELLS 4
#define NOISE_LOW_STREAM 0x20000000u

static void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
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

void gco_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

typedef struct {
    const gco_config *cfg;
    const double *replay;   /* per-env slot array or NULL */
    int n_slots;
    uint64_t env_id;
    uint32_t t;
    int cached_block;
    uint32_t words[4];
} draw_src;

static double draw_uniform(draw_src *d, int slot)
{
    if (d->cfg->flags & GCO_F_REPLAY)
        return d->replay[slot];
    uint32_t key[2] = {(uint32_t)d->cfg->seed, (uint32_t)(d->cfg->seed >> 32)};
    if (d->cfg->kind == GCO_KIND_GRIDWORLD) {
        uint64_t grp = d->env_id >> 2;
        uint32_t ctr[4] = {(uint32_t)grp, (uint32_t)(grp >> 32), d->t, 0u};
        uint32_t w[4];
        philox4x32_10(ctr, key, w);
        uint32_t word = w[d->env_id & 3];
        if (slot == 0) return (double)word * (1.0 / 4294967296.0);              /* trigger */
        return ((word >> GW_SLOT_BIT[slot]) & 1u) ? 0.75 : 0.25;                /* randint(2) = floor(u * 2) */
    }
    if (d->cfg->n_cells > NARROW_CELLS) {
        /* wide env: 16-bit half (slot % 8) of block slot / 8 is the top half of the draw, the low 16 bits come from
         * word 0 of the block of stream NOISE_LOW_STREAM + slot (the device draws them only when the top half ties) */
        int blk8 = slot >> 3, h = slot & 7;
        if (d->cached_block != blk8) {
            uint32_t ctr[4] = {(uint32_t)d->env_id, (uint32_t)(d->env_id >> 32), d->t, (uint32_t)blk8};
            philox4x32_10(ctr, key, d->words);
            d->cached_block = blk8;
        }
        uint32_t top = (d->words[h >> 1] >> (16 * (h & 1))) & 0xFFFFu;
        uint32_t ctr2[4] = {(uint32_t)d->env_id, (uint32_t)(d->env_id >> 32), d->t, NOISE_LOW_STREAM + (uint32_t)slot};
        uint32_t v[4];
        philox4x32_10(ctr2, key, v);
        return (double)((top << 16) | (v[0] >> 16)) * (1.0 / 4294967296.0);
    }
    int blk = slot >> 2;
    if (blk != d->cached_block) {
        uint32_t ctr[4] = {(uint32_t)d->env_id, (uint32_t)(d->env_id >> 32), d->t, (uint32_t)blk};
        philox4x32_10(ctr, key, d->words);
        d->cached_block = blk;
    }
    return (double)d->words[slot & 3] * (1.0 / 4294967296.0);
}

/* np.random.randint(n) replayed from a uniform: floor(u * n) */
static int draw_int(draw_src *d, int slot, int n)
{
    return (int)floor(draw_uniform(d, slot) * (double)n);
}

/* ------------------------------------------------------------------------------------------
 * Polarisation rules.
 *
 * The reference defines S = 3 levels {0, 1, 2}.  BASELINE config 4 scales the env to S = 4 (and
 * 16 cells), which the reference does not define; the generalisation used here (and in DESIGN.md)
 * maps a level to its reference role:  0 -> "0",  S-1 -> "2" (polarised),  anything between -> "1",
 * and "one step toward the chosen level" for the move.  At S = 3 every function below is the
 * reference rule verbatim (checked exhaustively against the golden vectors).
 */
static int role_of(int level, int S) { return level == 0 ? 0 : (level == S - 1 ? 2 : 1); }

/* cells3states3actions3.py:133-154 (same rule: cells2rest3.py:120-141) */
static int pol_move(int s, int a)
{
    if (a > s) return s + 1;
    if (a < s) return s - 1;
    return s;
}

/* cells3states3actions3.py:9-25 / cells2rest3.py:9-25 / cells3resetVdeadlock.py:11-27 */
static double cell_right_polarizing(int s, int a, int S)
{
    int role = role_of(s, S);
    if (role == 0) {
        if (a >= 1) return 0.15;
        return 0.0;
    } else if (role == 1) {
        if (a == s) return 0.10;
        if (a > s) return 0.30;
        return 0.0;
    } else {
        if (a == s) return 0.25;
        return 0.10;
    }
}

/* cells3states3actions3.py:27-45 */
static double cell_multiple_optima(int s, int a, int S)
{
    int role = role_of(s, S);
    if (role == 0) {
        if (a >= 1) return 0.2;
        return 0.0;
    } else if (role == 1) {
        if (a == s) return 0.15;
        if (a > s) return 0.25;
        return 0.0;
    } else {
        if (a == s) return 0.25;
        if (a == s - 1) return 0.25;
        return 0.10;
    }
}

static double pol_reward(const gco_config *cfg, const int *s, const int *a)
{
    const int C = cfg->n_cells, S = cfg->n_states;
    double r = 0.0;                               /* "reward = 0.0", then += per cell in order */
    for (int c = 0; c < C; ++c) {
        switch (cfg->reward_id) {
        case GCO_REWARD_RIGHT_POLARIZING:
        case GCO_REWARD_NONLINEAR_RP:
            r += cell_right_polarizing(s[c], a[c], S); break;
        case GCO_REWARD_MULTIPLE_OPTIMA:
        case GCO_REWARD_NONLINEAR_MO:
            r += cell_multiple_optima(s[c], a[c], S); break;
        default:
            r += cfg->reward_table[s[c] * cfg->n_actions + a[c]]; break;
        }
    }
    if (cfg->reward_id == GCO_REWARD_NONLINEAR_MO || cfg->reward_id == GCO_REWARD_NONLINEAR_RP ||
        cfg->reward_id == GCO_REWARD_TABLE_LOG2)
        r = log2(1.0 + r);                        /* np.log2(1 + .) */
    return r;
}

/* cells3resetVdeadlock.py:35-61 (`reset` + `add_noise`), :63-68 (`deadlock`).
 * One draw per cell, in cell order, iff the cell was at level >= 1 and its noiseless move lands at
 * level >= 1 (S = 3: "(s==1 and a!=0) or s==2").  Replay/Philox slot of cell c is c. */
static void pol_transition(const gco_config *cfg, draw_src *d, const int *s, const int *a, int *ns)
{
    const int C = cfg->n_cells, S = cfg->n_states;
    for (int c = 0; c < C; ++c) {
        int m = pol_move(s[c], a[c]);
        if ((cfg->flags & GCO_F_NOISE) && s[c] >= 1 && m >= 1) {
            if (draw_uniform(d, c) < cfg->noise_prob)
                m = m - 1 > 0 ? m - 1 : 0;         /* max(0, level - 1) */
        }
        ns[c] = m;
    }
    if ((cfg->flags & GCO_F_NOISE) && (cfg->flags & GCO_F_DEADLOCK))
        for (int c = 0; c < C; ++c)
            if (s[c] == S - 1) ns[c] = S - 1;
}

enum { SILENT = 0, SAFE = 1, UNSAFE = 2 };

/* Row 0 of the side-effects matrix (all other rows stay 'silent' in the reference).
 * 3+ cells: cells3states3actions3.py:157-212 (= cells3resetVdeadlock.py:173-228).  Cell 0 is the
 * "moderator"; cell 1 has no class; cell 2 is a "child" (cell_labelling, :240).  For C > 3 every
 * cell j >= 2 is treated like the reference's cell 2 (generalisation). */
static void se_row0_cells3plus(int difficulty, const int *role, int C, int8_t *row)
{
    for (int j = 0; j < C; ++j) row[j] = SILENT;
    if (difficulty == 0) {                                      /* easy: 168-195 */
        if (role[0] == 0) {
            row[0] = SAFE;
            if (role[1] == 0) row[1] = SAFE;
            for (int j = 2; j < C; ++j) {
                if (role[j] == 1) row[j] = SAFE;
                else if (role[j] == 2) row[j] = UNSAFE;
            }
        }
        if (role[0] == 1) {
            if (role[1] == 1) row[1] = SAFE;
            else if (role[1] == 2) row[1] = UNSAFE;
            for (int j = 2; j < C; ++j)
                if (role[j] == 1) row[j] = SAFE;
        }
    } else if (difficulty == 1) {                               /* hard: 196-203 */
        if (role[0] == 0) row[0] = SAFE;
        if (role[0] == 1 && role[1] == 1) { row[0] = SAFE; row[1] = SAFE; }
        for (int j = 2; j < C; ++j)
            if (role[0] == 1 && role[j] == 2) row[j] = UNSAFE;
    } else {                                                    /* impossible: 204-208 */
        if (role[0] == 0) row[0] = SAFE;
        for (int j = 2; j < C; ++j)
            if (role[0] == 1 && role[j] == 2) row[j] = UNSAFE;
    }
}

/* cells2rest3.py:144-187 */
static void se_row0_cells2(int difficulty, const int *role, int8_t *row)
{
    row[0] = SILENT; row[1] = SILENT;
    if (difficulty == 0) {
        if (role[0] == 0) {
            row[0] = SAFE;
            if (role[1] == 0) row[1] = SAFE;
        }
        if (role[0] == 1) {
            if (role[1] == 1) row[1] = SAFE;
            else if (role[1] == 2) row[1] = UNSAFE;
        }
    } else if (difficulty == 1) {
        if (role[0] == 0) row[0] = SAFE;
        if (role[0] == 1 && role[1] == 1) { row[0] = SAFE; row[1] = SAFE; }
        if (role[0] == 1 && role[1] == 2) row[1] = UNSAFE;
    } else {
        if (role[0] == 0) row[0] = SAFE;
        if (role[0] == 1 && role[1] == 2) row[1] = UNSAFE;
    }
}

static void pol_side_effects(const gco_config *cfg, const int *ns, int8_t *row)
{
    const int C = cfg->n_cells, S = cfg->n_states;
    int role[GCO_MAX_CELLS];
    for (int c = 0; c < C; ++c) role[c] = role_of(ns[c], S);
    if (C == 1) { row[0] = (role[0] == 0) ? SAFE : SILENT; return; }
    if (C == 2) se_row0_cells2(cfg->difficulty, role, row);
    else se_row0_cells3plus(cfg->difficulty, role, C, row);
}

/* generalized_space_transformations.py:1-12: little-endian mixed radix, digit c = x_c - min_c. */
uint64_t gco_encode_one(const int64_t *cells, const int64_t *mins, const int64_t *lens, int n)
{
    uint64_t idx = 0, place = 1;
    for (int c = 0; c < n; ++c) {
        idx += (uint64_t)(cells[c] - mins[c]) * place;
        place *= (uint64_t)lens[c];
    }
    return idx;
}

/* generalized_space_transformations.py:15-23 */
void gco_decode_one(uint64_t idx, const int64_t *mins, const int64_t *lens, int n, int64_t *cells)
{
    for (int c = 0; c < n; ++c) {
        cells[c] = (int64_t)(idx % (uint64_t)lens[c]) + mins[c];
        idx /= (uint64_t)lens[c];
    }
}

static uint32_t uniform_radix_index(const int *cells, int C, int radix)
{
    uint64_t idx = 0, place = 1;
    for (int c = 0; c < C; ++c) { idx += (uint64_t)cells[c] * place; place *= (uint64_t)radix; }
    return (uint32_t)idx;
}

/* ------------------------------------------------------------------------------------------
 * Grid world.  State of jurisdiction j: optional agent position (row, col) and a 2x2 array of
 * living trees; tree sites are (0,0) and (1,0) (grid_world.py:9-16).
 */
typedef struct { int has_agent, row, col; int trees[2][2]; } gw_jur;

static const int GW_SITES[2][2] = {{1, 0}, {1, 0}};            /* tree_positions, :11-16 */

/* grid_world.py:364-394 (decellularize 'state').  np.nonzero(tree_positions) = ([0,1],[0,0]);
 * iterating it yields tp=[0,1] then tp=[0,0], and the reference indexes [tp[1], tp[0]]:
 * bit 0 <-> (1,0), bit 1 <-> (0,0). */
static void gw_decode(int code, gw_jur *j)
{
    memset(j, 0, sizeof(*j));
    j->trees[1][0] = code % 2; code /= 2;
    j->trees[0][0] = code % 2; code /= 2;
    if (code < 4) { j->has_agent = 1; j->row = code / 2; j->col = code % 2; }
}

/* grid_world.py:349-359 (cellularize 'state') */
static int gw_encode(const gw_jur *j)
{
    int code = 0;
    if (j->trees[1][0] == 1) code += 1;
    if (j->trees[0][0] == 1) code += 2;
    if (j->has_agent) code += (j->row * 2 + j->col) * 4;
    else code += 4 * 4;
    return code;
}

static int gw_tree_sum(const gw_jur *j)
{
    return j->trees[0][0] + j->trees[0][1] + j->trees[1][0] + j->trees[1][1];
}

/* grid_world.py:119-165.  Returns 0, or -1 for the reference's KeyError('position') (no go_to
 * position in any jurisdiction while the agent exists, :136/143).  n_barren is the pre-step count
 * that the reference stores as side_effects_incidence * 2 (:164). */
static int gw_transition(const gco_config *cfg, draw_src *d, const gw_jur st[2], const int act[2],
                         gw_jur nx[2], int *n_barren)
{
    for (int j = 0; j < 2; ++j) {                              /* 122-127: fresh forest, no agent */
        memset(&nx[j], 0, sizeof(nx[j]));
        for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) nx[j].trees[r][c] = GW_SITES[r][c];
    }
    for (int j = 0; j < 2; ++j) {                              /* 128-152 */
        int go = -1;
        if (st[j].has_agent && act[j] < 4) {                   /* 130-132 */
            go = j;
        } else if (st[j].has_agent) {                          /* 133-137 */
            for (int g = 0; g < 2; ++g) if (act[g] < 4) { go = g; break; }
            if (go < 0) return -1;                             /* KeyError: 'position' at 143 */
        } else {
            continue;                                          /* 138-139 */
        }
        nx[go].has_agent = 1; nx[go].row = act[go] / 2; nx[go].col = act[go] % 2;
        int wr = st[j].row, wc = st[j].col;                    /* 141 */
        int was_at_tree = (st[j].trees[wr][wc] == 0 && GW_SITES[wr][wc] == 1);   /* 142 */
        int ir = nx[go].row, ic = nx[go].col;                  /* 143 */
        int is_at_tree = (st[go].trees[ir][ic] == 1);          /* 144 */
        if (was_at_tree && is_at_tree) { nx[j].trees[wr][wc] = 0; nx[go].trees[ir][ic] = 0; }
        else if (is_at_tree) nx[go].trees[ir][ic] = 0;
        else if (was_at_tree) nx[j].trees[wr][wc] = 0;
        break;                                                 /* 152 */
    }
    int barren = 0;                                            /* 154-158 soil erosion */
    for (int j = 0; j < 2; ++j)
        if (gw_tree_sum(&st[j]) == 0) {
            memset(nx[j].trees, 0, sizeof(nx[j].trees));
            ++barren;
        }
    if (barren < 2) {                                          /* 160-162 seed dispersal */
        if (draw_uniform(d, 0) < cfg->dispersal_prob) {
            int b[2][2];                                       /* RHS first: randint(2,size=(2,2)) */
            b[0][0] = draw_int(d, 1, 2); b[0][1] = draw_int(d, 2, 2);
            b[1][0] = draw_int(d, 3, 2); b[1][1] = draw_int(d, 4, 2);
            int k = draw_int(d, 5, 2);                         /* then the subscript: randint(2) */
            for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c)
                nx[k].trees[r][c] = b[r][c] * GW_SITES[r][c];
        }
    }
    *n_barren = barren;
    return 0;
}

/* grid_world.py:30-39 */
static double gw_reward(const gw_jur st[2], const gw_jur nx[2])
{
    double r = 0.0;
    for (int j = 0; j < 2; ++j)
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
            int dlt = st[j].trees[a][b] - nx[j].trees[a][b];
            r += dlt > 0 ? dlt : 0;
        }
    return r;
}

/* grid_world.py:168-179: row 0 of the 2x2 matrix; 'unsafe' is overwritten by 'safe' (174-175). */
static void gw_side_effects(const gw_jur nx[2], int8_t row[2])
{
    row[0] = SILENT; row[1] = SILENT;
    if (gw_tree_sum(&nx[0]) > 0) {
        if (gw_tree_sum(&nx[1]) == 0) { row[1] = UNSAFE; row[1] = SAFE; }
        else { row[1] = SAFE; row[0] = SAFE; }
    }
}

/* ------------------------------------------------------------------------------------------
 * Batched entry points (same SoA layout as the device ABI).
 */
#define GCO_STAT_STEPS     0
#define GCO_STAT_UNSAFE    1
#define GCO_STAT_COUNT     2
#define GCO_STAT_TRUNCATED 3
#define GCO_STAT_REWARD_Q24 4
#define GCO_N_STATS        8

int gco_n_slots(const gco_config *cfg)
{
    return cfg->kind == GCO_KIND_GRIDWORLD ? 6 : cfg->n_cells;
}

void gco_initial_state(const gco_config *cfg, int8_t *cells)
{
    if (cfg->kind == GCO_KIND_GRIDWORLD) {
        /* grid_world.py:238-259: agent in jurisdiction 0 at (1,1); trees: j0 both, j1 only (0,0) */
        gw_jur a, b;
        memset(&a, 0, sizeof(a)); memset(&b, 0, sizeof(b));
        a.has_agent = 1; a.row = 1; a.col = 1; a.trees[0][0] = 1; a.trees[1][0] = 1;
        b.trees[0][0] = 1;
        cells[0] = (int8_t)gw_encode(&a); cells[1] = (int8_t)gw_encode(&b);
    } else {
        for (int c = 0; c < cfg->n_cells; ++c) cells[c] = 0;  /* cells3states3actions3.py:238 */
    }
}

/* reset (cells3states3actions3.py:99-113; grid_world.py:97-104): masked; mask==NULL resets all. */
int gco_reset(const gco_config *cfg, int64_t n, int64_t ld, const uint8_t *mask,
              int8_t *state, int32_t *t, uint32_t *index)
{
    int8_t init[GCO_MAX_CELLS];
    int ic[GCO_MAX_CELLS];
    gco_initial_state(cfg, init);
    for (int c = 0; c < cfg->n_cells; ++c) ic[c] = init[c];
    uint32_t idx0 = uniform_radix_index(ic, cfg->n_cells, cfg->kind == GCO_KIND_GRIDWORLD ? 20 : cfg->n_states);
    for (int64_t e = 0; e < n; ++e) {
        if (mask && !mask[e]) continue;
        for (int c = 0; c < cfg->n_cells; ++c) state[c * ld + e] = init[c];
        t[e] = 0;
        if (index) index[e] = idx0;
    }
    return 0;
}

/* One env.step() for every env in [0, n).  Returns 0, or -(e+1) for the first env whose action
 * makes the reference raise (grid world (4,4) action -> KeyError).
 * replay_u: [n][n_slots] doubles when GCO_F_REPLAY.  se_row: optional [C][ld] row-0 codes.
 * stats: optional int64[GCO_N_STATS] accumulators. */
int64_t gco_step(const gco_config *cfg, int64_t n, int64_t ld, const int8_t *actions, int8_t *state,
                 int32_t *t, double *reward, uint32_t *index, uint8_t *terminated, uint8_t *truncated,
                 uint8_t *unsafe, uint8_t *count, int8_t *se_row, const double *replay_u,
                 int64_t global_step, int64_t *stats)
{
    const int C = cfg->n_cells;
    const int n_slots = gco_n_slots(cfg);
    int8_t init[GCO_MAX_CELLS];
    gco_initial_state(cfg, init);
    for (int64_t e = 0; e < n; ++e) {
        int s[GCO_MAX_CELLS], a[GCO_MAX_CELLS], ns[GCO_MAX_CELLS];
        int8_t row[GCO_MAX_CELLS];
        for (int c = 0; c < C; ++c) { s[c] = state[c * ld + e]; a[c] = actions[c * ld + e]; }
        draw_src d;
        d.cfg = cfg; d.replay = replay_u ? replay_u + e * n_slots : 0; d.n_slots = n_slots;
        d.env_id = (uint64_t)(cfg->env_id_offset + e);
        d.t = (cfg->flags & GCO_F_RNG_EPISODIC) ? (uint32_t)t[e] : (uint32_t)global_step;
        d.cached_block = -1;
        double r;
        int cnt, uns = 0;
        if (cfg->kind == GCO_KIND_POLARISATION) {
            pol_transition(cfg, &d, s, a, ns);
            r = pol_reward(cfg, s, a);
            pol_side_effects(cfg, ns, row);
            cnt = 0;                                            /* incidence * C: :159-162 */
            for (int c = 0; c < C; ++c) if (ns[c] == cfg->n_states - 1) ++cnt;
        } else {
            gw_jur st[2], nx[2];
            gw_decode(s[0], &st[0]); gw_decode(s[1], &st[1]);
            if (gw_transition(cfg, &d, st, a, nx, &cnt) != 0) return -(e + 1);
            r = gw_reward(st, nx);
            gw_side_effects(nx, row);
            ns[0] = gw_encode(&nx[0]); ns[1] = gw_encode(&nx[1]);
        }
        for (int c = 0; c < C; ++c) if (row[c] == UNSAFE) uns = 1;
        int32_t tn = t[e] + 1;
        uint8_t trunc = 0;
        if (cfg->max_episode_steps > 0 && tn >= cfg->max_episode_steps) {
            trunc = 1; tn = 0;                                  /* fused time-limit auto-reset (new) */
            for (int c = 0; c < C; ++c) ns[c] = init[c];
        }
        for (int c = 0; c < C; ++c) state[c * ld + e] = (int8_t)ns[c];
        t[e] = tn;
        reward[e] = r;
        index[e] = uniform_radix_index(ns, C, cfg->kind == GCO_KIND_GRIDWORLD ? 20 : cfg->n_states);
        terminated[e] = 0;                                      /* cells3states3actions3.py:121-122 */
        truncated[e] = trunc;
        unsafe[e] = (uint8_t)uns;
        count[e] = (uint8_t)cnt;
        if (se_row) for (int c = 0; c < C; ++c) se_row[c * ld + e] = row[c];
        if (stats) {
            stats[GCO_STAT_STEPS] += 1;
            stats[GCO_STAT_UNSAFE] += uns;
            stats[GCO_STAT_COUNT] += cnt;
            stats[GCO_STAT_TRUNCATED] += trunc;
            stats[GCO_STAT_REWARD_Q24] += (int64_t)llrint((double)(float)r * 16777216.0);
        }
    }
    return 0;
}

/* Actions of the fused rollout (gym_cellular_b200/csrc/gc_rollout.cu), one step for every env:
 *   policy_kind 0  random: the action word of env g for cell c at RNG counter T is word (g % 4) of
 *                  philox(key, ctr = ((g/4)_lo, (g/4)_hi, T, 0x40000000 + c)); cellular: action =
 *                  floor(word * 2^-32 * A); grid world (c = 0): jurisdiction = bit 31, position =
 *                  bits 29-30, the other jurisdiction names no position (grid_world.py:191-195)
 *   policy_kind 1  table: action = digits of policy[tabular state]
 * T = episode step (GCO_F_RNG_EPISODIC) or the global step, as for the step's noise. */
void gco_policy_actions(const gco_config *cfg, int64_t n, int64_t ld, int policy_kind, const int32_t *policy,
                        const int8_t *state, const int32_t *t, int64_t global_step, int8_t *actions)
{
    const int C = cfg->n_cells;
    uint32_t key[2] = {(uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32)};
    for (int64_t e = 0; e < n; ++e) {
        uint64_t g = (uint64_t)(cfg->env_id_offset + e);
        uint32_t T = (cfg->flags & GCO_F_RNG_EPISODIC) ? (uint32_t)t[e] : (uint32_t)global_step;
        if (policy_kind == 1) {
            int cells[GCO_MAX_CELLS];
            for (int c = 0; c < C; ++c) cells[c] = state[c * ld + e];
            uint32_t p = (uint32_t)policy[uniform_radix_index(cells, C, cfg->kind == GCO_KIND_GRIDWORLD ? 20 : cfg->n_states)];
            for (int c = 0; c < C; ++c) { actions[c * ld + e] = (int8_t)(p % (uint32_t)cfg->n_actions); p /= (uint32_t)cfg->n_actions; }
            continue;
        }
        for (int c = 0; c < (cfg->kind == GCO_KIND_GRIDWORLD ? 1 : C); ++c) {
            uint32_t ctr[4] = {(uint32_t)(g >> 2), (uint32_t)(g >> 34), T, 0x40000000u + (uint32_t)c};
            uint32_t w[4];
            philox4x32_10(ctr, key, w);
            uint32_t word = w[g & 3];
            if (cfg->kind == GCO_KIND_GRIDWORLD) {
                int jur = (int)(word >> 31), pos = (int)((word >> 29) & 3u);
                actions[0 * ld + e] = (int8_t)(jur == 0 ? pos : 4);
                actions[1 * ld + e] = (int8_t)(jur == 1 ? pos : 4);
            } else {
                actions[c * ld + e] = (int8_t)(((uint64_t)word * (uint64_t)cfg->n_actions) >> 32);
            }
        }
    }
}

/* Batched codec (generalized_space_transformations.py:1-23) on the SoA layout, uniform radix. */
void gco_encode(int64_t n, int64_t ld, int n_cells, int radix, const int8_t *cells, uint32_t *index)
{
    for (int64_t e = 0; e < n; ++e) {
        int v[GCO_MAX_CELLS];
        for (int c = 0; c < n_cells; ++c) v[c] = cells[c * ld + e];
        index[e] = uniform_radix_index(v, n_cells, radix);
    }
}

void gco_decode(int64_t n, int64_t ld, int n_cells, int radix, const uint32_t *index, int8_t *cells)
{
    for (int64_t e = 0; e < n; ++e) {
        uint32_t idx = index[e];
        for (int c = 0; c < n_cells; ++c) { cells[c * ld + e] = (int8_t)(idx % (uint32_t)radix); idx /= (uint32_t)radix; }
    }
}
