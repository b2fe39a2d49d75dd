/*
 * oracle/marlnav_oracle.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C restatement of the reference's batched environment step
 *   /root/reference/marlnav/environment.py:92-107  (Env.step)
 * and everything it calls, in the exact float32 operation order torch's CPU
 * backend uses.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (marlnav_b200/) never does.
 *
 * What is restated, with the reference lines each function follows:
 *   mo_move_agent      environment.py:113-137   _move_agents/_rotate_directions/_rotate
 *   mo_pair            environment.py:271-286   _get_distances (torch.cdist) + _get_angles
 *   mo_observe_env     environment.py:139-180   observations() incl. the cap rule :172-177
 *   mo_reward_env      environment.py:184-269   _rews_and_terms + the five score helpers
 *   mo_step            environment.py:92-107    step(): move, count, observe, reward,
 *                                               masked re-init (:76-90), observe again
 *   mo_triangle_template  utils.py:349-368      TriangleIntitializer agent constants
 *   obstacle box sampling utils.py:390-398      _sample_obstacles (RNG replaced, see below)
 *
 * Two deliberate, documented substitutions (DESIGN.md "Oracle"):
 *   1. cos/sin/acos use marlnav_trig.h (IEEE-only fp32 algorithms, max error
 *      1.38 / 1.48 / 1.12 ulp, exhaustively measured).  torch's MKL build routes
 *      these three ops through closed-source MKL VML, which cannot be restated;
 *      it differs even from torch's own other backend (SLEEF u10, restated in
 *      torch_cpu_math.h) by 1 ulp on 2-9 % of inputs.  Golden vectors are
 *      therefore produced twice (stock torch -> tolerance tests; torch with
 *      sin/cos/acos patched to marlnav_trig -> bit-exact tests), see
 *      tests/golden/make_golden.py.
 *   2. The reference re-samples obstacles from the global CPU mt19937 stream
 *      for the whole batch every step (utils.py:382-394).  Resets here draw
 *      from an addressed Philox4x32-10 stream; the same
 *      draws are injected into the reference through env._init_sampler when
 *      goldens are generated.
 *
 * Build: see oracle/Makefile (gcc -O2 -mfma -ffp-contract=off [-fopenmp]).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "marlnav_trig.h"
#include "torch_cpu_math.h"   /* SLEEF-u10 restatement: kept for mo_trig_sleef (comparison only) */

#define MO_MAX_AGENTS 26      /* torch.cdist leaves the direct formula above 25 columns */
#define MO_MAX_OBSTACLES 64

/* Field-for-field the same layout as include/marlnav_b200.h:marlnav_env_params
 * (defined independently on purpose: the oracle does not include product headers). */
typedef struct {
    int32_t num_envs, num_agents, num_obstacles, episode_len;
    float min_speed, max_speed, min_accel, max_accel;
    float risk_factor, distance_factor, heading_factor, target_factor, soft_factor, bond_factor;
    float ob_risk_dist, ag_risk_dist, ob_coll_dist, ag_coll_dist;
    float agents_min_d, agents_max_d, max_at_prop_d, max_angle_diff;
    float target_radius, cap_distance, bond_sharpness, ideal_dist, init_dist;
    float obst_x_range, obst_x_mean, obst_y_range, obst_y_mean;
} mo_params;

/* Reset source, mirrors include/marlnav_b200.h:marlnav_reset_spec. */
typedef struct {
    const float* tmpl_states;     /* (A,5) if stride 0, else (B,A,5) */
    const float* tmpl_obstacles;  /* NULL -> Philox box sampling; else (O,2)/(B,O,2) */
    const float* tmpl_target;     /* (2) if stride 0, else (B,2) */
    int64_t states_env_stride, obstacles_env_stride, target_env_stride; /* in floats */
    int32_t alias_first_step;     /* MockInitializer aliasing quirk, SURVEY Appendix B-6 */
    int32_t flags;
    uint64_t seed, step_counter, env_id_offset;
    /* TriangleIntitializer with noisy_ags = True (utils.py:25, 381-388): Gaussian position noise
     * and a random heading rotation on top of the agent template; 0 = off */
    int32_t noisy;
    float noise_chol;             /* cholesky(diag(ags_std, ags_std))[0][0] = sqrt(ags_std), utils.py:370-373 */
    float noise_mult;             /* ags_dist, utils.py:382 */
    float angle_range;            /* utils.py:383 */
} mo_reset;

/* ---------------------------------------------------------------- Philox */

static inline void mo_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void mo_philox_kat(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    mo_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

static inline float mo_u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

/* Obstacles of one env for one (seed, step_counter): utils.py:390-398 with the
 * uniform draws addressed by (global env id, step counter, obstacle pair). */
static void mo_sample_obstacles(const mo_params* p, uint64_t seed, uint64_t step_counter,
                                uint64_t env_id, float* obst /* (O,2) */) {
    const int O = p->num_obstacles;
    for (int pair = 0; 2 * pair < O; ++pair) {
        uint32_t r[4];
        mo_philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)step_counter,
                         (uint32_t)pair, (uint32_t)seed,
                         (uint32_t)(seed >> 32) ^ (uint32_t)(step_counter >> 32), r);
        for (int h = 0; h < 2; ++h) {
            int j = 2 * pair + h;
            if (j >= O) break;
            float ux = mo_u01(r[2 * h]), uy = mo_u01(r[2 * h + 1]);
            float sx = p->obst_x_range * (ux - 0.5f);
            float sy = p->obst_y_range * (uy - 0.5f);
            obst[2 * j + 0] = sx + p->obst_x_mean;
            obst[2 * j + 1] = sy + p->obst_y_mean;
        }
    }
}

/* The three draws of agent `agent` of one env for one (seed, step_counter): two standard normals
 * (Box-Muller) and one uniform, Philox-addressed like the obstacles (counter word 3 =
 * 0x40000000 + agent, disjoint from the obstacle pairs' 0..31). */
static void mo_agent_draw(uint64_t seed, uint64_t step_counter, uint64_t env_id, int agent,
                          float* z0, float* z1, float* u) {
    uint32_t r[4];
    mo_philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)step_counter,
                     0x40000000u + (uint32_t)agent, (uint32_t)seed,
                     (uint32_t)(seed >> 32) ^ (uint32_t)(step_counter >> 32), r);
    const float u1 = ((float)(r[0] >> 8) + 1.0f) * 5.9604644775390625e-08f;     /* (0, 1] */
    mt_box_muller(u1, mo_u01(r[1]), z0, z1);
    *u = mo_u01(r[2]);
}

/* utils.py:381-388 for one env, noisy_ags = 1:
 *   pos_noise = ags_dist * (scale_tril @ eps)          :382  (MultivariateNormal.sample = loc + L eps)
 *   angles    = angle_range * (rand - 0.5)             :383
 *   dirs      = [[c,-s],[s,c]] @ dir                   :384, 400-408  (unfused, like environment.py:131-137)
 *   positions = ags_pos + pos_noise                    :385 */
static void mo_sample_agents_noisy(const mo_reset* rs, uint64_t step_counter, uint64_t env_id, int A,
                                   const float* tmpl /* (A,5) */, float* out /* (A,5) */) {
    for (int i = 0; i < A; ++i) {
        float z0, z1, u;
        mo_agent_draw(rs->seed, step_counter, env_id, i, &z0, &z1, &u);
        const float ang = rs->angle_range * (u - 0.5f);
        float sn, cs;
        mt_sincosf(ang, &sn, &cs);
        const float* t = tmpl + 5 * i;
        out[5 * i + 0] = t[0] + (rs->noise_mult * (rs->noise_chol * z0));
        out[5 * i + 1] = t[1] + (rs->noise_mult * (rs->noise_chol * z1));
        out[5 * i + 2] = (cs * t[2]) + ((-sn) * t[3]);
        out[5 * i + 3] = (sn * t[2]) + (cs * t[3]);
        out[5 * i + 4] = t[4];
    }
}

/* whole batch: the sampler's agent states (init, and golden injection into the reference) */
void mo_noisy_agents(const mo_reset* rs, uint64_t step_counter, int64_t B, int A, float* out /* (B,A,5) */) {
    for (int64_t b = 0; b < B; ++b)
        mo_sample_agents_noisy(rs, step_counter, rs->env_id_offset + (uint64_t)b, A,
                               rs->tmpl_states + (size_t)b * rs->states_env_stride, out + (size_t)b * A * 5);
}
/* the raw draws (normals (B,A,2), uniforms (B,A)) -- injected into the REAL reference's
 * TriangleIntitializer when the noisy goldens are generated (tests/golden/make_golden.py) */
void mo_agent_draws(uint64_t seed, uint64_t step_counter, uint64_t env_id_offset, int64_t B, int A,
                    float* normals, float* uniforms) {
    for (int64_t b = 0; b < B; ++b)
        for (int i = 0; i < A; ++i)
            mo_agent_draw(seed, step_counter, env_id_offset + (uint64_t)b, i,
                          normals + ((size_t)b * A + i) * 2, normals + ((size_t)b * A + i) * 2 + 1,
                          uniforms + (size_t)b * A + i);
}
void mo_log01(const float* x, float* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = mt_logf01(x[i]); }

void mo_philox_obstacles(const mo_params* p, uint64_t seed, uint64_t step_counter,
                         uint64_t env_id_offset, float* obstacles /* (B,O,2) */) {
    const int O = p->num_obstacles;
    for (int64_t b = 0; b < p->num_envs; ++b)
        mo_sample_obstacles(p, seed, step_counter, env_id_offset + (uint64_t)b,
                            obstacles + (size_t)b * O * 2);
}

/* ------------------------------------------------------- torch reductions */

/* torch.sum over a contiguous inner dimension of n float32 (ATen SumKernel:
 * scalar row_sum with 4 interleaved accumulators below one 8-lane vector,
 * vectorised rows + sequential tail/lanes from 8 up).  Verified bit-equal to
 * torch.sum for n = 2..100 (tests/test_oracle_math.py).  Valid for n < 512. */
static float mo_torch_row_sum(const float* v, int n) {
    if (n < 8) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int q = n / 4;
        for (int i = 0; i < q; ++i)
            for (int k = 0; k < 4; ++k) acc[k] = acc[k] + v[4 * i + k];
        for (int j = 4 * q; j < n; ++j) acc[0] = acc[0] + v[j];
        return ((acc[0] + acc[1]) + acc[2]) + acc[3];
    }
    int nv = n / 8;
    float acc[4][8];
    memset(acc, 0, sizeof acc);
    int q = nv / 4;
    for (int i = 0; i < q; ++i)
        for (int k = 0; k < 4; ++k)
            for (int l = 0; l < 8; ++l) acc[k][l] = acc[k][l] + v[8 * (4 * i + k) + l];
    for (int j = 4 * q; j < nv; ++j)
        for (int l = 0; l < 8; ++l) acc[0][l] = acc[0][l] + v[8 * j + l];
    float fin = 0.f;
    for (int k = 8 * nv; k < n; ++k) fin = fin + v[k];
    for (int l = 0; l < 8; ++l) {
        float lane = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l];
        fin = fin + lane;
    }
    return fin;
}

float mo_row_sum(const float* v, int n) { return mo_torch_row_sum(v, n); }

/* ------------------------------------------------------------- the step */

static inline float mo_clampf(float x, float lo, float hi) {
    /* torch.clamp: NaN propagates; min(max(x, lo), hi) */
    if (x != x) return x;
    float t = x < lo ? lo : x;
    return t > hi ? hi : t;
}

/* environment.py:113-137 */
static void mo_move_agent(const mo_params* p, float* s /* [x,y,dx,dy,v] */, const float* a) {
    const float PI_F = 3.1415927410125732f;
    float th = mo_clampf(a[0], -PI_F, PI_F);
    float c, sn;
    mt_sincosf(th, &sn, &c);
    float dx = s[2], dy = s[3];
    float ndx = (c * dx) + ((-sn) * dy);
    float ndy = (sn * dx) + (c * dy);
    float acc = mo_clampf(a[1], p->min_accel, p->max_accel);
    float v = mo_clampf(s[4] + acc, p->min_speed, p->max_speed);
    s[2] = ndx; s[3] = ndy; s[4] = v;
    s[0] = s[0] + (ndx * v);
    s[1] = s[1] + (ndy * v);
}

/* environment.py:271-286 (+ cap :172-177): one (agent, object) pair. */
static inline void mo_pair(const mo_params* p, float ox, float oy, float hx, float hy,
                           float px, float py, float* angle, float* dist) {
    float cx = ox - px, cy = oy - py;              /* cdist: own - other */
    float d = sqrtf(fmaf(cy, cy, cx * cx));
    float ex = px - ox, ey = py - oy;              /* _get_angles: other - own */
    float nrm = sqrtf(fmaf(ey, ey, ex * ex));
    float den = nrm > 1e-12f ? nrm : 1e-12f;       /* clamp_min(eps) */
    float nx = ex / den, ny = ey / den;
    float dot = mo_clampf((hx * nx) + (hy * ny), -1.0f, 1.0f);
    float orthx = nx - (dot * hx);
    float sign = orthx > 0.0f ? -1.0f : 1.0f;
    float ang = sign * mt_acosf(dot);
    if (d < p->cap_distance) ang = 0.0f;
    *angle = ang; *dist = d;
}

/* environment.py:139-180.  obs is (A, obs_size) with the Observations field
 * order: [target_angle, target_distance, obstacles_angles(O),
 * obstacles_distances(O), others_angles(A-1), others_distances(A-1)]. */
static void mo_observe_env(const mo_params* p, const float* st, const float* ob,
                           const float* tg, float* obs) {
    const int A = p->num_agents, O = p->num_obstacles;
    const int S = 2 + 2 * O + 2 * (A - 1);
    for (int i = 0; i < A; ++i) {
        const float* s = st + 5 * i;
        float* o = obs + (size_t)S * i;
        mo_pair(p, s[0], s[1], s[2], s[3], tg[0], tg[1], &o[0], &o[1]);
        for (int j = 0; j < O; ++j)
            mo_pair(p, s[0], s[1], s[2], s[3], ob[2 * j], ob[2 * j + 1], &o[2 + j], &o[2 + O + j]);
        int k = 0;
        for (int j = 0; j < A; ++j) {
            if (j == i) continue;
            mo_pair(p, s[0], s[1], s[2], s[3], st[5 * j], st[5 * j + 1],
                    &o[2 + 2 * O + k], &o[2 + 2 * O + (A - 1) + k]);
            ++k;
        }
    }
}

/* environment.py:113-137 on a whole batch (phase-split parity: move only) */
void mo_move(const mo_params* p, float* states, const float* actions) {
    const int A = p->num_agents;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < p->num_envs; ++b)
        for (int i = 0; i < A; ++i)
            mo_move_agent(p, states + ((size_t)b * A + i) * 5, actions + ((size_t)b * A + i) * 2);
}

void mo_observe(const mo_params* p, const float* states, const float* obstacles,
                const float* target, float* obs) {
    const int A = p->num_agents, O = p->num_obstacles;
    const int S = 2 + 2 * O + 2 * (A - 1);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < p->num_envs; ++b)
        mo_observe_env(p, states + (size_t)b * A * 5, obstacles + (size_t)b * O * 2,
                       target + (size_t)b * 2, obs + (size_t)b * A * S);
}

/* environment.py:184-269.  Returns the env reward; writes coll_any / all_in. */
static float mo_reward_env(const mo_params* p, const float* obs, int* coll_any, int* all_in) {
    const int A = p->num_agents, O = p->num_obstacles, R = A - 1;
    const int S = 2 + 2 * O + 2 * R;
    float in_t[MO_MAX_AGENTS], risk[MO_MAX_AGENTS], dsc[MO_MAX_AGENTS], head[MO_MAX_AGENTS],
          soft[MO_MAX_AGENTS], bond[MO_MAX_AGENTS];
    int any_coll = 0, all_t = 1;
    for (int i = 0; i < A; ++i) {
        const float* o = obs + (size_t)S * i;
        const float ta = o[0], td = o[1];
        const float* od = o + 2 + O;
        const float* ad = o + 2 + 2 * O + R;
        float ob_risk = 0.f, ob_coll = 0.f, ag_risk = 0.f, ag_coll = 0.f;
        for (int j = 0; j < O; ++j) {
            if (od[j] < p->ob_risk_dist) ob_risk = 1.f;
            if (od[j] < p->ob_coll_dist) ob_coll = 1.f;
        }
        float cnt = 0.f, q[MO_MAX_AGENTS];
        for (int k = 0; k < R; ++k) {
            if (ad[k] < p->ag_risk_dist) ag_risk = 1.f;
            if (ad[k] < p->ag_coll_dist) ag_coll = 1.f;
            float above = p->agents_min_d < ad[k] ? 1.f : 0.f;
            float below = ad[k] < p->agents_max_d ? 1.f : 0.f;
            cnt = cnt + above * below;
            float sd = (ad[k] - p->ideal_dist) / p->bond_sharpness;
            q[k] = 1.0f / (1.0f + sd * sd);
        }
        float rk = ob_risk + ag_risk; if (rk > 1.f) rk = 1.f;
        float cl = ob_coll + ag_coll; if (cl > 1.f) cl = 1.f;
        if (cl > 0.f) any_coll = 1;
        in_t[i] = td < p->target_radius ? 1.f : 0.f;
        if (!(in_t[i] > 0.f)) all_t = 0;
        risk[i] = rk;
        float capped = cnt > p->max_at_prop_d ? p->max_at_prop_d : cnt;
        dsc[i] = capped / p->max_at_prop_d;
        head[i] = fabsf(ta) < p->max_angle_diff ? 1.f : 0.f;
        soft[i] = -1.0f * (td / p->init_dist);
        bond[i] = mo_torch_row_sum(q, R) / (float)R;
    }
    float tar = all_t ? 1.f : 0.f;
    float r[MO_MAX_AGENTS];
    for (int i = 0; i < A; ++i) {
        float v = (p->target_factor * tar) + (p->heading_factor * head[i]);
        v = v + (p->distance_factor * dsc[i]);
        v = v + (p->soft_factor * soft[i]);
        v = v + (p->bond_factor * bond[i]);
        v = v - (p->risk_factor * risk[i]);
        r[i] = v;
    }
    *coll_any = any_coll; *all_in = all_t;
    return mo_torch_row_sum(r, A) / (float)A;
}

/* environment.py:86-90, literally: (1-m)*old + m*new with m in {0,1}. */
static inline float mo_blend(float old_v, float new_v, float m) {
    return ((1.0f - m) * old_v) + (m * new_v);
}

/* environment.py:92-107.  All tensors in the reference's layouts:
 *   states (B,A,5) in-out, obstacles (B,O,2) in-out, target (B,2) in-out,
 *   step_num (B) f32 in-out, terminates (B) u8 in-out, actions (B,A,2),
 *   obs (B,A,obs_size) out (POST-reset), rewards (B), terminated/truncated (B) u8,
 *   stats[3] += (num_trunc, num_col, num_tar).
 * If obs_pre is non-NULL the pre-reset observations (the ones rewards are
 * computed from) are stored there too -- a debugging aid for parity triage. */
void mo_step(const mo_params* p, const mo_reset* rs, float* states, float* obstacles,
             float* target, float* step_num, uint8_t* terminates, const float* actions,
             float* obs, float* rewards, uint8_t* terminated, uint8_t* truncated,
             uint64_t* stats, float* obs_pre) {
    const int A = p->num_agents, O = p->num_obstacles;
    const int S = 2 + 2 * O + 2 * (A - 1);
    uint64_t n_trunc = 0, n_col = 0, n_tar = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_trunc, n_col, n_tar)
    for (int64_t b = 0; b < p->num_envs; ++b) {
        float* st = states + (size_t)b * A * 5;
        float* ob = obstacles + (size_t)b * O * 2;
        float* tg = target + (size_t)b * 2;
        float* o = obs + (size_t)b * A * S;
        const float* ac = actions + (size_t)b * A * 2;

        for (int i = 0; i < A; ++i) mo_move_agent(p, st + 5 * i, ac + 2 * i);
        float sn = step_num[b] + 1.0f;
        int trunc = sn > (float)(p->episode_len - 1);
        n_trunc += (uint64_t)trunc;

        mo_observe_env(p, st, ob, tg, o);
        if (obs_pre) memcpy(obs_pre + (size_t)b * A * S, o, sizeof(float) * A * S);
        int coll_any, all_in;
        rewards[b] = mo_reward_env(p, o, &coll_any, &all_in);
        n_col += (uint64_t)coll_any; n_tar += (uint64_t)all_in;
        int term_old = terminates[b] != 0;
        int term = coll_any || term_old;
        terminates[b] = (uint8_t)((!term_old) && all_in);
        terminated[b] = (uint8_t)term; truncated[b] = (uint8_t)trunc;

        int done = term || trunc;
        float m = done ? 1.0f : 0.0f;
        if (rs->alias_first_step) {
            /* new == current tensors (aliased): (1-m)*x + m*x, still literal */
            for (int k = 0; k < A * 5; ++k) st[k] = mo_blend(st[k], st[k], m);
            for (int k = 0; k < O * 2; ++k) ob[k] = mo_blend(ob[k], ob[k], m);
            for (int k = 0; k < 2; ++k) tg[k] = mo_blend(tg[k], tg[k], m);
        } else {
            const float* ts = rs->tmpl_states + (size_t)b * rs->states_env_stride;
            const float* tt = rs->tmpl_target + (size_t)b * rs->target_env_stride;
            float newob[2 * MO_MAX_OBSTACLES];
            const float* to;
            if (rs->tmpl_obstacles) {
                to = rs->tmpl_obstacles + (size_t)b * rs->obstacles_env_stride;
            } else {
                mo_sample_obstacles(p, rs->seed, rs->step_counter, rs->env_id_offset + (uint64_t)b, newob);
                to = newob;
            }
            float noisy[5 * MO_MAX_AGENTS];
            if (rs->noisy) {     /* the reference samples the whole batch every step and blends by mask */
                mo_sample_agents_noisy(rs, rs->step_counter, rs->env_id_offset + (uint64_t)b, A, ts, noisy);
                ts = noisy;
            }
            for (int k = 0; k < A * 5; ++k) st[k] = mo_blend(st[k], ts[k], m);
            for (int k = 0; k < O * 2; ++k) ob[k] = mo_blend(ob[k], to[k], m);
            for (int k = 0; k < 2; ++k) tg[k] = mo_blend(tg[k], tt[k], m);
        }
        step_num[b] = mo_blend(sn, 0.0f, m);

        /* environment.py:105 -- the reference re-observes every env after the blend */
        mo_observe_env(p, st, ob, tg, o);
    }
    stats[0] += n_trunc; stats[1] += n_col; stats[2] += n_tar;
}

/* utils.py:349-368: the 3-agent triangle, computed as torch does it:
 * float32(pos_const * float32(c)) + float32(centre).  Generalised ring for
 * A != 3 is built by the Python side and passed in. */
void mo_triangle_template(float ags_dist, float cx, float cy, float speed, float* tmpl /* (3,5) */) {
    const double k = 1.0 / sqrt(3.0);
    const float unit[3][2] = {{(float)(-k), 1.f}, {(float)(2.0 * k), 0.f}, {(float)(-k), -1.f}};
    const float pc = 0.5f * ags_dist;
    for (int i = 0; i < 3; ++i) {
        tmpl[5 * i + 0] = (pc * unit[i][0]) + cx;
        tmpl[5 * i + 1] = (pc * unit[i][1]) + cy;
        tmpl[5 * i + 2] = 1.f; tmpl[5 * i + 3] = 0.f; tmpl[5 * i + 4] = speed;
    }
}

/* the oracle's own sin(0)/cos(1)/acos(2) on arrays -- used to patch torch.sin/cos/acos
 * when goldens are generated from the real reference (tests/golden/refload.py) */
void mo_trig(int which, const float* x, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float s, c;
        if (which == 2) { y[i] = mt_acosf(x[i]); continue; }
        mt_sincosf(x[i], &s, &c);
        y[i] = which == 0 ? s : c;
    }
}
/* SLEEF u10 (torch's non-MKL CPU backend), for accuracy comparisons only */
void mo_trig_sleef(int which, const float* x, float* y, size_t n) {
    for (size_t i = 0; i < n; ++i)
        y[i] = which == 0 ? tcm_sinf(x[i]) : which == 1 ? tcm_cosf(x[i]) : tcm_acosf(x[i]);
}

int mo_abi_version(void) { return 2; }
