/* bgx_oracle.c — CPU restatement of the reference hot path (see bgx_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY — not product code, never linked into libbgx.
 * Written from the behaviour of the reference, not
 * from its text; every function cites the reference lines it restates.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: the float paths are
 * meant to round like the reference's separate torch ops).
 */
#include "bgx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ rules */

static inline int jailed(const int32_t *s, int player) { return s[24 + (player ? 1 : 0)]; }

/* game.cpp:416-457.  multi = +1 (PLAYER1) / -1 (PLAYER2).  A side with a
 * checker on the bar may only move from its bar origin (0 for P1, 25 for P2). */
int orc_valid_origin(const int32_t *s, int multi, int idx)
{
    if (multi == -1) {
        if (s[25] > 0) return idx == 25;
    } else if (multi == 1) {
        if (s[24] > 0) return idx == 0;
    }
    if (idx < 1 || idx > 24) return 0;
    return s[idx - 1] * multi > 0;
}

/* game.cpp:488-557.  Bear-off legality, quirks Q3/Q4 of SURVEY appendix A.3. */
int orc_can_free(const int32_t *s, int multi, int dice, int origin)
{
    int player = (multi == 1) ? 0 : 1;
    if (jailed(s, player) != 0) return 0;
    for (int pt = 1; pt <= 24; pt++) {
        if (player == 0) { if (pt < 19 && s[pt - 1] > 0) return 0; }   /* game.cpp:504 */
        else             { if (pt > 6 && s[pt - 1] < 0) return 0; }    /* game.cpp:512 */
    }
    if (player == 0) {
        if (dice > 25 - origin)                                        /* game.cpp:526-536 */
            for (int i = origin; i <= 23; i++)
                if (s[i] > 0) return 0;
    } else {
        if (dice > origin)                                             /* game.cpp:542-552 */
            for (int i = origin; i <= 6; i++)
                if (s[i] != 0) return 0;
    }
    return 1;
}

/* game.cpp:459-485 */
int orc_valid_destination(const int32_t *s, int multi, int idx, int dice, int origin)
{
    if (idx == 0 || idx >= 25) return orc_can_free(s, multi, dice, origin);
    if (idx < 1 || idx > 24) return 0;
    return s[idx - 1] * multi >= -1;       /* own, empty or a single enemy blot */
}

/* game.cpp:80-105: origins scanned 0..25 ascending for both sides, destination
 * clamped into [0,25]. */
int orc_legal_moves(const int32_t *s, int player, int die, int8_t *out)
{
    int multi = player == 0 ? 1 : -1, n = 0;
    for (int o = 0; o <= 25; o++) {
        if (!orc_valid_origin(s, multi, o)) continue;
        int d = o + multi * die;
        if (d > 25) d = 25;
        if (d < 0) d = 0;
        if (orc_valid_destination(s, multi, d, die, o)) {
            if (out) { out[2 * n] = (int8_t)o; out[2 * n + 1] = (int8_t)d; }
            n++;
        }
    }
    return n;
}

/* Pieces.cpp:45-55 (quirk Q10: the fall-through branch touches P2's counter) */
static void remove_jailed(int32_t *s, int player)
{
    if (player == 0 && s[24] > 0) s[24] -= 1;
    else s[25] -= 1;
}

/* game.cpp:573-663.  Returns an ORC_* code; the state is mutated only on
 * success.  Destinations 0 / 25 skip every legality check but "origin is a
 * board point" (quirk Q7). */
int orc_try_move(int32_t *s, int player, int dice, int origin, int dest)
{
    int multi = (player == 1) ? -1 : 1;
    if (!orc_valid_origin(s, multi, origin)) return ORC_ERR_INVALID_ORIGIN;
    if (origin < 0 || origin > 25) return ORC_ERR_ORIGIN_RANGE;
    if (dest < 0 || dest > 25) return ORC_ERR_DEST_RANGE;

    int diff = origin - dest;
    if (dest != 0 && dest != 25) {
        if (diff * (-multi) < 0) return ORC_ERR_DIRECTION;
        if (dice != abs(diff)) return ORC_ERR_DICE;
        if (!orc_valid_destination(s, multi, dest, dice, origin)) return ORC_ERR_INVALID_DEST;
        if (origin == 0 || origin == 25) remove_jailed(s, multi > 0 ? 0 : 1);
        else s[origin - 1] -= multi;
    }
    if (dest == 0 || dest == 25) {
        if (origin == 0 || origin == 25) return ORC_ERR_BEAR_FROM_JAIL;
        s[26 + (multi > 0 ? 0 : 1)] += 1;                              /* Pieces.cpp:84-94 */
        s[origin - 1] -= multi;
        return ORC_OK;
    }
    if (s[dest - 1] * multi == -1) {                                   /* hit a blot */
        s[dest - 1] = 0;
        s[24 + (multi > 0 ? 1 : 0)] += 1;
    }
    s[dest - 1] += multi;
    return ORC_OK;
}

/* game.cpp:388-407 */
int orc_game_over(const int32_t *s)
{
    if (s[26] == 15) return 0;
    if (s[27] == 15) return 1;
    return -1;
}

/* ------------------------------------------------------------ enumeration */

typedef struct {
    long cap, n;
    int8_t *moves, *lens;
    int32_t *states;
    int player;
    const int32_t *root;
} seq_sink;

/* evaluateTurnSequences, game.cpp:201-220: the afterstate of a sequence is the
 * root replayed move by move with die = |origin - dest|. */
static void sink_emit(seq_sink *k, const int8_t *prefix, int len)
{
    long i = k->n++;
    if (i >= k->cap) return;
    if (k->moves) {
        memset(k->moves + 8 * i, 0, 8);
        memcpy(k->moves + 8 * i, prefix, (size_t)(2 * len));
    }
    if (k->lens) k->lens[i] = (int8_t)len;
    if (k->states) {
        int32_t *st = k->states + 28 * i;
        memcpy(st, k->root, 28 * sizeof(int32_t));
        for (int j = 0; j < len; j++) {
            int o = prefix[2 * j], d = prefix[2 * j + 1];
            orc_try_move(st, k->player, abs(o - d), o, d);
        }
    }
}

/* collectDoubles, game.cpp:109-131 */
static void doubles_dfs(seq_sink *k, const int32_t *s, int die, int depth, int8_t *prefix)
{
    int8_t mv[52];
    int n = orc_legal_moves(s, k->player, die, mv);
    if (depth == 4 || n == 0) { sink_emit(k, prefix, depth); return; }
    for (int i = 0; i < n; i++) {
        int32_t child[28];
        memcpy(child, s, sizeof child);
        orc_try_move(child, k->player, die, mv[2 * i], mv[2 * i + 1]);
        prefix[2 * depth] = mv[2 * i];
        prefix[2 * depth + 1] = mv[2 * i + 1];
        doubles_dfs(k, child, die, depth + 1, prefix);
    }
}

/* legalTurnSequences, game.cpp:134-191 */
long orc_turn_sequences(const int32_t *s, int player, int d1, int d2, long cap,
                        int8_t *seq_moves, int8_t *seq_len, int32_t *states)
{
    seq_sink k = {cap, 0, seq_moves, seq_len, states, player, s};
    int8_t prefix[8];
    if (d1 != d2) {
        for (int order = 0; order < 2; order++) {
            int x = order ? d2 : d1, y = order ? d1 : d2;
            int8_t m1[52], m2[52];
            int n1 = orc_legal_moves(s, player, x, m1);
            for (int i = 0; i < n1; i++) {
                int32_t child[28];
                memcpy(child, s, sizeof child);
                orc_try_move(child, player, x, m1[2 * i], m1[2 * i + 1]);
                prefix[0] = m1[2 * i];
                prefix[1] = m1[2 * i + 1];
                int n2 = orc_legal_moves(child, player, y, m2);
                if (n2 == 0) sink_emit(&k, prefix, 1);
                for (int j = 0; j < n2; j++) {
                    prefix[2] = m2[2 * j];
                    prefix[3] = m2[2 * j + 1];
                    sink_emit(&k, prefix, 2);
                }
            }
        }
    } else {
        doubles_dfs(&k, s, d1, 0, prefix);
    }
    return k.n <= cap ? k.n : -k.n;
}

/* ---- summary: N, U, digest (definition shared with the CUDA kernel; see
 * DESIGN.md "enumeration digest") */

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

/* the five bit-planes of a state row: bit i of plane b = bit b of |s[i]|,
 * plane 4 = sign bits */
static void state_planes(const int32_t *st, uint32_t w[5])
{
    memset(w, 0, 5 * sizeof(uint32_t));
    for (int i = 0; i < 28; i++) {
        uint32_t mag = (uint32_t)abs(st[i]);
        for (int b = 0; b < 4; b++) w[b] |= ((mag >> b) & 1u) << i;
        if (st[i] < 0) w[4] |= 1u << i;
    }
}

static uint64_t leaf_hash(const int32_t *st, const int8_t *mv, int len)
{
    uint32_t w[5];
    state_planes(st, w);
    uint64_t m = (uint64_t)len << 40;
    for (int j = 0; j < len; j++)
        m |= ((uint64_t)(uint8_t)mv[2 * j] | ((uint64_t)(uint8_t)mv[2 * j + 1] << 5)) << (10 * j);
    uint64_t h = mix64((uint64_t)w[0] | ((uint64_t)w[1] << 32));
    h = mix64(h ^ ((uint64_t)w[2] | ((uint64_t)w[3] << 32)));
    h = mix64(h ^ (uint64_t)w[4]);
    h = mix64(h ^ m);
    return h;
}

static int cmp_state(const void *a, const void *b) { return memcmp(a, b, 28 * sizeof(int32_t)); }

void orc_turn_summary(const int32_t *s, int player, int d1, int d2,
                      int64_t *n_seq, int64_t *n_unique, uint64_t *digest)
{
    long cap = 1024, n;
    int8_t *mv = NULL, *ln = NULL;
    int32_t *st = NULL;
    for (;;) {
        mv = malloc((size_t)cap * 8); ln = malloc((size_t)cap); st = malloc((size_t)cap * 28 * 4);
        n = orc_turn_sequences(s, player, d1, d2, cap, mv, ln, st);
        if (n >= 0) break;
        free(mv); free(ln); free(st);
        cap = -n;
    }
    uint64_t dg = 0;
    for (long i = 0; i < n; i++)
        dg = dg * 0x9E3779B97F4A7C15ULL + leaf_hash(st + 28 * i, mv + 8 * i, ln[i]);
    long u = 0;
    if (n > 0) {
        qsort(st, (size_t)n, 28 * sizeof(int32_t), cmp_state);
        u = 1;
        for (long i = 1; i < n; i++)
            if (memcmp(st + 28 * i, st + 28 * (i - 1), 28 * sizeof(int32_t)) != 0) u++;
    }
    if (n_seq) *n_seq = n;
    if (n_unique) *n_unique = u;
    if (digest) *digest = dg;
    free(mv); free(ln); free(st);
}

/* orc_turn_summary over a batch of 32-byte records (28 state bytes, mover, d1, d2, pad) on
 * `threads` POSIX threads: lets the tests check EVERY position of the 10^6-position sweep
 * (BASELINE.json configs[1]) against this restatement instead of a sample. */
typedef struct {
    const int8_t *rec; long lo, hi;
    int64_t *n_seq, *n_unique; uint64_t *digest;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    for (long i = j->lo; i < j->hi; i++) {
        const int8_t *r = j->rec + 32 * i;
        int32_t s[28];
        for (int k = 0; k < 28; k++) s[k] = r[k];
        orc_turn_summary(s, r[28], r[29], r[30], j->n_seq + i, j->n_unique + i, j->digest + i);
    }
    return NULL;
}

void orc_turn_summary_batch(const int8_t *records, long n, int threads,
                            int64_t *n_seq, int64_t *n_unique, uint64_t *digest)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    batch_job job[256];
    /* interleaved blocks would balance better, but the sweep is shuffled: contiguous is fine */
    for (int t = 0; t < threads; t++) {
        job[t].rec = records; job[t].lo = n * t / threads; job[t].hi = n * (t + 1) / threads;
        job[t].n_seq = n_seq; job[t].n_unique = n_unique; job[t].digest = digest;
        pthread_create(&tid[t], NULL, batch_worker, &job[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
}

/* ------------------------------------------------------------- the model */

/* _encode_states_np, model.py:111-144.  Slots 0-3 of a point are PLAYER1's
 * checkers, 4-7 PLAYER2's: [n>=1, n>=2, n>=3, (n-3)/2].  192/193 = the MOVER's
 * flag (not flipped for afterstates, quirk Q12); 194/195 = jailed/2;
 * 196/197 = borne-off/15.0 (a float64 divide rounded to float32). */
void orc_encode(const int32_t *states, long n, int turn, float *X)
{
    for (long r = 0; r < n; r++) {
        const int32_t *s = states + 28 * r;
        float *x = X + ORC_FEATS * r;
        memset(x, 0, ORC_FEATS * sizeof(float));
        for (int i = 0; i < 24; i++) {
            int v = s[i], c = v < 0 ? -v : v;
            int base = 8 * i + (v > 0 ? 0 : 4);
            if (c >= 1) x[base] = 1.0f;
            if (c >= 2) x[base + 1] = 1.0f;
            if (c >= 3) x[base + 2] = 1.0f;
            if (c >= 4) x[base + 3] = (float)((double)(c - 3) / 2.0);
        }
        x[192] = turn == 0 ? 1.0f : 0.0f;
        x[193] = turn == 0 ? 0.0f : 1.0f;
        x[194] = (float)((double)s[24] / 2.0);
        x[195] = (float)((double)s[25] / 2.0);
        x[196] = (float)((double)s[26] / 15.0);
        x[197] = (float)((double)s[27] / 15.0);
    }
}

static inline float sigmoidf_(float z) { return 1.0f / (1.0f + expf(-z)); }

/* TDLGammonModel.forward, model.py:63-67: V = sigmoid(w2 . sigmoid(W1 x + b1) + b2),
 * fp32, features accumulated in ascending index order. */
void orc_forward(const float *W1, const float *b1, const float *w2, const float *b2,
                 const float *X, long n, float *V, float *H)
{
    float h[ORC_HIDDEN];
    for (long r = 0; r < n; r++) {
        const float *x = X + ORC_FEATS * r;
        for (int j = 0; j < ORC_HIDDEN; j++) {
            const float *w = W1 + ORC_FEATS * j;
            float z = 0.0f;
            for (int f = 0; f < ORC_FEATS; f++)
                if (x[f] != 0.0f) z += w[f] * x[f];
            h[j] = sigmoidf_(z + b1[j]);
        }
        float y = 0.0f;
        for (int j = 0; j < ORC_HIDDEN; j++) y += w2[j] * h[j];
        V[r] = sigmoidf_(y + b2[0]);
        if (H) memcpy(H + ORC_HIDDEN * r, h, sizeof h);
    }
}

/* apply_td_updates, train.py:124-172 (closed-form gradients of the 2-layer
 * sigmoid net replace autograd).  Per step:
 *   v' = V(s_{t+1}) with the CURRENT weights, v = V(s_t), delta = v' - v
 *   e <- lambda*e + grad V(s_t) ;  p <- p + (lr*delta)*e        (train.py:136-147)
 * terminal step: delta = (1|0) - V(s_{T-1})                      (train.py:165-170)
 * lambda and lr*delta reach the tensors as fp32 scalars (torch scalar-tensor
 * arithmetic), the product lr*delta itself is a Python float64. */
void orc_td_replay(float *W1, float *b1, float *w2, float *b2,
                   const float *X, long T, int player1_won,
                   double lr, double lambda, double *sq_errors)
{
    float *eW1 = calloc(ORC_FEATS * ORC_HIDDEN, sizeof(float));
    float eb1[ORC_HIDDEN] = {0}, ew2[ORC_HIDDEN] = {0}, eb2 = 0.0f;
    float h[ORC_HIDDEN], gh[ORC_HIDDEN];
    const float lam = (float)lambda;

    for (long t = 0; t < T; t++) {
        const float *x = X + ORC_FEATS * t;
        double delta;
        float v_cur;
        if (t < T - 1) {
            float v_next;
            orc_forward(W1, b1, w2, b2, X + ORC_FEATS * (t + 1), 1, &v_next, NULL);
            orc_forward(W1, b1, w2, b2, x, 1, &v_cur, h);
            float d32 = v_next - v_cur;
            delta = (double)d32;
            if (sq_errors) sq_errors[t] = delta * delta;
        } else {
            orc_forward(W1, b1, w2, b2, x, 1, &v_cur, h);
            delta = (player1_won ? 1.0 : 0.0) - (double)v_cur;
        }
        const float c = (float)(lr * delta);
        /* gradients of v_cur w.r.t. the PRE-update weights */
        const float gv = (1.0f - v_cur) * v_cur;
        for (int j = 0; j < ORC_HIDDEN; j++) gh[j] = ((gv * w2[j]) * (1.0f - h[j])) * h[j];   /* torch sigmoid_backward order */
        /* named_parameters order: fc1.weight, fc1.bias, fc2.weight, fc2.bias */
        for (int j = 0; j < ORC_HIDDEN; j++) {
            float *e = eW1 + ORC_FEATS * j, *w = W1 + ORC_FEATS * j;
            for (int f = 0; f < ORC_FEATS; f++) {
                float g = gh[j] * x[f];
                e[f] = lam * e[f] + g;
                w[f] = w[f] + c * e[f];
            }
        }
        for (int j = 0; j < ORC_HIDDEN; j++) {
            eb1[j] = lam * eb1[j] + gh[j];
            b1[j] = b1[j] + c * eb1[j];
        }
        for (int j = 0; j < ORC_HIDDEN; j++) {
            float g = gv * h[j];
            ew2[j] = lam * ew2[j] + g;
            w2[j] = w2[j] + c * ew2[j];
        }
        eb2 = lam * eb2 + gv;
        b2[0] = b2[0] + c * eb2;
    }
    free(eW1);
}

/* ------------------------------------------------------------------- dice */

static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

/* Philox4x32-10 (Salmon et al., SC'11), the published algorithm */
void orc_philox4x32(uint32_t seed_lo, uint32_t seed_hi,
                    uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t c[4] = {c0, c1, c2, c3}, k[2] = {seed_lo, seed_hi};
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

int orc_die(uint32_t x) { return 1 + (int)(((uint64_t)x * 6u) >> 32); }

/* ------------------------------------------------------------ greedy ply */

/* make_move with epsilon = 0, model.py:180-222 */
long orc_greedy_ply(const float *W1, const float *b1, const float *w2, const float *b2,
                    const int32_t *s, int player, int d1, int d2,
                    int32_t *out, float *v_best, int64_t *n_seq)
{
    long cap = 512, n;
    int32_t *st = NULL;
    for (;;) {
        st = malloc((size_t)cap * 28 * 4);
        n = orc_turn_sequences(s, player, d1, d2, cap, NULL, NULL, st);
        if (n >= 0) break;
        free(st);
        cap = -n;
    }
    if (n_seq) *n_seq = n;
    if (n == 0) { free(st); return -1; }
    float *X = malloc((size_t)n * ORC_FEATS * sizeof(float));
    float *V = malloc((size_t)n * sizeof(float));
    orc_encode(st, n, player, X);
    orc_forward(W1, b1, w2, b2, X, n, V, NULL);
    long best = 0;
    for (long i = 1; i < n; i++)
        if (player == 0 ? V[i] > V[best] : V[i] < V[best]) best = i;
    if (out) memcpy(out, st + 28 * best, 28 * sizeof(int32_t));
    if (v_best) *v_best = V[best];
    free(X); free(V); free(st);
    return best;
}

/* orc_greedy_ply over a batch of 32-byte records on `threads` POSIX threads, returning also the first-index sequence
 * itself (model.py:212-220): lets the GPU tests compare choice, value, N AND the reported moves on 10^5 positions. */
typedef struct {
    const float *W1, *b1, *w2, *b2;
    const int8_t *rec; long lo, hi;
    int8_t *after, *moves, *len; float *value; int64_t *n_seq;
} greedy_job;

static void *greedy_worker(void *arg)
{
    greedy_job *j = (greedy_job *)arg;
    long cap = 16384;
    int32_t *st = malloc((size_t)cap * 28 * 4);
    int8_t *mv = malloc((size_t)cap * 8), *ln = malloc((size_t)cap);
    float *X = malloc((size_t)cap * ORC_FEATS * sizeof(float)), *V = malloc((size_t)cap * sizeof(float));
    for (long i = j->lo; i < j->hi; i++) {
        const int8_t *r = j->rec + 32 * i;
        int32_t s[28];
        for (int k = 0; k < 28; k++) s[k] = r[k];
        const int player = r[28];
        long n = orc_turn_sequences(s, player, r[29], r[30], cap, mv, ln, st);
        if (n < 0) {                                   /* more than cap sequences: grow and redo */
            cap = -n;
            free(st); free(mv); free(ln); free(X); free(V);
            st = malloc((size_t)cap * 28 * 4); mv = malloc((size_t)cap * 8); ln = malloc((size_t)cap);
            X = malloc((size_t)cap * ORC_FEATS * sizeof(float)); V = malloc((size_t)cap * sizeof(float));
            n = orc_turn_sequences(s, player, r[29], r[30], cap, mv, ln, st);
        }
        j->n_seq[i] = n;
        j->len[i] = 0; j->value[i] = 0.f;
        memset(j->moves + 8 * i, 0, 8);
        for (int k = 0; k < 28; k++) j->after[28 * i + k] = (int8_t)s[k];
        if (n <= 0) continue;
        orc_encode(st, n, player, X);
        orc_forward(j->W1, j->b1, j->w2, j->b2, X, n, V, NULL);
        long best = 0;
        for (long k = 1; k < n; k++)
            if (player == 0 ? V[k] > V[best] : V[k] < V[best]) best = k;
        for (int k = 0; k < 28; k++) j->after[28 * i + k] = (int8_t)st[28 * best + k];
        memcpy(j->moves + 8 * i, mv + 8 * best, 8);
        j->len[i] = ln[best];
        j->value[i] = V[best];
    }
    free(st); free(mv); free(ln); free(X); free(V);
    return NULL;
}

void orc_greedy_batch(const float *W1, const float *b1, const float *w2, const float *b2,
                      const int8_t *records, long n, int threads,
                      int8_t *after, float *value, int64_t *n_seq, int8_t *moves, int8_t *len)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    greedy_job job[256];
    for (int t = 0; t < threads; t++) {
        greedy_job g = {W1, b1, w2, b2, records, n * t / threads, n * (t + 1) / threads, after, moves, len, value, n_seq};
        job[t] = g;
        pthread_create(&tid[t], NULL, greedy_worker, &job[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
}
