/* bgx_oracle.h — CPU restatement of the reference's self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under backgammon-engine_b200/ may include,
 * link or dlopen this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity pinning: every function here is checked (tests/test_oracle_*.py)
 * against (i) the known answers of the reference's own cppsrc/tests.cpp,
 * (ii) golden vectors generated from the reference itself
 * (tests/golden/make_golden.py, run in the build container against
 * oracle/_ref), and (iii) the live oracle/_ref build whenever it is present.
 *
 * All file:line citations are relative to the reference checkout.
 * A position is the reference's TurnEval row (cppsrc/game.hpp:17-28):
 *   s[0..23] board (+ = PLAYER1 count, - = PLAYER2 count, index = point-1),
 *   s[24] jailed P1, s[25] jailed P2, s[26] borne-off P1, s[27] borne-off P2.
 */
#ifndef BGX_ORACLE_H
#define BGX_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_STATE 28
#define ORC_FEATS 198
#define ORC_HIDDEN 128
#define ORC_NPARAMS (ORC_FEATS * ORC_HIDDEN + ORC_HIDDEN + ORC_HIDDEN + 1) /* 25601 */

/* tryMove error codes, in the order game.cpp:583-643 can raise them */
enum {
    ORC_OK = 0,
    ORC_ERR_INVALID_ORIGIN = 1,   /* "Invalid origin"                 game.cpp:585 */
    ORC_ERR_ORIGIN_RANGE = 2,     /* "Origin out of range"            game.cpp:590 */
    ORC_ERR_DEST_RANGE = 3,       /* "Destination out of range"       game.cpp:595 */
    ORC_ERR_DIRECTION = 4,        /* "Cannot move in that direction." game.cpp:608 */
    ORC_ERR_DICE = 5,             /* "Move does not match dice."      game.cpp:614 */
    ORC_ERR_INVALID_DEST = 6,     /* "Invalid destination."           game.cpp:619 */
    ORC_ERR_BEAR_FROM_JAIL = 7    /* "Cannot bear off from jail"      game.cpp:642 */
};

/* rules ------------------------------------------------------------------ */
int orc_valid_origin(const int32_t *s, int multi, int idx);                 /* game.cpp:416-457 */
int orc_valid_destination(const int32_t *s, int multi, int idx, int dice, int origin); /* 459-485 */
int orc_can_free(const int32_t *s, int multi, int dice, int origin);        /* game.cpp:488-557 */
int orc_legal_moves(const int32_t *s, int player, int die, int8_t *out_pairs /*[26][2]*/); /* 80-105 */
int orc_try_move(int32_t *s, int player, int dice, int origin, int dest);   /* game.cpp:573-663 */
int orc_game_over(const int32_t *s);                                        /* game.cpp:388-407; -1 / 0 / 1 */

/* enumeration: legalTurnSequences + evaluateTurnSequences (game.cpp:134-222).
 * seq_moves[k][4][2] (origin,dest), seq_len[k], states[k][28].  Returns N, or
 * -(N needed) if cap is too small (nothing beyond cap is written).
 * Any output pointer may be NULL. */
long orc_turn_sequences(const int32_t *s, int player, int d1, int d2, long cap,
                        int8_t *seq_moves, int8_t *seq_len, int32_t *states);

/* summary of one enumeration: N, number of distinct afterstates U and an
 * order-dependent 64-bit digest over (len, moves, state) of every sequence. */
void orc_turn_summary(const int32_t *s, int player, int d1, int d2,
                      int64_t *n_seq, int64_t *n_unique, uint64_t *digest);

/* the same over n 32-byte records (28 state bytes, mover, d1, d2, pad), on `threads` threads */
void orc_turn_summary_batch(const int8_t *records, long n, int threads,
                            int64_t *n_seq, int64_t *n_unique, uint64_t *digest);

/* encoding + model (pysrc/TD(λ) model/model.py) ---------------------------- */
void orc_encode(const int32_t *states, long n, int turn, float *X /*[n][198]*/);   /* model.py:111-144 */
/* weights in state_dict layout: W1[128][198], b1[128], w2[128], b2[1] */
void orc_forward(const float *W1, const float *b1, const float *w2, const float *b2,
                 const float *X, long n, float *V, float *H /*[n][128] or NULL*/); /* model.py:63-67 */

/* apply_td_updates (train.py:124-172): online TD(lambda) replay of one game,
 * weights updated in place, traces start at zero.  X is [T][198].
 * sq_errors[T-1] receives td_error^2 of the non-terminal steps (may be NULL). */
void orc_td_replay(float *W1, float *b1, float *w2, float *b2,
                   const float *X, long T, int player1_won,
                   double lr, double lambda, double *sq_errors);

/* dice: Philox4x32-10, key = (seed_lo, seed_hi), counter = (c0,c1,c2,c3).
 * This is the engine's dice spec (the reference has no seed API; its dice are
 * injected through Game.setDice, backgammon_bindings.cpp:86). */
void orc_philox4x32(uint32_t seed_lo, uint32_t seed_hi,
                    uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
/* die from a 32-bit draw: 1 + ((x*6) >> 32) */
int orc_die(uint32_t x);

/* greedy 1-ply ply (model.py:180-222 with epsilon = 0): enumerate, encode with
 * the mover's flag, forward, argmax (P1) / argmin (P2) with first-index ties.
 * Returns the chosen sequence index or -1 when there is no sequence; on a
 * choice the afterstate is written to `out`, its value to *v_best. */
long orc_greedy_ply(const float *W1, const float *b1, const float *w2, const float *b2,
                    const int32_t *s, int player, int d1, int d2,
                    int32_t *out, float *v_best, int64_t *n_seq);

/* the same over n 32-byte records on `threads` threads; moves[n][4][2] / len[n] = the chosen (first-index) sequence */
void orc_greedy_batch(const float *W1, const float *b1, const float *w2, const float *b2,
                      const int8_t *records, long n, int threads,
                      int8_t *after /*[n][28]*/, float *value, int64_t *n_seq, int8_t *moves, int8_t *len);

#ifdef __cplusplus
}
#endif
#endif
