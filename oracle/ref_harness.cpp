// ref_harness.cpp — C-ABI shim over the UNMODIFIED reference engine.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the
// reference's own cppsrc/{game,Pieces,player}.cpp, read in place from
// $(REF)/cppsrc (never copied into this repo), into oracle/_ref/libref_harness.so.
// It exists because the reference's Python module cannot set bar counts
// (backgammon_bindings.cpp:57-59 binds only getters) while Pieces' fields are
// public (Pieces.hpp:8-12): going through C++ lets tests drive the real
// Game::evaluateTurnSequences (game.cpp:193-222) on arbitrary positions.
#include "game.hpp"
#include <cstdint>
#include <cstring>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

namespace {
struct Fixture {
    Player p1{"A", Player::PLAYER1}, p2{"B", Player::PLAYER2};
    Game game{0};
    Fixture() { game.setPlayers(&p1, &p2); }
    void load(const int32_t *s)
    {
        for (int i = 0; i < 24; i++) game.gameboard[i] = s[i];
        game.pieces.numPieces_p1 = s[24];
        game.pieces.numPieces_p2 = s[25];
        game.pieces.freedPieces_p1 = s[26];
        game.pieces.freedPieces_p2 = s[27];
    }
    void store(int32_t *s)
    {
        for (int i = 0; i < 24; i++) s[i] = game.gameboard[i];
        s[24] = game.pieces.numPieces_p1;
        s[25] = game.pieces.numPieces_p2;
        s[26] = game.pieces.freedPieces_p1;
        s[27] = game.pieces.freedPieces_p2;
    }
};
Fixture &fx()
{
    static thread_local Fixture f;
    return f;
}
} // namespace

extern "C" {

// Game::legalMoves (game.cpp:80-105)
int ref_legal_moves(const int32_t *s, int player, int die, int8_t *out_pairs)
{
    Fixture &f = fx();
    f.load(s);
    auto mv = f.game.legalMoves(player, die);
    for (size_t i = 0; i < mv.size(); i++) {
        out_pairs[2 * i] = (int8_t)mv[i].first;
        out_pairs[2 * i + 1] = (int8_t)mv[i].second;
    }
    return (int)mv.size();
}

// Game::tryMove (game.cpp:573-663); state updated in place; err gets the message
int ref_try_move(int32_t *s, int player, int dice, int origin, int dest, char *err, int errcap)
{
    Fixture &f = fx();
    f.load(s);
    std::string e;
    bool ok = f.game.tryMove(player == 0 ? &f.p1 : &f.p2, dice, origin, dest, e);
    f.store(s);
    if (err && errcap > 0) {
        std::strncpy(err, e.c_str(), (size_t)errcap - 1);
        err[errcap - 1] = 0;
    }
    return ok ? 1 : 0;
}

// Game::over (game.cpp:388-407): -1 not over, else winner
int ref_game_over(const int32_t *s)
{
    Fixture &f = fx();
    f.load(s);
    int w = -1;
    return f.game.over(&w) ? w : -1;
}

// Game::evaluateTurnSequences (game.cpp:193-222).  Same output layout as
// orc_turn_sequences: returns N, or -N when cap is too small.
long ref_turn_sequences(const int32_t *s, int player, int d1, int d2, long cap,
                        int8_t *seq_moves, int8_t *seq_len, int32_t *states)
{
    Fixture &f = fx();
    f.load(s);
    TurnEval ev = f.game.evaluateTurnSequences(player, d1, d2);
    long n = (long)ev.sequences.size();
    if (n > cap) return -n;
    for (long i = 0; i < n; i++) {
        const auto &q = ev.sequences[i];
        if (seq_moves) {
            std::memset(seq_moves + 8 * i, 0, 8);
            for (size_t j = 0; j < q.size() && j < 4; j++) {
                seq_moves[8 * i + 2 * j] = (int8_t)q[j].first;
                seq_moves[8 * i + 2 * j + 1] = (int8_t)q[j].second;
            }
        }
        if (seq_len) seq_len[i] = (int8_t)q.size();
        if (states)
            for (int j = 0; j < 28; j++) states[28 * i + j] = ev.states[i][j];
    }
    return n;
}

// CPU baseline: run the reference's evaluateTurnSequences over a batch of
// (state, player, d1, d2) queries; returns seconds, writes total sequences.
double ref_bench_enumerate(const int32_t *states, const int8_t *player, const int8_t *dice,
                           long n, int64_t *total_sequences)
{
    Fixture &f = fx();
    int64_t tot = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (long i = 0; i < n; i++) {
        f.load(states + 28 * i);
        TurnEval ev = f.game.evaluateTurnSequences(player[i], dice[2 * i], dice[2 * i + 1]);
        tot += (int64_t)ev.sequences.size();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (total_sequences) *total_sequences = tot;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Game::evaluateTurnSequences over a batch of 32-byte records (28 state bytes, mover, d1, d2, pad) on `threads` threads:
// per query the number of sequences and the enumeration digest (D <- D * 0x9E3779B97F4A7C15 + h(leaf), h = four
// splitmix64 rounds over the state's five bit-planes and the packed moves) computed HERE from the reference's own output,
// so that the 10^6-position sweep can be compared with the reference itself without moving 1.4 GB of sequences.
static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

void ref_turn_summary_batch(const int8_t *records, long n, int threads, int64_t *n_seq, uint64_t *digest)
{
    std::atomic<long> next{0};
    auto work = [&]() {
        Fixture &f = fx();
        int32_t s[28];
        for (;;) {
            const long lo = next.fetch_add(64);
            if (lo >= n) break;
            const long hi = lo + 64 < n ? lo + 64 : n;
            for (long i = lo; i < hi; i++) {
                const int8_t *r = records + 32 * i;
                for (int j = 0; j < 28; j++) s[j] = r[j];
                f.load(s);
                TurnEval ev = f.game.evaluateTurnSequences(r[28], r[29], r[30]);
                uint64_t dg = 0;
                for (size_t k = 0; k < ev.sequences.size(); k++) {
                    uint32_t w[5] = {0, 0, 0, 0, 0};
                    for (int j = 0; j < 28; j++) {
                        const int v = ev.states[k][j];
                        const uint32_t mag = (uint32_t)(v < 0 ? -v : v);
                        for (int b = 0; b < 4; b++) w[b] |= ((mag >> b) & 1u) << j;
                        if (v < 0) w[4] |= 1u << j;
                    }
                    const auto &q = ev.sequences[k];
                    uint64_t m = (uint64_t)q.size() << 40;
                    for (size_t j = 0; j < q.size(); j++)
                        m |= ((uint64_t)(uint8_t)q[j].first | ((uint64_t)(uint8_t)q[j].second << 5)) << (10 * j);
                    uint64_t h = mix64((uint64_t)w[0] | ((uint64_t)w[1] << 32));
                    h = mix64(h ^ ((uint64_t)w[2] | ((uint64_t)w[3] << 32)));
                    h = mix64(h ^ (uint64_t)w[4]);
                    h = mix64(h ^ m);
                    dg = dg * 0x9E3779B97F4A7C15ULL + h;
                }
                n_seq[i] = (int64_t)ev.sequences.size();
                digest[i] = dg;
            }
        }
    };
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
}

} // extern "C"
