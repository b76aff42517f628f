// bgx_host.cpp — section 1 of include/bgx.h: single-position host functions behind the
// compat Game object.  Pure C++, no CUDA, no dependency on oracle/.
//
// Legal-move generation goes through the same mask algebra the kernels use
// (bgx_core.h: legal_origins), so the CPU test-suite exercises the device rules too.
// tryMove is restated as a validator with the reference's decision order, because the
// order decides which error string a bad move gets (cppsrc/game.cpp:583-643).
#include "../../include/bgx.h"
#include "bgx_core.h"
#include "bgx_internal.h"

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace bgx {

thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// ---- scalar predicates (any die value, any counts): used by tryMove and by the generic
// generator path for arguments outside the kernels' domain (die not in 1..6, |count| > 15)

static bool origin_ok(const int32_t *s, int multi, int idx)               // game.cpp:416-457
{
    if (multi == -1 && s[25] > 0) return idx == 25;
    if (multi == +1 && s[24] > 0) return idx == 0;
    return idx >= 1 && idx <= 24 && s[idx - 1] * multi > 0;
}

static bool bear_off_ok(const int32_t *s, int multi, int dice, int origin) // game.cpp:488-557
{
    const bool p1 = multi == +1;
    if (s[p1 ? 24 : 25] != 0) return false;
    for (int pt = 1; pt <= 24; pt++) {
        if (p1 ? (pt <= 18 && s[pt - 1] > 0) : (pt >= 7 && s[pt - 1] < 0)) return false;
    }
    if (p1 && dice > 25 - origin) {
        for (int pt = origin + 1; pt <= 24; pt++)
            if (s[pt - 1] > 0) return false;
    }
    if (!p1 && dice > origin) {
        for (int pt = origin + 1; pt <= 7; pt++)
            if (s[pt - 1] != 0) return false;
    }
    return true;
}

static bool dest_ok(const int32_t *s, int multi, int idx, int dice, int origin) // game.cpp:459-485
{
    if (idx == 0 || idx >= 25) return bear_off_ok(s, multi, dice, origin);
    if (idx < 1) return false;
    return s[idx - 1] * multi >= -1;
}

static bool in_kernel_domain(const int32_t *s, int die)
{
    if (die < 1 || die > 6) return false;
    for (int i = 0; i < 24; i++)
        if (s[i] > 15 || s[i] < -15) return false;
    return true;
}

static int legal_moves_impl(const int32_t *s, int player, int die, int8_t *out)
{
    int n = 0;
    if (in_kernel_domain(s, die) && (player == 0 || player == 1)) {
        Masks m = masks_of_row(s);
        uint32_t legal = legal_origins(player, die, m, s[24 + player]);
        while (legal) {
            int o = lowest_bit(legal);
            legal &= legal - 1;
            if (out) { out[2 * n] = (int8_t)o; out[2 * n + 1] = (int8_t)destination(player, o, die); }
            n++;
        }
        return n;
    }
    const int multi = player == 0 ? +1 : -1;                               // game.cpp:83
    for (int o = 0; o <= 25; o++) {
        if (!origin_ok(s, multi, o)) continue;
        long d = (long)o + (long)multi * die;
        d = d > 25 ? 25 : (d < 0 ? 0 : d);
        if (dest_ok(s, multi, (int)d, die, o)) {
            if (out) { out[2 * n] = (int8_t)o; out[2 * n + 1] = (int8_t)d; }
            n++;
        }
    }
    return n;
}

static int try_move_impl(int32_t *s, int player, int dice, int origin, int dest)
{
    const int multi = player == 1 ? -1 : +1;                               // game.cpp:579-580
    if (!origin_ok(s, multi, origin)) return BGX_MOVE_INVALID_ORIGIN;
    if (origin < 0 || origin > 25) return BGX_MOVE_ORIGIN_RANGE;
    if (dest < 0 || dest > 25) return BGX_MOVE_DEST_RANGE;
    const bool to_edge = dest == 0 || dest == 25;
    const bool from_bar = origin == 0 || origin == 25;
    if (to_edge) {
        // quirk Q7: no direction / die / legality check at all on this path
        if (from_bar) return BGX_MOVE_BEAR_FROM_JAIL;
        s[multi > 0 ? 26 : 27] += 1;
        s[origin - 1] -= multi;
        return BGX_MOVE_OK;
    }
    const int diff = origin - dest;
    if (diff * -multi < 0) return BGX_MOVE_DIRECTION;
    if (dice != std::abs(diff)) return BGX_MOVE_DICE;
    if (!dest_ok(s, multi, dest, dice, origin)) return BGX_MOVE_INVALID_DEST;
    if (from_bar) {
        // Pieces::removeJailedPiece (Pieces.cpp:45-55) incl. its fall-through on P2's counter
        if (multi > 0 && s[24] > 0) s[24] -= 1;
        else s[25] -= 1;
    } else {
        s[origin - 1] -= multi;
    }
    if (s[dest - 1] * multi == -1) {
        s[dest - 1] = 0;
        s[multi > 0 ? 25 : 24] += 1;
    }
    s[dest - 1] += multi;
    return BGX_MOVE_OK;
}

struct SeqOut {
    int64_t cap, n;
    int8_t *moves, *lens;
    int32_t *states;
    void emit(const int32_t *st, const int8_t *prefix, int len)
    {
        int64_t i = n++;
        if (i >= cap) return;
        if (moves) {
            std::memset(moves + 8 * i, 0, 8);
            std::memcpy(moves + 8 * i, prefix, (size_t)(2 * len));
        }
        if (lens) lens[i] = (int8_t)len;
        if (states) std::memcpy(states + 28 * i, st, 28 * sizeof(int32_t));
    }
};

// One pass of the turn tree: dice[k] is the die of the k-th move (2 entries for a
// non-double order, 4 for a double).  A node with no move is a leaf (game.cpp:117-121,
// 148-151); the root of a non-double pass is not (quirk Q5).
static void walk(SeqOut &out, const int32_t *s, int player, const int *dice, int maxlen, int depth,
                 int8_t *prefix, bool root_is_leaf)
{
    int8_t mv[52];
    int n = depth < maxlen ? legal_moves_impl(s, player, dice[depth], mv) : 0;
    if (n == 0) {
        if (depth > 0 || root_is_leaf) out.emit(s, prefix, depth);
        return;
    }
    for (int i = 0; i < n; i++) {
        int32_t child[28];
        std::memcpy(child, s, sizeof child);
        // the reference replays through tryMove (game.cpp:126,146,169,209); for generated
        // moves that is exactly the unchecked state update
        try_move_impl(child, player, dice[depth], mv[2 * i], mv[2 * i + 1]);
        prefix[2 * depth] = mv[2 * i];
        prefix[2 * depth + 1] = mv[2 * i + 1];
        walk(out, child, player, dice, maxlen, depth + 1, prefix, root_is_leaf);
    }
}

} // namespace bgx

using namespace bgx;

extern "C" {

const char *bgx_last_error(void) { return g_err; }
int bgx_abi_version(void) { return 2; }

int bgx_legal_moves(const int32_t *position, int player, int die, int8_t *out_pairs, int cap, int *n)
{
    if (!position || !n) { set_error("bgx_legal_moves: null argument"); return BGX_E_INVALID; }
    int8_t tmp[52];
    int k = legal_moves_impl(position, player, die, tmp);
    *n = k;
    if (k > cap) { set_error("bgx_legal_moves: %d moves, cap %d", k, cap); return BGX_E_CAPACITY; }
    if (out_pairs) std::memcpy(out_pairs, tmp, (size_t)(2 * k));
    return BGX_OK;
}

int bgx_try_move(int32_t *position, int player, int dice, int origin, int dest, int *move_code)
{
    if (!position || !move_code) { set_error("bgx_try_move: null argument"); return BGX_E_INVALID; }
    *move_code = try_move_impl(position, player, dice, origin, dest);
    return BGX_OK;
}

const char *bgx_move_error_string(int code)
{
    switch (code) {
    case BGX_MOVE_OK: return "";
    case BGX_MOVE_INVALID_ORIGIN: return "Invalid origin";
    case BGX_MOVE_ORIGIN_RANGE: return "Origin out of range";
    case BGX_MOVE_DEST_RANGE: return "Destination out of range";
    case BGX_MOVE_DIRECTION: return "Cannot move in that direction.";
    case BGX_MOVE_DICE: return "Move does not match dice.";
    case BGX_MOVE_INVALID_DEST: return "Invalid destination.";
    case BGX_MOVE_BEAR_FROM_JAIL: return "Cannot bear off from jail";
    }
    return "?";
}

int bgx_game_over(const int32_t *position, int *winner)
{
    if (!position || !winner) { set_error("bgx_game_over: null argument"); return BGX_E_INVALID; }
    *winner = position[26] == 15 ? 0 : (position[27] == 15 ? 1 : -1);       // game.cpp:393-402
    return BGX_OK;
}

int bgx_turn_sequences(const int32_t *position, int player, int d1, int d2, int64_t cap,
                       int8_t *seq_moves, int8_t *seq_len, int32_t *states, int64_t *n)
{
    if (!position || !n) { set_error("bgx_turn_sequences: null argument"); return BGX_E_INVALID; }
    SeqOut out = {cap, 0, seq_moves, seq_len, states};
    int8_t prefix[8] = {0};
    if (d1 != d2) {
        int a[2] = {d1, d2}, b[2] = {d2, d1};
        walk(out, position, player, a, 2, 0, prefix, false);
        walk(out, position, player, b, 2, 0, prefix, false);
    } else {
        int d[4] = {d1, d1, d1, d1};
        walk(out, position, player, d, 4, 0, prefix, true);
    }
    *n = out.n;
    if (out.n > cap) { set_error("bgx_turn_sequences: %lld sequences, cap %lld", (long long)out.n, (long long)cap); return BGX_E_CAPACITY; }
    return BGX_OK;
}

// host threads of bgx_advance_host: bgx_set_host_threads, else 1..4 by the cores at hand (one thread does 65,536 games in ~0.3 ms).
// Several ranks on one box should say how many they are (bgx_set_host_threads(cores / ranks)): the library reads no environment.
static int g_host_threads = 0;
static int advance_threads()
{
    if (g_host_threads > 0) return g_host_threads;
    int t = (int)std::thread::hardware_concurrency();
    if (t > 4) t = 4;
    return t < 1 ? 1 : t;
}

int bgx_set_host_threads(int n)
{
    if (n < 0 || n > 64) { bgx::set_error("bgx_set_host_threads: 0 (automatic) .. 64"); return BGX_E_INVALID; }
    g_host_threads = n;
    return BGX_OK;
}

} // extern "C"

namespace bgx {

constexpr int kBlock = 16;      // work is split on multiples of this many games

// one game: ~20 ns, almost all of it the ten dependent Philox rounds (hand-written AVX2 over 8 games gained 30 % on the
// build box, gcc's own vectorisation at -O3 LOST 40 %: the scalar form stays)
static void advance_range(const int8_t *chosen, int8_t *next, int64_t lo, int64_t hi, uint32_t k0, uint32_t k1,
                          const int32_t *ply, const int64_t *game_id, int8_t *winner)
{
    for (int64_t i = lo; i < hi; i++) {
        const int8_t *c = chosen + 32 * i;
        int8_t *o = next + 32 * i;
        const int win = c[26] == 15 ? 0 : (c[27] == 15 ? 1 : -1);           // game.cpp:388-407
        const uint64_t g = game_id ? (uint64_t)game_id[i] : (uint64_t)i;
        const Philox r = philox4x32_10(k0, k1, ply ? (uint32_t)ply[i] : 0u, (uint32_t)g, (uint32_t)(g >> 32), 0u);
        const int mover = c[28] ? 1 : 0;
        if (o != c) std::memcpy(o, c, 28);
        o[28] = (int8_t)(win < 0 ? mover ^ 1 : mover);
        o[29] = (int8_t)die_of(r.x[0]);
        o[30] = (int8_t)die_of(r.x[1]);
        o[31] = (int8_t)(win + 1);
        if (winner) winner[i] = (int8_t)win;
    }
}

// A few persistent helper threads, asleep on a condition variable between calls (no spinning: several ranks share the
// box's cores, and an OpenMP team of 2-4 with the runtime's idle threads spinning measured 5x SLOWER than one thread).
class HostPool {
public:
    explicit HostPool(int helpers)
    {
        for (int t = 0; t < helpers; t++) threads_.emplace_back([this, t] { work(t); });
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int helpers() const { return (int)threads_.size(); }
    // fn(part, parts) runs for part = 1..helpers on the helpers and for part 0 on the caller
    void run(const std::function<void(int, int)> &fn)
    {
        const int parts = helpers() + 1;
        { std::lock_guard<std::mutex> g(m_); fn_ = &fn; pending_ = helpers(); gen_++; }
        cv_.notify_all();
        fn(0, parts);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(t + 1, helpers() + 1);
            { std::lock_guard<std::mutex> g(m_); pending_--; }
            done_.notify_one();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

} // namespace bgx

extern "C" {

int bgx_advance_host(const int8_t *chosen, int8_t *next, int64_t n, uint64_t seed, const int32_t *ply,
                     const int64_t *game_id, int8_t *winner)
{
    if (!chosen || !next || n < 0) { set_error("bgx_advance_host: bad argument"); return BGX_E_INVALID; }
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int threads = advance_threads();
    if (threads <= 1 || n < 16384) {
        bgx::advance_range(chosen, next, 0, n, k0, k1, ply, game_id, winner);
        return BGX_OK;
    }
    static std::mutex pool_mutex;                       // one batch at a time through the pool
    static bgx::HostPool *pool = nullptr;               // lives until process exit (helpers are blocked, not spinning)
    std::lock_guard<std::mutex> g(pool_mutex);
    if (!pool) pool = new bgx::HostPool(threads - 1);
    pool->run([&](int part, int parts) {
        const int64_t per = ((n + parts - 1) / parts + bgx::kBlock - 1) / bgx::kBlock * bgx::kBlock;
        const int64_t lo = per * part, hi = lo + per < n ? lo + per : n;
        if (lo < hi) bgx::advance_range(chosen, next, lo, hi, k0, k1, ply, game_id, winner);
    });
    return BGX_OK;
}

} // extern "C"
