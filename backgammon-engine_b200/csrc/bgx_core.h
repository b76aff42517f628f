// bgx_core.h — rules of the reference engine as branch-light bit-mask algebra, shared
// by the host shim (bgx_host.cpp) and the CUDA kernels (bgx_kernels.cu).
//
// The reference scans origins 0..25 and calls three predicates per origin
// (cppsrc/game.cpp:80-105, 416-557).  Here the same answer is computed for all 26
// origins at once from four occupancy masks, so a warp that holds one board point per
// lane gets the whole legal-move set from four ballots and a handful of ALU ops.
//
// "Code space": bit c (0..25) stands for the reference's origin/destination code c:
//   c = 0      PLAYER1's bar (origin) / PLAYER2's bear-off target
//   c = 1..24  board points
//   c = 25     PLAYER2's bar (origin) / PLAYER1's bear-off target
// Board masks are built with bit c = point c, i.e. lane/row index i (= point-1) << 1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BGX_HD __host__ __device__ __forceinline__
#else
#define BGX_HD inline
#endif

namespace bgx {

constexpr uint32_t kPoints = 0x01FFFFFEu;      // bits 1..24
constexpr uint32_t kP1Outside = 0x0007FFFEu;   // points 1..18: PLAYER1 not yet home (game.cpp:504)
constexpr uint32_t kP2Outside = 0x01FFFF80u;   // points 7..24: PLAYER2 not yet home (game.cpp:512)
constexpr uint32_t kP2Window = 0x000000FEu;    // points 1..7: the PLAYER2 over-bear scan (game.cpp:546)

struct Masks {
    uint32_t occ1;   // points holding PLAYER1 checkers   (board > 0)
    uint32_t occ2;   // points holding PLAYER2 checkers   (board < 0)
    uint32_t wall1;  // points PLAYER1 cannot land on     (board <= -2)
    uint32_t wall2;  // points PLAYER2 cannot land on     (board >= +2)
};

BGX_HD int highest_bit(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz(x);
#else
    return 31 - __builtin_clz(x);
#endif
}

BGX_HD int lowest_bit(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs(x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// All legal origins of `player` for one die, as a code-space mask; bit order = the
// reference's emission order (ascending origin, game.cpp:84).  jail_mover = checkers of
// the mover on the bar.
//   entering      game.cpp:418-441 (only the bar origin is valid) + 472-480 (landing)
//   normal moves  game.cpp:449, 472-480
//   bearing off   game.cpp:488-557 incl. the asymmetric over-bear rules
BGX_HD uint32_t legal_origins(int player, int die, const Masks &m, int jail_mover)
{
    if (player == 0) {
        if (jail_mover > 0) return ((m.wall1 >> die) & 1u) ? 0u : 1u;
        uint32_t legal = m.occ1 & ((~m.wall1 & kPoints) >> die);
        if (m.occ1 != 0 && (m.occ1 & kP1Outside) == 0) {
            legal |= m.occ1 & (1u << (25 - die));                 // exact bear-off
            int hi = highest_bit(m.occ1);                           // over-bear: only from the
            if (hi + die > 25) legal |= 1u << hi;                   // highest occupied point
        }
        return legal;
    }
    if (jail_mover > 0) return ((m.wall2 >> (25 - die)) & 1u) ? 0u : (1u << 25);
    uint32_t legal = m.occ2 & ((~m.wall2 & kPoints) << die) & kPoints;
    if (m.occ2 != 0 && (m.occ2 & kP2Outside) == 0) {
        legal |= m.occ2 & (1u << die);                              // exact bear-off
        // over-bear from c < die: nothing of EITHER colour on points c+1..7, i.e. c is the
        // highest occupied bit of the 1..7 window
        int hi = highest_bit((m.occ1 | m.occ2) & kP2Window);
        if (hi < die && ((m.occ2 >> hi) & 1u)) legal |= 1u << hi;
    }
    return legal;
}

// destination code of a legal origin (clamped like game.cpp:89-97)
BGX_HD int destination(int player, int origin, int die)
{
    int d = player == 0 ? origin + die : origin - die;
    return d > 25 ? 25 : (d < 0 ? 0 : d);
}

// ---- scalar helpers on the 28-int row (host shim; also handy in device tails) -------

template <typename T>
BGX_HD Masks masks_of_row(const T *s)
{
    Masks m = {0, 0, 0, 0};
    for (int i = 0; i < 24; i++) {
        int v = (int)s[i];
        uint32_t bit = 1u << (i + 1);
        if (v > 0) m.occ1 |= bit;
        if (v < 0) m.occ2 |= bit;
        if (v <= -2) m.wall1 |= bit;
        if (v >= 2) m.wall2 |= bit;
    }
    return m;
}

// apply a move the generator produced (no validation): game.cpp:624-659 + Pieces.cpp
template <typename T>
BGX_HD void apply_generated(T *s, int player, int o, int d)
{
    const int m = player == 0 ? 1 : -1;
    if (o == 0 || o == 25) s[24 + player] -= 1;
    else s[o - 1] -= m;
    if (d == 0 || d == 25) { s[26 + player] += 1; return; }
    if ((int)s[d - 1] == -m) { s[d - 1] = 0; s[25 - player] += 1; }
    s[d - 1] += m;
}

// ---- dice: Philox4x32-10, counter-based ------------------------------------------------

BGX_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

struct Philox {
    uint32_t x[4];
};

BGX_HD Philox philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox p = {{c0, c1, c2, c3}};
    return p;
}

BGX_HD int die_of(uint32_t x) { return 1 + (int)mulhi32(x, 6u); }

// ---- the enumeration digest -------------------------------------------------

BGX_HD uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// k[0..3] magnitude bit-planes, k[4] sign plane of the 28-int row; moves = packed sequence
BGX_HD uint64_t leaf_hash(const uint32_t k[5], uint64_t moves_and_len)
{
    uint64_t h = mix64((uint64_t)k[0] | ((uint64_t)k[1] << 32));
    h = mix64(h ^ ((uint64_t)k[2] | ((uint64_t)k[3] << 32)));
    h = mix64(h ^ (uint64_t)k[4]);
    h = mix64(h ^ moves_and_len);
    return h;
}
constexpr uint64_t kDigestMul = 0x9E3779B97F4A7C15ULL;

// packed sequence: move j occupies bits 10j..10j+9 as origin | dest<<5, length in bits 40..42
BGX_HD uint64_t pack_move(int o, int d, int j) { return ((uint64_t)(o | (d << 5))) << (10 * j); }

} // namespace bgx
