// bgx_device.cuh — device-side building blocks of the self-play hot path (sm_100a).
//
// Execution model: ONE WARP PER POSITION.  Lane i (0..27) holds element i of the
// reference's 28-int state row (cppsrc/game.hpp:17-28): lanes 0..23 the board points,
// 24/25 the bar counts, 26/27 the borne-off counts; lanes 28..31 hold 0.  Everything a
// rule needs about the whole board is four __ballot_sync masks (bgx_core.h), a move is
// two predicated register updates, and the turn tree of cppsrc/game.cpp:109-191 is walked
// with warp-uniform control flow (no divergence, no per-thread stacks in memory).  The
// 198-128-1 network (model.py:63-67) is evaluated by the same warp with 4 hidden units
// per lane against a feature-major weight table resident in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bgx_core.h"

namespace bgx {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kFeatures = 198;
constexpr int kHidden = 128;
constexpr int kTableFloats = kFeatures * kHidden;                 // 25 344
constexpr int kTableBytes = kTableFloats * 4;                     // 101 376
constexpr int kFixedRows = kFeatures + 30;                        // + borne-off step rows (bgx_ply.cuh)
constexpr int kFixedInts = kFixedRows * kHidden;                  // 29 184
constexpr int kFixedBytes = kFixedInts * 4;                       // 116 736

// status byte (record byte 31) of a self-play slot
enum { kRunning = 0, kP1Won = 1, kP2Won = 2, kTruncated = 3 };

// ------------------------------------------------------------------------------------
// lane-distributed state
// ------------------------------------------------------------------------------------

__device__ __forceinline__ Masks masks_on_lanes(int v, int player, bool &mover_on_bar)
{
    const uint32_t pos = __ballot_sync(kFull, v > 0);
    const uint32_t neg = __ballot_sync(kFull, v < 0);
    const uint32_t w1 = __ballot_sync(kFull, v <= -2);
    const uint32_t w2 = __ballot_sync(kFull, v >= 2);
    mover_on_bar = (pos >> (24 + player)) & 1u;
    Masks m;
    m.occ1 = (pos & 0xFFFFFFu) << 1;
    m.occ2 = (neg & 0xFFFFFFu) << 1;
    m.wall1 = (w1 & 0xFFFFFFu) << 1;
    m.wall2 = (w2 & 0xFFFFFFu) << 1;
    return m;
}

// apply a generated move (origin code o -> destination code d): cppsrc/game.cpp:624-659
__device__ __forceinline__ int apply_on_lanes(int v, int lane, int player, int o, int d)
{
    const int m = player ? -1 : 1;
    const bool from_bar = (o == 0) | (o == 25);
    const bool off = (d == 0) | (d == 25);
    const int src = from_bar ? 24 + player : o - 1;
    const int dst = off ? 26 + player : d - 1;
    const int dv = __shfl_sync(kFull, v, dst);
    const bool hit = !off && dv == -m;                 // single enemy blot on the landing point
    int nv = v;
    if (lane == src) nv -= from_bar ? 1 : m;
    if (lane == dst) nv = off ? nv + 1 : (hit ? m : nv + m);
    if (hit && lane == 25 - player) nv += 1;           // the enemy's bar count
    return nv;
}

// the five bit-planes of the row: the canonical 160-bit key of a position
__device__ __forceinline__ void key_planes(int v, uint32_t k[5])
{
    const int mag = v < 0 ? -v : v;
    k[0] = __ballot_sync(kFull, mag & 1);
    k[1] = __ballot_sync(kFull, mag & 2);
    k[2] = __ballot_sync(kFull, mag & 4);
    k[3] = __ballot_sync(kFull, mag & 8);
    k[4] = __ballot_sync(kFull, v < 0);
}

__device__ __forceinline__ uint32_t hash_planes(const uint32_t k[5])
{
    uint32_t h = k[0] * 0x9E3779B1u;
    h ^= k[1] * 0x85EBCA77u;
    h ^= k[2] * 0xC2B2AE3Du;
    h ^= k[3] * 0x27D4EB2Fu;
    h ^= k[4] * 0x165667B1u;
    return h ^ (h >> 15);
}

// ------------------------------------------------------------------------------------
// the turn tree: legalTurnSequences (cppsrc/game.cpp:134-191) + collectDoubles (109-131)
// ------------------------------------------------------------------------------------
// Leaf is called as leaf(state_of_this_lane, packed_moves_with_length, length) once per
// legal turn sequence, in the reference's order, duplicates included.  All control flow
// is warp-uniform; the per-depth stack lives in registers.
#define BGX_STK_GET(a, d) ((d) == 0 ? a##0 : (d) == 1 ? a##1 : (d) == 2 ? a##2 : a##3)
#define BGX_STK_SET(a, d, x)          \
    do {                              \
        if ((d) == 0) a##0 = (x);     \
        else if ((d) == 1) a##1 = (x);\
        else if ((d) == 2) a##2 = (x);\
        else a##3 = (x);              \
    } while (0)

template <class Leaf>
__device__ __forceinline__ void walk_turn(int root, int lane, int player, int d1, int d2, Leaf &leaf)
{
    const bool dbl = d1 == d2;
    const int maxlen = dbl ? 4 : 2;
    const int npass = dbl ? 1 : 2;
    for (int pass = 0; pass < npass; pass++) {
        const int dieA = pass ? d2 : d1, dieB = pass ? d1 : d2;
        int cur = root;
        int sv0 = 0, sv1 = 0, sv2 = 0, sv3 = 0;             // node state per depth (this lane)
        uint32_t lg0 = 0, lg1 = 0, lg2 = 0, lg3 = 0;        // origins still to try per depth
        uint64_t prefix = 0;
        int depth = 0;
        bool entering = true;
        for (;;) {
            if (entering) {
                uint32_t legal = 0;
                if (depth < maxlen) {
                    bool on_bar;
                    const Masks mk = masks_on_lanes(cur, player, on_bar);
                    legal = legal_origins(player, (depth & 1) ? dieB : dieA, mk, on_bar ? 1 : 0);
                }
                if (legal == 0) {
                    // a node without a move ends the sequence (game.cpp:117-121, 148-151); the
                    // root of a non-double pass emits nothing
                    if (dbl || depth > 0)
                        leaf(cur, (prefix & ((1ull << (10 * depth)) - 1)) | ((uint64_t)depth << 40), depth);
                    if (depth == 0) break;
                    depth--;
                    cur = BGX_STK_GET(sv, depth);
                    entering = false;
                    continue;
                }
                BGX_STK_SET(lg, depth, legal);
                BGX_STK_SET(sv, depth, cur);
                entering = false;
            }
            const uint32_t rest = BGX_STK_GET(lg, depth);
            if (rest == 0) {
                if (depth == 0) break;
                depth--;
                cur = BGX_STK_GET(sv, depth);
                continue;
            }
            const int o = lowest_bit(rest);
            BGX_STK_SET(lg, depth, rest & (rest - 1));
            const int d = destination(player, o, (depth & 1) ? dieB : dieA);
            cur = apply_on_lanes(cur, lane, player, o, d);
            prefix = (prefix & ~(0x3FFull << (10 * depth))) | pack_move(o, d, depth);
            depth++;
            entering = true;
        }
    }
}

__device__ __forceinline__ float sigmoid_f32(float z) { return 1.0f / (1.0f + expf(-z)); }

__device__ __forceinline__ float off_feature(int k) { return __fdiv_rn((float)k, 15.0f); } // == (float)(k/15.0), k<=15

// ------------------------------------------------------------------------------------
// shared-memory staging of the weight table: one TMA bulk copy per CTA
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Every thread of the CTA must call this; returns once the table is visible.
__device__ __forceinline__ void stage_table(void *dst_smem, const void *src_gmem, uint64_t *bar, int bytes)
{
    const uint32_t b = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(b)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(b), "r"(0)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------
// one greedy / epsilon-greedy ply: make_move (model.py:180-222)
// ------------------------------------------------------------------------------------
struct Choice {
    int v;            // this lane's element of the chosen afterstate
    uint64_t moves;   // packed sequence | length << 40
    float value;
    int n_seq;        // sequences enumerated (reference count)
    int n_scored;     // afterstates sent through the network
    int n_visited;    // tree edges walked (moves applied)
    bool any;         // a sequence exists (N > 0)
};

struct CountLeaf {
    int n = 0;
    __device__ __forceinline__ void operator()(int, uint64_t, int) { n++; }
};

struct PickLeaf {
    int target, n = 0, v = 0;
    uint64_t moves = 0;
    __device__ __forceinline__ explicit PickLeaf(int t) : target(t) {}
    __device__ __forceinline__ void operator()(int vv, uint64_t mv, int)
    {
        if (n == target) { v = vv; moves = mv; }
        n++;
    }
};

// opening position (cppsrc/game.cpp:251) for this lane
__device__ __forceinline__ int start_value(int lane)
{
    // 2 0 0 0 0 -5 0 -3 0 0 0 5 -5 0 0 0 3 0 5 0 0 0 0 -2
    switch (lane) {
    case 0: return 2;
    case 5: return -5;
    case 7: return -3;
    case 11: return 5;
    case 12: return -5;
    case 16: return 3;
    case 18: return 5;
    case 23: return -2;
    default: return 0;
    }
}

// first mover of game `gid`: play_game's roll-off by dice sums (train.py:89-97) on Philox
// stream 1, or the parity rule of benchmark.py:74
__device__ __forceinline__ int first_mover_of(uint32_t k0, uint32_t k1, uint64_t gid, int rule)
{
    if (rule == 1) return (int)(gid & 1);
    for (uint32_t attempt = 0;; attempt++) {
        const Philox r = philox4x32_10(k0, k1, attempt, (uint32_t)gid, (uint32_t)(gid >> 32), 1u);
        const int s1 = die_of(r.x[0]) + die_of(r.x[1]);
        const int s2 = die_of(r.x[2]) + die_of(r.x[3]);
        if (s1 != s2) return s1 > s2 ? 0 : 1;
    }
}

} // namespace bgx
