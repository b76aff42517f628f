// bgx_ply.cuh — the greedy ply of the fused self-play path, second generation.
//
// Same contract as choose_ply() in bgx_device.cuh (make_move, model.py:180-222: enumerate
// in reference order, score every distinct afterstate, first-index arg-best), three changes
// in how the work is done, all driven by the ncu capture profiles/r1a:
//
//  1. DELTA EVALUATION.  The hidden pre-activation z = W1 x + b1 of a node is kept per tree
//     depth; a move changes 2-4 features (one checker leaves a stack, one joins a stack, a
//     hit also flips the blot and the enemy bar), so z(child) = z(parent) + sum of 2-4 rows
//     of the raw feature-major weight table.  Scoring a new afterstate costs 2-4 LDS.128 +
//     the sigmoid epilogue instead of a walk over all ~16 occupied points.
//  2. TREE STACK IN SHARED MEMORY.  Per-depth node state / z / remaining-origin masks live in
//     a per-warp scratch block indexed by depth; every word is only ever touched by one lane
//     (or written with identical values by all), so the walk needs no __syncwarp.
//  3. MEMOISED DOUBLES.  In a double, the same position is reached at the same depth through
//     many move orders (start 3-3: 536 sequences, 73 distinct afterstates).  Interior nodes
//     at depth >= 2 are remembered with the number of sequences below them; meeting one again
//     adds that count to N and skips the whole sub-tree: all its afterstates were already
//     scored earlier in reference order, so the first-index arg-best is unchanged and N stays
//     exact.  The memo shares the per-warp two-way cache with the scored-afterstate set; it is
//     lossy in the safe direction only (a miss re-walks, never mis-counts).
#pragma once
#include "bgx_device.cuh"

namespace bgx {

constexpr int kPlyEntryBytes = 32;       // one byte per lane: 28 state bytes, node tag, ply generation, sub-tree count, spare

// Per-warp scratch in shared memory.  The cache is kSets two-way sets of 32-byte entries; lane l
// owns byte l of both ways of every set: it alone reads and writes that byte, so the walk needs
// no __syncwarp.  Depth 0 (the root) lives in registers; rows 0..2 hold depths 1..3.
template <int kSets>
struct __align__(16) PlyScratch {
    float4 zs[3][32];                    // hidden pre-activations of the nodes at depth 1..3 (lane's 4 units)
    int sv[3][32];                       // node state at depth 1..3 (lane's element)
    uint32_t lg[4];                      // origins still to try, per depth
    uint32_t ent[4];                     // N when the node was entered
    uint32_t mv[4];                      // move taken at each depth: origin | dest << 5
    uint32_t pad[4];
    uint8_t cache[kSets * 2 * kPlyEntryBytes];
};

__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }

struct PlyEvaluator {
    const float4 *W4;    // shared memory, raw feature-major table Wt[198][32] float4
    float4 b1, w2;
    float b2;

    __device__ __forceinline__ void load_params(const float *b1g, const float *w2g, const float *b2g, int lane)
    {
        b1 = reinterpret_cast<const float4 *>(b1g)[lane];
        w2 = reinterpret_cast<const float4 *>(w2g)[lane];
        b2 = b2g[0];
    }
    __device__ __forceinline__ static void axpy(float4 &z, float a, const float4 &t)
    {
        z.x = fmaf(a, t.x, z.x); z.y = fmaf(a, t.y, z.y); z.z = fmaf(a, t.z, z.z); z.w = fmaf(a, t.w, z.w);
    }
    // full pre-activation of a position (once per ply, for the root), features in ascending order
    __device__ __forceinline__ float4 preactivation(int v, int lane, int turn) const
    {
        const int n = v < 0 ? -v : v;
        const int packed = (8 * lane + (v > 0 ? 0 : 4)) | (n << 8);
        uint32_t occ = __ballot_sync(kFull, lane < 24 && v != 0);
        uint32_t side = __ballot_sync(kFull, lane >= 24 && v != 0);
        float4 z = b1;
        while (occ) {
            const int i = lowest_bit(occ);
            occ &= occ - 1;
            const int p = __shfl_sync(kFull, packed, i);
            const int base = p & 0xFF, cnt = p >> 8;
            const float4 *row = W4 + base * 32 + lane;
            axpy(z, 1.0f, row[0]);
            if (cnt >= 2) axpy(z, 1.0f, row[32]);
            if (cnt >= 3) axpy(z, 1.0f, row[64]);
            if (cnt >= 4) axpy(z, (float)(cnt - 3) * 0.5f, row[96]);
        }
        axpy(z, 1.0f, W4[(192 + turn) * 32 + lane]);
        while (side) {
            const int i = lowest_bit(side);
            side &= side - 1;
            const int c = __shfl_sync(kFull, v, i);
            axpy(z, i < 26 ? (float)c * 0.5f : off_feature(c), W4[(170 + i) * 32 + lane]);
        }
        return z;
    }
    __device__ __forceinline__ float finish(const float4 &z) const
    {
        float y = w2.x * fast_sigmoid(z.x) + w2.y * fast_sigmoid(z.y) + w2.z * fast_sigmoid(z.z) + w2.w * fast_sigmoid(z.w);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) y += __shfl_xor_sync(kFull, y, s);
        return fast_sigmoid(y + b2);
    }
};

// what the last move changed, in table rows (warp-uniform)
struct MoveDelta {
    int row_src, row_dst, row_opp;
    float c_src, c_dst;
    bool hit;
    __device__ __forceinline__ float4 apply(float4 z, const float4 *W4, int lane, int player) const
    {
        PlyEvaluator::axpy(z, c_src, W4[row_src * 32 + lane]);
        PlyEvaluator::axpy(z, c_dst, W4[row_dst * 32 + lane]);
        if (hit) {
            PlyEvaluator::axpy(z, -1.0f, W4[row_opp * 32 + lane]);
            PlyEvaluator::axpy(z, 0.5f, W4[(195 - player) * 32 + lane]);   // the enemy's bar feature
        }
        return z;
    }
};

// apply a generated move on the lanes and describe it as table rows (game.cpp:624-659)
__device__ __forceinline__ int apply_with_delta(int v, int lane, int player, int o, int d, MoveDelta &md)
{
    const int m = player ? -1 : 1;
    const int c_me = player ? 4 : 0;
    const bool from_bar = (o == 0) | (o == 25);
    const bool off = (d == 0) | (d == 25);
    const int src = from_bar ? 24 + player : o - 1;
    const int dst = off ? 26 + player : d - 1;
    const int sval = __shfl_sync(kFull, v, src);
    const int dval = __shfl_sync(kFull, v, dst);
    const bool hit = !off && dval == -m;
    if (from_bar) {
        md.row_src = 194 + player;
        md.c_src = -0.5f;
    } else {
        const int n = sval < 0 ? -sval : sval;
        md.row_src = 8 * src + c_me + (n < 4 ? n : 4) - 1;
        md.c_src = n >= 4 ? -0.5f : -1.0f;
    }
    if (off) {
        md.row_dst = 196 + player;
        md.c_dst = off_feature(dval + 1) - off_feature(dval);
    } else if (hit) {
        md.row_dst = 8 * dst + c_me;
        md.c_dst = 1.0f;
        md.row_opp = 8 * dst + (4 - c_me);
    } else {
        const int k = (dval < 0 ? -dval : dval) + 1;
        md.row_dst = 8 * dst + c_me + (k < 4 ? k : 4) - 1;
        md.c_dst = k >= 4 ? 0.5f : 1.0f;
    }
    md.hit = hit;
    int nv = v;
    if (lane == src) nv -= from_bar ? 1 : m;
    if (lane == dst) nv = off ? nv + 1 : (hit ? m : nv + m);
    if (hit && lane == 25 - player) nv += 1;
    return nv;
}

// The per-warp cache of one ply.  Two kinds of entries share it:
//   tag 0      an afterstate that was already scored in this ply (the exact-duplicate filter)
//   tag 2, 3   an interior node of a double at that depth, with the number of sequences below it
// An entry is the position itself, one byte per lane, so a probe is one byte load per way, one
// compare and one vote; the set index comes from a single warp-wide integer reduction
// (REDUX.SUM) of value x per-lane multiplier.  A hit is an exact match of all 28 state bytes,
// the tag and the ply generation, so the cache can only lose time, never exactness.
template <int kSets>
struct PlyCache {
    uint8_t *slots;
    uint32_t gen;        // 1..255, bumped per ply; stale entries are the preferred victims
    uint32_t kmul;       // this lane's hash multiplier

    struct Probe {
        uint32_t off;    // byte offset of the set
        uint32_t m0, m1; // per-lane equality masks of way 0 / way 1
        uint32_t h, got0, got1;
        __device__ __forceinline__ bool hit() const { return m0 == kFull || m1 == kFull; }
    };

    __device__ __forceinline__ void reset(uint8_t *p, int lane)
    {
        slots = p;
        gen = 0;
        kmul = (0x9E3779B1u * (uint32_t)(2 * lane + 1)) ^ (0x85EBCA77u >> (lane & 7));
        uint32_t *w = reinterpret_cast<uint32_t *>(p);
        for (int i = lane; i < kSets * 2 * kPlyEntryBytes / 4; i += 32) w[i] = 0;
    }
    __device__ __forceinline__ void next_ply(int lane)
    {
        gen = (gen + 1) & 0xFFu;
        if (gen == 0) {
            uint32_t *w = reinterpret_cast<uint32_t *>(slots);
            for (int i = lane; i < kSets * 2 * kPlyEntryBytes / 4; i += 32) w[i] = 0;
            gen = 1;
            __syncwarp();
        }
    }
    // what this lane's byte of a matching entry holds (lanes 30, 31 match anything)
    __device__ __forceinline__ uint32_t byte_of(int v, int tag, int lane) const
    {
        const uint32_t meta = lane == 28 ? (uint32_t)tag : gen;
        return lane < 28 ? ((uint32_t)v & 0xFFu) : meta;
    }
    __device__ __forceinline__ Probe probe(int v, int tag, int lane) const
    {
        Probe p;
        uint32_t h = __reduce_add_sync(kFull, (uint32_t)(v + 16) * kmul) + (uint32_t)tag * 0x9E3779B1u;
        h ^= h >> 15;
        p.h = h;
        p.off = (((h & 0xFFFFu) * (uint32_t)kSets) >> 16) * (2 * kPlyEntryBytes);
        const uint8_t *e = slots + p.off + lane;
        p.got0 = e[0];
        p.got1 = e[kPlyEntryBytes];
        const uint32_t mine = byte_of(v, tag, lane);
        p.m0 = __ballot_sync(kFull, lane >= 30 || p.got0 == mine);
        p.m1 = __ballot_sync(kFull, lane >= 30 || p.got1 == mine);
        return p;
    }
    // victim: the matching way if there is one, else a way left over from an earlier ply, else pseudo-random
    __device__ __forceinline__ void write(const Probe &p, int v, int tag, int count, int lane) const
    {
        const int way = p.m0 == kFull ? 0 : p.m1 == kFull ? 1 : !((p.m0 >> 29) & 1u) ? 0 : !((p.m1 >> 29) & 1u) ? 1 : (int)((p.h >> 20) & 1u);
        const uint32_t mine = lane == 30 ? (uint32_t)count : byte_of(v, tag, lane);
        slots[p.off + way * kPlyEntryBytes + lane] = (uint8_t)mine;
    }
    // scored-afterstate set (tag 0): true if this exact state was scored earlier in this ply
    __device__ __forceinline__ bool seen_or_insert(int v, int lane) const
    {
        const Probe p = probe(v, 0, lane);
        if (p.hit()) return true;
        write(p, v, 0, 0, lane);
        return false;
    }
    // memo of interior nodes (tag = depth): sub-tree sequence count (<= 225), or -1
    __device__ __forceinline__ int lookup(int v, int tag, int lane) const
    {
        const Probe p = probe(v, tag, lane);
        const int cnt = (int)__shfl_sync(kFull, p.m0 == kFull ? p.got0 : p.got1, 30);
        return p.hit() ? cnt : -1;
    }
    __device__ __forceinline__ void store(int v, int tag, int count, int lane) const
    {
        if (count > 255) return;                      // cannot happen (<= 15 x 15 sequences below depth 2); never mis-count
        const Probe p = probe(v, tag, lane);
        write(p, v, tag, count, lane);
    }
};

template <int kSets>
__device__ __forceinline__ Choice greedy_ply(int root, int lane, int player, int d1, int d2, const PlyEvaluator &ev,
                                             PlyScratch<kSets> &S, PlyCache<kSets> &cache)
{
    cache.next_ply(lane);
    const float4 zroot = ev.preactivation(root, lane, player);
    const bool dbl = d1 == d2;
    const int maxlen = dbl ? 4 : 2;
    const int npass = dbl ? 1 : 2;
    Choice best;
    best.v = root; best.moves = 0; best.value = __int_as_float(0x7fc00000); best.n_seq = 0; best.n_scored = 0; best.n_visited = 0; best.any = false;
    int best_len = 0;
    uint32_t best_mv0 = 0, best_mv1 = 0, best_mv2 = 0, best_mv3 = 0;

    for (int pass = 0; pass < npass; pass++) {
        const int dieA = pass ? d2 : d1, dieB = pass ? d1 : d2;
        int depth = 0, cur = root;
        bool entering = true;
        MoveDelta md;
        md.row_src = md.row_dst = md.row_opp = 0; md.c_src = md.c_dst = 0.f; md.hit = false;
        for (;;) {
            if (entering) {
                entering = false;
                uint32_t legal = 0;
                if (depth < maxlen) {
                    bool on_bar;
                    const Masks mk = masks_on_lanes(cur, player, on_bar);
                    legal = legal_origins(player, (depth & 1) ? dieB : dieA, mk, on_bar ? 1 : 0);
                }
                if (legal == 0) {
                    if (dbl || depth > 0) {                       // a sequence ends here
                        best.n_seq++;
                        if (!cache.seen_or_insert(cur, lane)) {
                            const float4 z = depth == 0 ? zroot : md.apply(depth == 1 ? zroot : S.zs[depth - 2][lane], ev.W4, lane, player);
                            const float val = ev.finish(z);
                            best.n_scored++;
                            if (!best.any || (player == 0 ? val > best.value : val < best.value)) {
                                best.any = true; best.value = val; best.v = cur; best_len = depth;
                                best_mv0 = S.mv[0]; best_mv1 = S.mv[1]; best_mv2 = S.mv[2]; best_mv3 = S.mv[3];
                            }
                        }
                    }
                    if (depth == 0) break;
                    depth--;
                    continue;
                }
                if (dbl && depth >= 2) {                           // same position, same dice left: seen before?
                    const int below = cache.lookup(cur, depth, lane);
                    if (below >= 0) {
                        best.n_seq += below;
                        depth--;
                        continue;
                    }
                }
                if (depth > 0) {
                    S.zs[depth - 1][lane] = md.apply(depth == 1 ? zroot : S.zs[depth - 2][lane], ev.W4, lane, player);
                    S.sv[depth - 1][lane] = cur;
                }
                S.lg[depth] = legal;
                S.ent[depth] = (uint32_t)best.n_seq;
            }
            const uint32_t rest = S.lg[depth];
            if (rest == 0) {                                       // all children done
                if (dbl && depth >= 2) {
                    cache.store(S.sv[depth - 1][lane], depth, best.n_seq - (int)S.ent[depth], lane);
                }
                if (depth == 0) break;
                depth--;
                continue;
            }
            const int o = lowest_bit(rest);
            S.lg[depth] = rest & (rest - 1);
            const int d = destination(player, o, (depth & 1) ? dieB : dieA);
            cur = apply_with_delta(depth == 0 ? root : S.sv[depth - 1][lane], lane, player, o, d, md);
            best.n_visited++;
            S.mv[depth] = (uint32_t)(o | (d << 5));
            depth++;
            entering = true;
        }
    }
    if (best.any) {
        uint64_t mv = (uint64_t)best_len << 40;
        if (best_len > 0) mv |= (uint64_t)best_mv0;
        if (best_len > 1) mv |= (uint64_t)best_mv1 << 10;
        if (best_len > 2) mv |= (uint64_t)best_mv2 << 20;
        if (best_len > 3) mv |= (uint64_t)best_mv3 << 30;
        best.moves = mv;
    }
    return best;
}

// greedy or exploring ply; kExplore = false compiles the epsilon path out (smaller, fewer registers)
template <int kSets, bool kExplore>
__device__ __forceinline__ Choice choose_ply_fast(int root, int lane, int player, int d1, int d2, const PlyEvaluator &ev,
                                                  PlyScratch<kSets> &S, PlyCache<kSets> &cache, bool explore, uint32_t u)
{
    if (kExplore && explore) {
        CountLeaf cnt;
        walk_turn(root, lane, player, d1, d2, cnt);
        Choice c;
        c.v = root; c.moves = 0; c.value = __int_as_float(0x7fc00000); c.n_seq = cnt.n; c.n_scored = 0; c.n_visited = 0; c.any = cnt.n > 0;
        if (cnt.n > 0) {
            PickLeaf pick((int)mulhi32(u, (uint32_t)cnt.n));
            walk_turn(root, lane, player, d1, d2, pick);
            c.v = pick.v;
            c.moves = pick.moves;
        }
        return c;
    }
    return greedy_ply<kSets>(root, lane, player, d1, d2, ev, S, cache);
}

} // namespace bgx
