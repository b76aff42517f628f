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
//     exact.  The memo shares the per-warp direct-mapped cache with the scored-afterstate
//     set; it is lossy in the safe direction only (a miss re-walks, never mis-counts).
#pragma once
#include "bgx_device.cuh"

namespace bgx {

constexpr int kPlyEntryWords = 6;        // 4 magnitude planes (+depth tag), sign|generation, sub-tree count

// Per-warp scratch in shared memory.  kSets two-way sets of 6-word entries; lane (way*8 + word)
// owns word `word` of way `way` of every set, so no lane ever reads a word another lane wrote
// and the walk needs no __syncwarp.  Depth 0 (the root) lives in registers; rows 0..2 hold
// depths 1..3.
template <int kSets>
struct __align__(16) PlyScratch {
    float4 zs[3][32];                    // hidden pre-activations of the nodes at depth 1..3 (lane's 4 units)
    int sv[3][32];                       // node state at depth 1..3 (lane's element)
    uint32_t lg[4];                      // origins still to try, per depth
    uint32_t ent[4];                     // N when the node was entered
    uint32_t mv[4];                      // move taken at each depth: origin | dest << 5
    uint32_t pad[4];
    uint32_t cache[kSets * 12];          // two 6-word entries per set
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

template <int kSets>
struct PlyCache {
    uint32_t *slots;
    uint32_t gen;

    __device__ __forceinline__ void reset(uint32_t *p, int lane)
    {
        slots = p;
        gen = 0;
        for (int i = lane; i < kSets * 12; i += 32) p[i] = 0;
    }
    __device__ __forceinline__ void next_ply(int lane)
    {
        gen = (gen + 1) & 0xFFu;
        if (gen == 0) {
            for (int i = lane; i < kSets * 12; i += 32) slots[i] = 0;
            gen = 1;
            __syncwarp();
        }
    }
    // what lane (way*8 + w) expects in word w of a matching entry
    __device__ __forceinline__ uint32_t word_of(const uint32_t k[5], int tag, int lane) const
    {
        const int w = lane & 7;
        return w == 0 ? (k[0] | ((uint32_t)tag << 28)) : w == 1 ? k[1] : w == 2 ? k[2] : w == 3 ? k[3] : (k[4] | (gen << 24));
    }
    // probe both ways of the set with one LDS: bit 0 / bit 1 of the result = way 0 / way 1 matches
    __device__ __forceinline__ uint32_t probe(const uint32_t k[5], int tag, int lane, uint32_t &h, uint32_t &got) const
    {
        h = hash_planes(k) + (uint32_t)tag * 0x9E3779B1u;
        const uint32_t *set = slots + (h & (kSets - 1)) * 12;
        got = (lane < 16 && (lane & 7) < 6) ? set[(lane >> 3) * 6 + (lane & 7)] : 0u;
        const uint32_t same = __ballot_sync(kFull, (lane & 7) >= 5 || got == word_of(k, tag, lane));
        return ((same & 0xFFu) == 0xFFu ? 1u : 0u) | ((same & 0xFF00u) == 0xFF00u ? 2u : 0u);
    }
    // victim: a way left over from an earlier ply if there is one, else pseudo-random
    __device__ __forceinline__ void write(const uint32_t k[5], int tag, uint32_t h, uint32_t got, uint32_t count, int lane) const
    {
        const uint32_t stale = __ballot_sync(kFull, lane < 16 && (lane & 7) == 4 && (got >> 24) != gen);
        const int way = (stale & 0x10u) ? 0 : (stale & 0x1000u) ? 1 : (int)((h >> 20) & 1u);
        uint32_t *e = slots + (h & (kSets - 1)) * 12 + way * 6;
        const int w = lane & 7;
        if ((lane >> 3) == way) {
            if (w < 5) e[w] = word_of(k, tag, lane);
            else if (w == 5) e[5] = count;
        }
    }
    // scored-afterstate set (tag 0): true if this exact state was scored earlier in this ply
    __device__ __forceinline__ bool seen_or_insert(const uint32_t k[5], int lane) const
    {
        uint32_t h, got;
        if (probe(k, 0, lane, h, got)) return true;
        write(k, 0, h, got, 0u, lane);
        return false;
    }
    // memo of interior nodes (tag = depth): sub-tree sequence count, or -1
    __device__ __forceinline__ int lookup(const uint32_t k[5], int tag, int lane) const
    {
        uint32_t h, got;
        const uint32_t hit = probe(k, tag, lane, h, got);
        const int cnt = (int)__shfl_sync(kFull, got, (hit & 1u) ? 5 : 13);
        return hit ? cnt : -1;
    }
    __device__ __forceinline__ void store(const uint32_t k[5], int tag, int count, int lane) const
    {
        uint32_t h, got;
        const uint32_t hit = probe(k, tag, lane, h, got);   // refresh in place if it is still there
        if (hit) {
            if (lane == ((hit & 1u) ? 5 : 13)) slots[(h & (kSets - 1)) * 12 + (lane >> 3) * 6 + 5] = (uint32_t)count;
        } else {
            write(k, tag, h, got, (uint32_t)count, lane);
        }
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
                        uint32_t k[5];
                        key_planes(cur, k);
                        if (!cache.seen_or_insert(k, lane)) {
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
                    uint32_t k[5];
                    key_planes(cur, k);
                    const int below = cache.lookup(k, depth, lane);
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
                    uint32_t k[5];
                    key_planes(S.sv[depth - 1][lane], k);
                    cache.store(k, depth, best.n_seq - (int)S.ent[depth], lane);
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
