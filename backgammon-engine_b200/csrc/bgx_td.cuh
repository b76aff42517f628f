// bgx_td.cuh — exact online TD(lambda) replay, apply_td_updates (train.py:124-172).
//
// One game per CTA, weights AND eligibility traces of the 198 x 128 first layer in SHARED MEMORY (2 x 101,376 B of the
// 227 KB a CTA may own), visited SPARSELY: a backgammon position has ~24 of its 198 features non-zero, and of the
// reference's dense per-step work - two forwards, e <- lambda*e + grad, p <- p + (lr*delta)*e over all 25,344 parameters -
// only the rows of the non-zero features are needed NOW:
//   * the forwards of s_t and s_t+1 read the rows of their non-zero features;
//   * the gradient is non-zero only in the rows of s_t's features;
//   * every other row merely decays (e <- lambda*e) and drifts (w <- w + c_k*e) and is not looked at until its feature
//     fires again.  Such a row is brought up to date LAZILY, when it enters the window of the next states, by replaying
//     the steps it missed one by one with the recorded c_k = (float)(lr*delta_k): the same fp32 operations in the same
//     order as the dense update, so the result is bit-identical to updating all 198 rows every step (which is what
//     torch does), not an algebraic shortcut.  The arithmetic per row and step is 2 packed instructions (FMUL2, FFMA2);
//     shared-memory traffic per row is one load and one store however long it slept.
//
// Roles (288 threads): 8 worker warps + 1 lister warp, ONE block-wide barrier per step.
//   worker thread = (pair of hidden units, row class): lane = u + 8q, the pair is 8*warp + u, the class q = 0..3 is a fixed
//     4-colouring of the 198 feature rows (td_class: the four features of a point land in four different classes, so a
//     position's ~45 live rows split evenly).  Element (row f, unit j) of W and E belongs to exactly one thread for the
//     whole kernel - nobody else reads or writes it - so the window update of step t, the lazy catch-up and the forward of
//     step t+1 need no synchronisation between them.  The four class partial sums of a pre-activation meet in two
//     xor-shuffles; the 8 per-warp partial output sums cross warps through 64 bytes of shared memory: the one barrier.
//   lister warp: decodes the trajectory records two states ahead and writes, per step, the compact list of live rows
//     (feature values of s_t and s_t+1, first missed step of rows that re-enter the window), class by class in ascending
//     feature order; at the end of the game the list of every touched row for the final catch-up / accumulate / restore.
//
// The replay is the reference's, step for step: two forwards per step with the CURRENT weights, closed-form gradients of
// the 2-layer sigmoid net, lr*delta formed in float64 then rounded to fp32 as torch does, packed
// FMAs for trace and weight (e = fma(lambda, e, g*x), w = fma(c, e, w)).
#pragma once
#include "bgx_device.cuh"

namespace bgx {

constexpr int kTdMaxSteps = 2048;                  // recorded plies per trajectory the replay accepts (c_k history in shared memory)
constexpr int kTdSchedLen = 64;                    // entries of the lr / lambda schedule tables (model.py:69-73; clamped long before)

struct TdParams {
    const int8_t *traj;        // [n_games][traj_cap][32] pre-move records (byte 28 = turn flag)
    const int8_t *slots;       // [n_games][32], byte 31 = 1 P1 won / 2 P2 won (others: skipped)
    const int32_t *ply;        // [n_games] number of recorded states
    long long n_games;
    int traj_cap;
    double lr;
    float lambda;
    const double *sched;       // NULL: lr / lambda above for every game; else [2][kTdSchedLen]: lr by episode / 40000, lambda by episode / 30000
    long long episode_first;   // episode number (train.py:538: games_done + k + 1) of local game 0
    const float *flat;         // snapshot, state_dict order (b1, w2, b2 are read from here)
    const float *wt;           // snapshot W1 transposed [198][128]
    float *partial;            // [gridDim.x][25604] per-CTA sum of (w_final - w_snapshot), feature-major W1
    float *final_weights;      // optional [25604] state_dict order (single-game calls)
    double *sq_errors;         // optional [T-1] (single-game calls)
    unsigned long long *queue; // work queue: next game index (zeroed by the host)
    unsigned long long *stats; // [3] games replayed, [6] TD steps, [7] row-steps caught up lazily
    double *dstats;            // [0] sum of squared TD errors
    unsigned long long *prof;  // k_td_replay<true>: [16] cycles per phase, CTA 0 (bgx_td_profile)
};

// packed fp32 (sm_100): d = a * b + c on two lanes, one rounding each
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long A, B, Cc, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(Cc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(Cc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(D));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    unsigned long long A, B, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(D) : "l"(A), "l"(B));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(D));
    return d;
}

constexpr int kTdWorkerWarps = 8;
constexpr int kTdWorkers = kTdWorkerWarps * 32;
constexpr int kTdThreads = kTdWorkers + 32;                 // + the lister warp
constexpr int kTdClassCap = 50;                             // a class has 50 or 49 rows

// the fixed 4-colouring of the feature rows: the four features of (point, side) go to four different classes,
// rotated by the point so that "at least one checker" rows do not pile up in one class
__host__ __device__ constexpr int td_class(int f) { return f < 192 ? ((f & 3) + (f >> 3)) & 3 : (f & 3); }
// rows of class q in ascending order: position pos (0 .. td_class_rows(q) - 1) -> feature
__host__ __device__ constexpr int td_row_of(int q, int pos) { return pos < 48 ? 4 * pos + ((q - (pos >> 1)) & 3) : 192 + q + 4 * (pos - 48); }
__host__ __device__ constexpr int td_class_rows(int q) { return q < 2 ? 50 : 49; }

// one live row of step t: its feature value in s_t (gradient), in s_t+1 and s_t+2 (the forwards of step t+1), bookkeeping
struct __align__(16) TdEnt {
    float x0, x1, x2;
    uint32_t meta;             // bits 0..7 row, 9 needs catch-up, 16..31 first missed step
};
constexpr uint32_t kTdLate = 0x200u;
struct __align__(16) TdList {
    TdEnt ent[4][kTdClassCap]; // per class, ascending rows
    int n[4];
};
struct __align__(16) TdCtrl {
    long long game;
    double lr;
    float lam;
    int T, won, done;
};

// shared memory map (bytes)
constexpr int kTdOffW = 0;
constexpr int kTdOffE = kTdOffW + kTableBytes;
constexpr int kTdOffC = kTdOffE + kTableBytes;                       // c_k of every step so far
constexpr int kTdOffLists = kTdOffC + kTdMaxSteps * 4;               // 3 rotating step lists + the end-of-game list
constexpr int kTdOffZ = kTdOffLists + 4 * (int)sizeof(TdList);       // class partial pre-activations [4][128 units][2 states]
constexpr int kTdOffH = kTdOffZ + 4 * kHidden * 2 * 4;               // hidden activations [128 units][2 states]
constexpr int kTdOffB1 = kTdOffH + kHidden * 2 * 4;                  // b1 [128]
constexpr int kTdOffW2 = kTdOffB1 + kHidden * 4;                     // w2, double-buffered by step parity [2][128]
constexpr int kTdOffRed = kTdOffW2 + 2 * kHidden * 4;                // per-warp output partials [2 parities][2 states][8]
constexpr int kTdOffLast = kTdOffRed + 2 * 2 * 8 * 4;                // last step applied to each row, -1 = untouched (lister) [200] int16
constexpr int kTdOffCtrl = kTdOffLast + 200 * 2;
constexpr int kTdSmem = kTdOffCtrl + (int)sizeof(TdCtrl);
static_assert(kTdSmem <= 232448, "k_td_replay: shared memory per CTA");
static_assert(kTdOffCtrl % 16 == 0 && sizeof(TdList) % 16 == 0, "alignment");

__device__ __forceinline__ void td_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(kTdWorkers) : "memory"); }
__device__ __forceinline__ void td_bar_all() { asm volatile("bar.sync 2, %0;" ::"n"(kTdThreads) : "memory"); }

// value of feature k (0..3) of a point side holding c checkers (model.py:111-144): [c>=1, c>=2, c>=3, (c-3)/2]
__device__ __forceinline__ float td_feat(int c, int k) { return k < 3 ? (c > k ? 1.f : 0.f) : (c > 3 ? (float)(c - 3) * 0.5f : 0.f); }

// Lister.  Lane l < 24 holds point l of a record; lanes 24..28 hold the turn flag, the two bar counts and the two borne-off
// counts IN THAT ORDER (record bytes 28, 24, 25, 26, 27), so that within every class the rows a lane can contribute ascend
// with the lane: points give rows 8l + k (PLAYER1 side) and 8l + 4 + k, lane 24 rows 192 / 193, 25: 194, 26: 195, 27: 196, 28: 197.
__device__ __forceinline__ int td_record_lane(int lane) { return lane < 24 ? lane : (lane == 24 ? 28 : (lane < 29 ? lane - 1 : 31)); }

struct TdCand { float x0, x1, x2; int f; bool live; };

__device__ __forceinline__ void td_candidates(int q, int lane, int r0, int r1, int r2, bool has1, bool has2, TdCand &A, TdCand &B)
{
    A.live = B.live = false;
    A.f = B.f = 0;
    A.x0 = A.x1 = A.x2 = B.x0 = B.x1 = B.x2 = 0.f;
    if (lane < 24) {
        const int k = (q - lane) & 3;
        A.f = 8 * lane + k;
        B.f = A.f + 4;
        A.x0 = td_feat(max(r0, 0), k); B.x0 = td_feat(max(-r0, 0), k);
        if (has1) { A.x1 = td_feat(max(r1, 0), k); B.x1 = td_feat(max(-r1, 0), k); }
        if (has2) { A.x2 = td_feat(max(r2, 0), k); B.x2 = td_feat(max(-r2, 0), k); }
    } else if (lane == 24) {
        if (q < 2) {                                         // 192: PLAYER1 to move, 193: PLAYER2 to move
            A.f = 192 + q;
            A.x0 = (r0 != 0) == (q == 1) ? 1.f : 0.f;
            if (has1) A.x1 = (r1 != 0) == (q == 1) ? 1.f : 0.f;
            if (has2) A.x2 = (r2 != 0) == (q == 1) ? 1.f : 0.f;
        }
    } else if (lane < 29) {
        const int f = 169 + lane;                            // 25 -> 194 ... 28 -> 197
        if ((f & 3) == q) {
            A.f = f;
            A.x0 = f < 196 ? (float)r0 * 0.5f : off_feature(r0);
            if (has1) A.x1 = f < 196 ? (float)r1 * 0.5f : off_feature(r1);
            if (has2) A.x2 = f < 196 ? (float)r2 * 0.5f : off_feature(r2);
        }
    }
    A.live = A.x0 != 0.f || A.x1 != 0.f || A.x2 != 0.f;
    B.live = B.x0 != 0.f || B.x1 != 0.f || B.x2 != 0.f;
}

// the live rows of step s = non-zero features of s_s, s_s+1, s_s+2 (r0, r1, r2: this lane's byte of their records)
__device__ __forceinline__ void td_build_list(TdList *L, short *last, int s, int T, int lane, int r0, int r1, int r2)
{
    const bool has1 = s + 1 < T, has2 = s + 2 < T;
    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        TdCand A, B;
        td_candidates(q, lane, r0, r1, r2, has1, has2, A, B);
        const uint32_t mA = __ballot_sync(kFull, A.live), mB = __ballot_sync(kFull, B.live);
        const int at = __popc(mA & below) + __popc(mB & below);
        if (A.live) {
            const int la = last[A.f];
            TdEnt e;
            e.x0 = A.x0; e.x1 = A.x1; e.x2 = A.x2;
            e.meta = (uint32_t)A.f | (la >= 0 && la < s - 1 ? kTdLate : 0u) | ((uint32_t)(la + 1) << 16);
            L->ent[q][at] = e;
            last[A.f] = (short)s;
        }
        if (B.live) {
            const int la = last[B.f];
            TdEnt e;
            e.x0 = B.x0; e.x1 = B.x1; e.x2 = B.x2;
            e.meta = (uint32_t)B.f | (la >= 0 && la < s - 1 ? kTdLate : 0u) | ((uint32_t)(la + 1) << 16);
            L->ent[q][at + (A.live ? 1 : 0)] = e;
            last[B.f] = (short)s;
        }
        if (lane == 0) L->n[q] = __popc(mA) + __popc(mB);
    }
}

// end of the game: every row the game touched, with the first step it still misses; `last` is reset
__device__ __forceinline__ void td_build_final(TdList *L, short *last, int T, int lane)
{
    const uint32_t below = (1u << lane) - 1u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int base = 0;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int pos = lane + 32 * pass;
            const bool ok = pos < td_class_rows(q);
            const int f = ok ? td_row_of(q, pos) : 0;
            const int la = ok ? (int)last[f] : -1;
            const uint32_t mask = __ballot_sync(kFull, la >= 0);
            if (la >= 0) {
                TdEnt e;
                e.x0 = e.x1 = e.x2 = 0.f;
                e.meta = (uint32_t)f | (la < T - 1 ? kTdLate : 0u) | ((uint32_t)(la + 1) << 16);
                L->ent[q][base + __popc(mask & below)] = e;
                last[f] = -1;
            }
            base += __popc(mask);
        }
        if (lane == 0) L->n[q] = base;
    }
}

// worker: rows that (re-)enter the window replay the steps they slept through, k = first missed .. upto, exactly as the
// dense update would have done them: e <- fl(lambda e), w <- fma(c_k, e, w).  Warp-uniform: a warp owns one class.
__device__ __forceinline__ unsigned td_catch_up(const TdEnt *ent, int n, int upto, int pair, float2 *W2, float2 *E2, const float *chist, float2 lam2)
{
    unsigned done = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t m = ent[i].meta;
        if (m & kTdLate) {
            const int idx = (int)(m & 0xFFu) * 64 + pair;
            float2 e = E2[idx], w = W2[idx];
            for (int k = (int)(m >> 16); k <= upto; k++) {
                const float c = chist[k];
                e = mul2(lam2, e);
                w = fma2(make_float2(c, c), e, w);
            }
            done += (unsigned)(upto + 1 - (int)(m >> 16));
            E2[idx] = e;
            W2[idx] = w;
        }
    }
    return done;
}

template <bool kProf>
__global__ void __launch_bounds__(kTdThreads, 1) k_td_replay(TdParams p)
{
    extern __shared__ __align__(16) unsigned char td_smem[];
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;       // kProf: cycles per phase as seen by one worker thread / the lister
#define TD_MARK(i) do { if (kProf) { const long long now_ = clock64(); pc[i] += now_ - pt; pt = now_; } } while (0)
    float2 *W2 = reinterpret_cast<float2 *>(td_smem + kTdOffW), *E2 = reinterpret_cast<float2 *>(td_smem + kTdOffE);
    float *chist = reinterpret_cast<float *>(td_smem + kTdOffC);
    TdList *lists = reinterpret_cast<TdList *>(td_smem + kTdOffLists);
    float *zpart = reinterpret_cast<float *>(td_smem + kTdOffZ);
    float *hs = reinterpret_cast<float *>(td_smem + kTdOffH);
    float *b1s = reinterpret_cast<float *>(td_smem + kTdOffB1);
    float *w2s = reinterpret_cast<float *>(td_smem + kTdOffW2);
    float *red = reinterpret_cast<float *>(td_smem + kTdOffRed);
    short *last = reinterpret_cast<short *>(td_smem + kTdOffLast);
    TdCtrl *ctrl = reinterpret_cast<TdCtrl *>(td_smem + kTdOffCtrl);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float *mine = p.partial + (size_t)blockIdx.x * BGX_NPARAMS_PADDED;
    for (int i = tid; i < BGX_NPARAMS_PADDED; i += kTdThreads) mine[i] = 0.f;

    if (warp == kTdWorkerWarps) {
        // ------------------------------------------------------------------ the lister
        for (int f = lane; f < 200; f += 32) last[f] = -1;
        __syncwarp();
        const int src = td_record_lane(lane);
        for (;;) {
            long long g = -1;
            int T = 0, status = 0;
            for (;;) {                                                       // next finished game of the queue
                unsigned long long take = 0;
                if (lane == 0) take = atomicAdd(p.queue, 1ull);
                take = __shfl_sync(kFull, take, 0);
                if (take >= (unsigned long long)p.n_games) break;
                status = (int)p.slots[take * 32 + 31];
                T = min(min(p.ply[take], p.traj_cap), kTdMaxSteps);
                if ((status == kP1Won || status == kP2Won) && T > 0) { g = (long long)take; break; }   // others: still running or truncated
            }
            if (lane == 0) {
                ctrl->done = g < 0;
                ctrl->game = g;
                ctrl->T = T;
                ctrl->won = status == kP1Won;
                if (p.sched && g >= 0) {                                     // the per-game schedule of train.py:538 / model.py:69-73
                    const long long ep = p.episode_first + g;
                    ctrl->lr = p.sched[min(ep / 40000, (long long)kTdSchedLen - 1)];
                    ctrl->lam = (float)p.sched[kTdSchedLen + min(ep / 30000, (long long)kTdSchedLen - 1)];
                } else {
                    ctrl->lr = p.lr;
                    ctrl->lam = p.lambda;
                }
            }
            if (g < 0) { td_bar_all(); return; }
            const int8_t *traj = p.traj + (size_t)g * p.traj_cap * 32 + src;
            int r0 = (int)traj[0], r1 = T > 1 ? (int)traj[32] : 0, r2 = T > 2 ? (int)traj[64] : 0;
            int a0 = T > 3 ? (int)traj[3 * 32] : 0, a1 = T > 4 ? (int)traj[4 * 32] : 0;   // two records in flight
            td_build_list(lists, last, 0, T, lane, r0, r1, r2);
            td_bar_all();                                                    // game start: ctrl and list 0 are ready
            if (kProf) pt = clock64();
            for (int t = 0; t < T; t++) {
                r0 = r1; r1 = r2; r2 = a0;                                   // states t+1, t+2, t+3
                a0 = a1;
                a1 = t + 5 < T ? (int)traj[(size_t)(t + 5) * 32] : 0;
                __syncwarp();
                if (t + 1 < T) td_build_list(lists + (t + 1) % 3, last, t + 1, T, lane, r0, r1, r2);
                TD_MARK(0);
                td_bar_all();                                                // the step's barrier
                TD_MARK(1);
            }
            __syncwarp();
            td_build_final(lists + 3, last, T, lane);
            td_bar_all();                                                    // game end: the final list is ready
            if (kProf && blockIdx.x == 0 && lane == 0) { atomicAdd(p.prof + 8, (unsigned long long)pc[0]); atomicAdd(p.prof + 9, (unsigned long long)pc[1]); pc[0] = pc[1] = 0; }
        }
    }

    // ---------------------------------------------------------------------- the workers
    const int q = warp & 3;                                  // this warp's row class
    const int pair = 32 * (warp >> 2) + lane;                // this thread's hidden units: col, col + 1
    const int col = 2 * pair;
    const bool owner = q == 0;                               // the class-0 thread of a pair also keeps its b1 / w2 entries
    const int unit = 16 * warp + (lane & 15), sig_state = lane >> 4;   // hidden-layer duty: one (unit, state) sigmoid per lane
    const float2 *wt2 = reinterpret_cast<const float2 *>(p.wt);
    for (int pos = 0; pos < td_class_rows(q); pos++) {
        const int f = td_row_of(q, pos);
        W2[f * 64 + pair] = wt2[f * 64 + pair];
        E2[f * 64 + pair] = make_float2(0.f, 0.f);
    }
    unsigned long long steps = 0, games = 0, lazy = 0;
    double sq_sum = 0.0;

    for (;;) {
        td_bar_all();                                        // game start
        if (ctrl->done) break;
        const int T = ctrl->T;
        const bool p1_won = ctrl->won != 0;
        const float lam = ctrl->lam;
        const double lr = ctrl->lr;
        const float2 lam2 = make_float2(lam, lam);
        if (owner) {
            *reinterpret_cast<float2 *>(b1s + col) = *reinterpret_cast<const float2 *>(p.flat + kTableFloats + col);
            *reinterpret_cast<float2 *>(w2s + col) = *reinterpret_cast<const float2 *>(p.flat + kTableFloats + kHidden + col);
        }
        float b2 = p.flat[kTableFloats + 2 * kHidden], eb2 = 0.f;            // every thread keeps its own copy
        float2 eb1 = make_float2(0.f, 0.f), ew2 = make_float2(0.f, 0.f);     // owners only

        // first layer of step 0: this class's live rows against x(s_0) and x(s_1)
        float2 z0 = make_float2(0.f, 0.f), z1 = make_float2(0.f, 0.f);
        {
            const TdEnt *ent = lists[0].ent[q];
            const int n = lists[0].n[q];
            for (int i = 0; i < n; i++) {
                const TdEnt e = ent[i];
                const float2 w = W2[(int)(e.meta & 0xFFu) * 64 + pair];
                z0 = fma2(make_float2(e.x0, e.x0), w, z0);
                z1 = fma2(make_float2(e.x1, e.x1), w, z1);
            }
        }

        if (kProf) pt = clock64();
        for (int t = 0; t < T; t++) {
            const TdList *L = lists + t % 3;
            const TdEnt *ent = L->ent[q];
            const int n = L->n[q];
            const bool terminal = t == T - 1;
            const float *w2c = w2s + (t & 1) * kHidden;
            // (1) class partials of both pre-activations -> shared memory, [class][unit][state]
            *reinterpret_cast<float4 *>(zpart + (q * kHidden + col) * 2) = make_float4(z0.x, z1.x, z0.y, z1.y);
            TD_MARK(0);
            td_bar_workers();
            TD_MARK(1);
            // (2) hidden layer: lane = (unit, state); output partials of this warp's 16 units
            {
                const float za = zpart[(0 * kHidden + unit) * 2 + sig_state], zb = zpart[(1 * kHidden + unit) * 2 + sig_state];
                const float zc = zpart[(2 * kHidden + unit) * 2 + sig_state], zd = zpart[(3 * kHidden + unit) * 2 + sig_state];
                const float h = sigmoid_f32(((za + zb) + (zc + zd)) + b1s[unit]);
                hs[unit * 2 + sig_state] = h;
                float y = w2c[unit] * h;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) y += __shfl_xor_sync(kFull, y, o);
                if ((lane & 15) == 0) red[((t & 1) * 2 + sig_state) * 8 + warp] = y;
            }
            TD_MARK(2);
            // (3) rows entering the window catch up through step t-1 while the partials travel
            lazy += td_catch_up(ent, n, t - 1, pair, W2, E2, chist, lam2);
            TD_MARK(3);
            td_bar_all();                                    // the step's barrier
            TD_MARK(4);
            // (4) values, TD error: odd lanes evaluate s_t+1, even lanes s_t (one sigmoid stream per warp)
            const float4 ra = *reinterpret_cast<const float4 *>(red + ((t & 1) * 2 + (lane & 1)) * 8);
            const float4 rb = *reinterpret_cast<const float4 *>(red + ((t & 1) * 2 + (lane & 1)) * 8 + 4);
            const float v_mine = sigmoid_f32((((ra.x + ra.y) + (ra.z + ra.w)) + ((rb.x + rb.y) + (rb.z + rb.w))) + b2);
            const float v_cur = __shfl_sync(kFull, v_mine, 0);
            float c;                                         // (float)(lr * delta), lr a double: train.py:147
            if (!terminal) {
                const float v_next = __shfl_sync(kFull, v_mine, 1);
                const float d = __fsub_rn(v_next, v_cur);                    // train.py:160
                if (tid == 0) {
                    const double delta = (double)d;
                    sq_sum += delta * delta;
                    if (p.sq_errors) p.sq_errors[t] = delta * delta;         // train.py:162
                }
                c = (float)(lr * (double)d);
            } else {
                c = (float)(lr * ((p1_won ? 1.0 : 0.0) - (double)v_cur));    // train.py:168
            }
            if (tid == 0) chist[t] = c;
            // (5) gradients w.r.t. the pre-update weights
            const float gv = __fmul_rn(__fsub_rn(1.0f, v_cur), v_cur);
            const float4 hh = *reinterpret_cast<const float4 *>(hs + col * 2);           // h(s_t), h(s_t+1) of col, then of col + 1
            const float2 w2v = *reinterpret_cast<const float2 *>(w2c + col);
            const float2 gh = make_float2(__fmul_rn(__fmul_rn(__fmul_rn(gv, w2v.x), __fsub_rn(1.0f, hh.x)), hh.x),
                                          __fmul_rn(__fmul_rn(__fmul_rn(gv, w2v.y), __fsub_rn(1.0f, hh.z)), hh.z));
            // (6) one pass over the live rows: e <- lambda*e + grad ; w <- w + c*e (train.py:141-147), and with the NEW weights
            // the first layer of step t+1 (s_t+1 and s_t+2 have all their non-zero features among these rows)
            const float2 c2 = make_float2(c, c);
            z0 = make_float2(0.f, 0.f); z1 = make_float2(0.f, 0.f);
            TD_MARK(5);
#pragma unroll 2
            for (int i = 0; i < n; i++) {
                const TdEnt e = ent[i];
                const int idx = (int)(e.meta & 0xFFu) * 64 + pair;
                const float2 tr = fma2(lam2, E2[idx], mul2(gh, make_float2(e.x0, e.x0)));
                const float2 w = fma2(c2, tr, W2[idx]);
                E2[idx] = tr;
                W2[idx] = w;
                z0 = fma2(make_float2(e.x1, e.x1), w, z0);
                z1 = fma2(make_float2(e.x2, e.x2), w, z1);
            }
            if (owner) {                                     // fc1.bias, fc2.weight: one thread per unit pair
                eb1.x = __fadd_rn(__fmul_rn(lam, eb1.x), gh.x); eb1.y = __fadd_rn(__fmul_rn(lam, eb1.y), gh.y);
                ew2.x = __fadd_rn(__fmul_rn(lam, ew2.x), __fmul_rn(gv, hh.x)); ew2.y = __fadd_rn(__fmul_rn(lam, ew2.y), __fmul_rn(gv, hh.z));
                float2 b = *reinterpret_cast<float2 *>(b1s + col);
                b.x = __fadd_rn(b.x, __fmul_rn(c, eb1.x)); b.y = __fadd_rn(b.y, __fmul_rn(c, eb1.y));
                *reinterpret_cast<float2 *>(b1s + col) = b;
                *reinterpret_cast<float2 *>(w2s + ((t + 1) & 1) * kHidden + col) =
                    make_float2(__fadd_rn(w2v.x, __fmul_rn(c, ew2.x)), __fadd_rn(w2v.y, __fmul_rn(c, ew2.y)));
            }
            eb2 = __fadd_rn(__fmul_rn(lam, eb2), gv);
            b2 = __fadd_rn(b2, __fmul_rn(c, eb2));
            TD_MARK(6);
        }

        td_bar_all();                                        // game end: the list of touched rows is ready
        const TdList *F = lists + 3;
        lazy += td_catch_up(F->ent[q], F->n[q], T - 1, pair, W2, E2, chist, lam2);
        const float *w2f = w2s + (T & 1) * kHidden;
        if (p.final_weights) {                               // single-game calls: the weights after the replay, state_dict order
            for (int pos = 0; pos < td_class_rows(q); pos++) {
                const int f = td_row_of(q, pos);
                const float2 w = W2[f * 64 + pair];
                p.final_weights[col * kFeatures + f] = w.x;
                p.final_weights[(col + 1) * kFeatures + f] = w.y;
            }
            if (owner) {
                p.final_weights[kTableFloats + col] = b1s[col]; p.final_weights[kTableFloats + col + 1] = b1s[col + 1];
                p.final_weights[kTableFloats + kHidden + col] = w2f[col]; p.final_weights[kTableFloats + kHidden + col + 1] = w2f[col + 1];
            }
            if (tid == 0) p.final_weights[kTableFloats + 2 * kHidden] = b2;
        }
        // this game's weight change, accumulated per CTA (feature-major W1, then b1, w2, b2); touched rows go back to the snapshot
        {
            float2 *acc = reinterpret_cast<float2 *>(mine);
            const int nF = F->n[q];
            for (int i = 0; i < nF; i++) {
                const int f = (int)(F->ent[q][i].meta & 0xFFu);
                const float2 w = W2[f * 64 + pair], o = wt2[f * 64 + pair];
                float2 a = acc[f * 64 + pair];
                a.x += w.x - o.x; a.y += w.y - o.y;
                acc[f * 64 + pair] = a;
                W2[f * 64 + pair] = o;
                E2[f * 64 + pair] = make_float2(0.f, 0.f);
            }
            if (owner) {
                mine[kTableFloats + col] += b1s[col] - p.flat[kTableFloats + col];
                mine[kTableFloats + col + 1] += b1s[col + 1] - p.flat[kTableFloats + col + 1];
                mine[kTableFloats + kHidden + col] += w2f[col] - p.flat[kTableFloats + kHidden + col];
                mine[kTableFloats + kHidden + col + 1] += w2f[col + 1] - p.flat[kTableFloats + kHidden + col + 1];
            }
            if (tid == 0) mine[kTableFloats + 2 * kHidden] += b2 - p.flat[kTableFloats + 2 * kHidden];
        }
        steps += (unsigned long long)T;
        games++;
        TD_MARK(7);
    }
    if (kProf && blockIdx.x == 0 && tid == 0) {
        for (int i = 0; i < 8; i++) p.prof[i] = (unsigned long long)pc[i];
        p.prof[10] = steps;
    }
#undef TD_MARK
    if (lane == 0 && (warp >> 2) == 0) atomicAdd(p.stats + 7, lazy);          // row-steps replayed lazily, all four classes
    if (tid == 0) {
        atomicAdd(p.stats + 3, games);
        atomicAdd(p.stats + 6, steps);
        atomicAdd(p.dstats, sq_sum);
    }
}

// delta[state_dict order] = sum over CTAs of their feature-major partials
__global__ void k_td_reduce(const float *__restrict__ partial, int n_parts, float *__restrict__ delta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BGX_NPARAMS_PADDED) return;
    int src = i;
    if (i < kTableFloats) {
        const int j = i / kFeatures, f = i % kFeatures;
        src = f * kHidden + j;
    }
    float s = 0.f;
    for (int c = 0; c < n_parts; c++) s += partial[(size_t)c * BGX_NPARAMS_PADDED + src];
    delta[i] = i < BGX_NPARAMS ? s : 0.f;
}

__global__ void k_axpy(float *__restrict__ y, const float *__restrict__ x, float a, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}

} // namespace bgx
