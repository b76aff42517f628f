// bgx_td.cuh — exact online TD(lambda) replay, apply_td_updates (train.py:124-172).
//
// One game per CTA.  A backgammon position has ~24 of its 198 features non-zero, and of the reference's dense per-step work
// - two forwards, e <- lambda*e + grad, p <- p + (lr*delta)*e over all 25,344 first-layer parameters - only the rows of the
// non-zero features are needed NOW:
//   * the forwards of s_t and s_t+1 read the rows of their non-zero features;
//   * the gradient is non-zero only in the rows of s_t's features;
//   * every other row merely decays (e <- lambda*e) and drifts (w <- w + c_k*e) and is not looked at until its feature
//     fires again.  Such a row is brought up to date LAZILY, when it enters the window of the next states, by replaying
//     the steps it missed one by one with the recorded c_k = (float)(lr*delta_k): the same fp32 operations in the same
//     order as the dense update, so the result is bit-identical to updating all 198 rows every step (which is what
//     torch does), not an algebraic shortcut.  Per row and missed step: 2 packed instructions (FMUL2, FFMA2).
//
// Layout (256 threads, 8 warps, two CTAs = two games per SM): the 198 feature rows are 8-coloured (td_class: the eight
// features of a point land in eight different classes, rotated by the point, so a position's ~35 live rows spread evenly);
// warp = class, lane = four hidden units.  A thread owns element (row, 4 units) of every row of its class for the whole
// kernel - nobody else reads or writes it - so update, lazy replay and the next forward need no synchronisation.
//   * Weights and traces of a game (2 x 101,376 B) live in a per-CTA scratch in GLOBAL memory that never leaves L2, and the
//     ~35 live rows of the moment (35 KB) are L1 hits (96.8 % measured): L1 is the cache of the window, by itself.  Keeping
//     both tables in shared memory instead allows one game per SM (the step is a chain of dependent phases; one game leaves
//     60 % of the issue slots empty) and pushes ~80 KB per step through the same 128 B/cycle port; explicit register slots
//     for the live rows cost more instructions (slot tests, slot moves) than the loads they save (measured: 8 slots 95,
//     4 slots 116, 3 slots 125, none 118 M TD steps/s) - see profiles/r2_td_replay.md for every variant tried.
//   * Every warp lists the live rows of its own class itself, lane r = row r of the class, from the records of the window
//     kept in a shared-memory ring that a cp.async stream fills 8 states ahead: one ballot gives the live set, the feature
//     values travel by shuffle; per step one new feature value per lane (the window moves on by one state).
//   * ONE pass over the live rows per step updates trace and weight AND accumulates, with the new weights, the first-layer
//     partial sums of the next step's two forwards (s_t+1 and s_t+2 have all their non-zero features inside the window).
// Two block barriers per step: class partials -> hidden layer (one sigmoid per lane), output partials -> values.
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
    float *home;               // [gridDim.x][2][198][128] per-CTA home copies of W1 (transposed) and of its traces; L2-resident
    float *final_weights;      // optional [25604] state_dict order (single-game calls)
    double *sq_errors;         // optional [T-1] (single-game calls)
    unsigned long long *stats; // [3] games replayed, [5] truncated games skipped, [6] TD steps, [7] row-steps replayed lazily, [8] live rows summed over the steps
    double *dstats;            // [0] sum of squared TD errors
    unsigned long long *prof;  // k_td_replay<true>: [8 warps][16] cycles per phase, CTA 0 (bgx_td_profile)
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

constexpr int kTdWarps = 4;
constexpr int kTdCtasPerSm = 4;                             // four games per SM: the others hide a game's dependent-issue latency
constexpr int kTdThreads = kTdWarps * 32;
constexpr int kTdClasses = 4;                               // = warps: warp c owns class c

// the fixed 4-colouring of the feature rows: the four features of a point side go to four different classes, rotated by
// the point so that "at least one checker" rows do not pile up in one class.  A class has two rows per point (one per
// side) and one or two of the six tail rows: 50 or 49 rows, two per lane.
__host__ __device__ constexpr int td_class(int f) { return f < 192 ? ((f & 3) + (f >> 3)) & 3 : (f & 3); }

struct __align__(16) TdCtrl {
    long long game;
    double lr;
    float lam;
    int T, won, done;
};

// shared memory map (bytes)
constexpr int kTdOffC = 0;                                           // c_k of every step so far
constexpr int kTdOffZ = kTdOffC + kTdMaxSteps * 4;                   // class partial pre-activations [4][128 units][2 states]
constexpr int kTdOffH = kTdOffZ + kTdClasses * kHidden * 2 * 4;      // hidden activations [128 units][2 states]
constexpr int kTdOffB1 = kTdOffH + kHidden * 2 * 4;                  // b1 [128]
constexpr int kTdOffW2 = kTdOffB1 + kHidden * 4;                     // w2, double-buffered by step parity [2][128]
constexpr int kTdOffRed = kTdOffW2 + 2 * kHidden * 4;                // output partials [2 parities][2 states][4 classes]
constexpr int kTdRing = 8;                                           // records in the ring (a power of two)
constexpr int kTdAhead = 6;                                          // how many states ahead of the step the loader fetches
static_assert(kTdAhead + 2 <= kTdRing && kTdAhead >= 4, "the ring holds s_t-1 .. s_t+kTdAhead while s_t .. s_t+2 are read");
constexpr int kTdOffRing = kTdOffRed + 2 * 2 * kTdClasses * 4;       // records of consecutive states [kTdRing][32]
constexpr int kTdOffOff = kTdOffRing + kTdRing * 32;                       // the borne-off feature k / 15.0 for k = 0..15
constexpr int kTdOffCtrl = kTdOffOff + 16 * 4;
constexpr int kTdSmem = kTdOffCtrl + (int)sizeof(TdCtrl);
// four games per SM inside the 64 KB shared-memory carve-out (1 KB per CTA is the system's, 32 B are static): the other 192 KB
// of the SM's array are the L1 that holds the four games' windows of live rows
constexpr int kTdCarveoutBytes = 64 * 1024;
static_assert(kTdCtasPerSm * (kTdSmem + 32 + 1024) <= kTdCarveoutBytes && kTdOffCtrl % 16 == 0, "k_td_replay: shared memory per CTA");

// loader: 4 bytes of a record, global -> shared, without passing through a register (nothing waits for the load)
__device__ __forceinline__ void td_fetch(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void td_fetch_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void td_fetch_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__device__ __forceinline__ void td_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kTdThreads) : "memory"); }

// What a lane contributes to its warp's class: up to two feature rows (A, B), each defined by the record byte it looks at and
// how that byte becomes the feature value (model.py:111-144), branch-free: d = sign * byte - threshold, then
//   form 0 (step):  d > 0 ? 1 : 0      point features "at least k+1 checkers" (sign picks the side), the two turn flags
//   form 1 (ramp):  max(d, 0) / 2      "(n - 3) / 2" of a point side, bar counts / 2
//   form 2 (table): off[byte]          borne-off counts / 15.0 (the fp32 divide)
//   form 3: no row
// Lane r < 24 speaks for point r: A = the class's feature of the PLAYER1 side (row 8r + k, k = (class - r) & 3), B = the same
// feature of the PLAYER2 side (row 8r + 4 + k).  Lane 24: class 0 -> 192 (turn == 0), 1 -> 193 (turn != 0), 2 -> 194, 3 -> 195
// (bar counts); lane 25: class 0 -> 196, 1 -> 197 (borne-off counts).  Within a class rows ascend with (lane, A before B).
struct TdLaneRow {
    int row, byte, sign, thr, form;                          // row: the feature index, -1 if none
    __device__ __forceinline__ void init(int cls, int lane, int side)
    {
        const int k = (cls - lane) & 3;
        row = 8 * lane + 4 * side + k; byte = lane; sign = side ? -1 : 1; thr = k; form = k < 3 ? 0 : 1;
        if (lane >= 24) {
            row = -1; byte = 31; sign = 1; thr = 0; form = 3;
            if (side == 0 && lane == 24) {
                row = 192 + cls;
                byte = cls < 2 ? 28 : 22 + cls;              // the turn flag | bar counts (bytes 24, 25)
                sign = cls == 0 ? -1 : 1; thr = cls == 0 ? -1 : 0;
                form = cls < 2 ? 0 : 1;
            } else if (side == 0 && lane == 25 && cls < 2) {
                row = 196 + cls; byte = 26 + cls; form = 2;  // borne-off counts (bytes 26, 27)
            }
        }
    }
    __device__ __forceinline__ float value(int v, const float *off) const
    {
        const int d = sign * v - thr;
        const float step = d > 0 ? 1.f : 0.f, ramp = (float)max(d, 0) * 0.5f, tab = off[v & 15];
        return form == 0 ? step : (form == 1 ? ramp : (form == 2 ? tab : 0.f));
    }
};

__device__ __forceinline__ float4 fma4(float s, float4 a, float4 c)      // s * a + c
{
    const float2 lo = fma2(make_float2(s, s), make_float2(a.x, a.y), make_float2(c.x, c.y));
    const float2 hi = fma2(make_float2(s, s), make_float2(a.z, a.w), make_float2(c.z, c.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 mul4(float s, float4 a)
{
    const float2 lo = mul2(make_float2(s, s), make_float2(a.x, a.y)), hi = mul2(make_float2(s, s), make_float2(a.z, a.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// a row that (re-)enters the window replays the steps it slept through, k = from .. upto, exactly as the dense update would
// have done them: e <- fl(lambda e), w <- fma(c_k, e, w).  Four steps per trip with their c_k loaded together (one shared-
// memory latency per trip instead of per step); steps past `upto` run with lambda' = 1, c' = 0, which changes nothing, bit for bit.
__device__ __forceinline__ void td_replay_row(float4 &e, float4 &w, int from, int upto, const float *chist, float lam)
{
    for (int k = from; k <= upto; k += 4) {
        float c[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const bool on = k + i <= upto;
            c[i] = chist[min(k + i, upto)];
            c[i] = on ? c[i] : 0.f;
            l[i] = on ? lam : 1.f;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            e = mul4(l[i], e);
            w = fma4(c[i], e, w);
        }
    }
}

// (float)(lr * (double)d) for a double lr = lr_hi + lr_lo without float64 instructions (three conversions and a DMUL are
// ~120 cycles on the step's critical path): p = fl(lr_hi d), its exact error by FMA, plus the low part's product.  The sum
// p + (err + lr_lo d) carries the double product to ~2^-48 relative and rounds it to fp32; it can differ from the directly
// rounded product only when that product lies within 2^-24 ulp of a rounding boundary (the golden TD tests see none).
__device__ __forceinline__ float td_scale(float lr_hi, float lr_lo, float d)
{
    const float prod = __fmul_rn(lr_hi, d);
    const float err = __fmaf_rn(lr_hi, d, -prod);
    return __fadd_rn(prod, __fmaf_rn(lr_lo, d, err));
}

template <bool kProf>
__global__ void __launch_bounds__(kTdThreads, kTdCtasPerSm) k_td_replay(TdParams p)
{
    extern __shared__ __align__(16) unsigned char td_smem[];
    long long pc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;   // kProf: cycles per phase as seen by lane 0 of every warp
    __shared__ long long td_arrive[kTdWarps];
#define TD_MARK(i) do { if (kProf) { const long long now_ = clock64(); pc[i] += now_ - pt; pt = now_; } } while (0)
    float4 *W4 = reinterpret_cast<float4 *>(p.home + (size_t)blockIdx.x * 2 * kTableFloats), *E4 = W4 + kTableFloats / 4;
    float *chist = reinterpret_cast<float *>(td_smem + kTdOffC);
    float *zpart = reinterpret_cast<float *>(td_smem + kTdOffZ);
    float *hs = reinterpret_cast<float *>(td_smem + kTdOffH);
    float *b1s = reinterpret_cast<float *>(td_smem + kTdOffB1);
    float *w2s = reinterpret_cast<float *>(td_smem + kTdOffW2);
    float *red = reinterpret_cast<float *>(td_smem + kTdOffRed);
    int8_t *ring = reinterpret_cast<int8_t *>(td_smem + kTdOffRing);
    float *offtab = reinterpret_cast<float *>(td_smem + kTdOffOff);
    TdCtrl *ctrl = reinterpret_cast<TdCtrl *>(td_smem + kTdOffCtrl);
    const int tid = threadIdx.x, lane = tid & 31, cls = tid >> 5;            // warp = row class
    const int col = 4 * lane;                                // this thread's hidden units: col .. col + 3
    const bool owner = cls == 0;                             // the class-0 thread of four units also keeps their b1 / w2 entries
    const int unit = 32 * cls + lane;                        // hidden-layer duty: this lane runs both states of one unit
    const bool loader = cls == kTdWarps - 1;                 // this warp feeds the record ring and picks the games
    TdLaneRow meA, meB;
    meA.init(cls, lane, 0);
    meB.init(cls, lane, 1);
    const int idxA = max(meA.row, 0) * 32, idxB = max(meB.row, 0) * 32;      // where this lane's rows start in a table (float4 units)

    float *mine = p.partial + (size_t)blockIdx.x * BGX_NPARAMS_PADDED;
    for (int i = tid; i < BGX_NPARAMS_PADDED; i += kTdThreads) mine[i] = 0.f;
    if (tid < 16) offtab[tid] = off_feature(tid);
    const float4 *wt4 = reinterpret_cast<const float4 *>(p.wt);
    {   // the CTA's home tables start as the snapshot with zero traces; every thread initialises the elements it owns
        const uint32_t haveA = __ballot_sync(kFull, meA.row >= 0), haveB = __ballot_sync(kFull, meB.row >= 0);
        for (int pass = 0; pass < 2; pass++) {
            uint32_t rows = pass ? haveB : haveA;
            while (rows) {
                const int j = __ffs(rows) - 1;
                rows &= rows - 1;
                const int idx = __shfl_sync(kFull, pass ? idxB : idxA, j) + lane;
                W4[idx] = wt4[idx];
                E4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    int lastA = -1, lastB = -1;                              // last step applied to the home copy of this lane's rows, -1 = untouched in this game
    const int8_t *traj = nullptr;                            // loader: the game's records
    long long cursor = blockIdx.x;                           // loader: the next game of this CTA
    unsigned long long steps = 0, games = 0, lazy = 0, rows_live = 0;
    double sq_sum = 0.0;

    // loader: take the next finished game off the queue, publish it, stage its first records
    auto next_game = [&]() {
        td_fetch_wait<0>();                                  // nothing of the previous game is still on its way into the ring
        long long g = -1;
        int T = 0, status = 0;
        for (;;) {                                           // CTA b replays games b, b + grid, ...: a fixed order, so the CTA's
            const long long take = cursor;                   // fp32 sum of weight changes is the same in every run
            cursor += gridDim.x;
            if (take >= p.n_games) break;
            status = (int)p.slots[take * 32 + 31];
            T = min(min(p.ply[take], p.traj_cap), kTdMaxSteps);
            if ((status == kP1Won || status == kP2Won) && T > 0) { g = (long long)take; break; }   // others: still running or truncated
            if (status == kTruncated && lane == 0) atomicAdd(p.stats + 5, 1ull);
        }
        if (lane == 0) {
            ctrl->done = g < 0;
            ctrl->game = g;
            ctrl->T = T;
            ctrl->won = status == kP1Won;
            if (p.sched && g >= 0) {                         // the per-game schedule of train.py:538 / model.py:69-73
                const long long ep = p.episode_first + g;
                ctrl->lr = p.sched[min(ep / 40000, (long long)kTdSchedLen - 1)];
                ctrl->lam = (float)p.sched[kTdSchedLen + min(ep / 30000, (long long)kTdSchedLen - 1)];
            } else {
                ctrl->lr = p.lr;
                ctrl->lam = p.lambda;
            }
        }
        if (g >= 0) {                                        // records 0 .. kTdAhead - 1 (8 lanes x 4 bytes each), complete before the barrier
            traj = p.traj + (size_t)g * p.traj_cap * 32;
            if (lane < 8)
                for (int k = 0; k < kTdAhead; k++) td_fetch(ring + k * 32 + lane * 4, traj + (size_t)min(k, T - 1) * 32 + lane * 4);
            td_fetch_commit();
            td_fetch_wait<0>();
        }
    };
    if (loader) next_game();

    // one pass over the live rows named by `rows` (lane j speaks for the row at idx_lane of lane j, with feature values
    // x0 / x1 / x2 in s_t / s_t+1 / s_t+2): e <- lambda*e + grad ; w <- w + c*e (train.py:141-147), and with the NEW weights the
    // first layer of step t+1.  Two rows per trip: their loads travel together.
    auto row_pass = [&](uint32_t rows, int idx_lane, float x0, float x1, float x2, float lam, float c, const float4 &gh, float4 &z0, float4 &z1) {
        while (rows) {
            const int ja = __ffs(rows) - 1;
            rows &= rows - 1;
            const bool two = rows != 0;
            const int jb = two ? __ffs(rows) - 1 : ja;
            rows &= rows - 1;
            const int ia = __shfl_sync(kFull, idx_lane, ja) + lane, ib = __shfl_sync(kFull, idx_lane, jb) + lane;
            float4 ea = E4[ia], wa = W4[ia], eb = E4[ib], wb = W4[ib];
            ea = fma4(lam, ea, mul4(__shfl_sync(kFull, x0, ja), gh));
            wa = fma4(c, ea, wa);
            E4[ia] = ea;
            W4[ia] = wa;
            z0 = fma4(__shfl_sync(kFull, x1, ja), wa, z0);
            z1 = fma4(__shfl_sync(kFull, x2, ja), wa, z1);
            if (two) {
                eb = fma4(lam, eb, mul4(__shfl_sync(kFull, x0, jb), gh));
                wb = fma4(c, eb, wb);
                E4[ib] = eb;
                W4[ib] = wb;
                z0 = fma4(__shfl_sync(kFull, x1, jb), wb, z0);
                z1 = fma4(__shfl_sync(kFull, x2, jb), wb, z1);
            }
        }
    };
    // rows that enter the window after a sleep replay what they missed (first missed step .. upto) in their home copies
    auto replay_late = [&](uint32_t late, int idx_lane, int last, int upto, float lam) {
        while (late) {
            const int j = __ffs(late) - 1;
            late &= late - 1;
            const int from = __shfl_sync(kFull, last, j) + 1;
            const int idx = __shfl_sync(kFull, idx_lane, j) + lane;
            float4 e = E4[idx], w = W4[idx];
            td_replay_row(e, w, from, upto, chist, lam);
            E4[idx] = e;
            W4[idx] = w;
            lazy += (unsigned)(upto + 1 - from);
        }
    };

    for (;;) {
        td_bar();                                            // game start: ctrl and the first records are ready
        if (ctrl->done) break;
        const int T = ctrl->T;
        const bool p1_won = ctrl->won != 0;
        const float lam = ctrl->lam;
        const double lr = ctrl->lr;
        const float lr_hi = (float)lr, lr_lo = (float)(lr - (double)lr_hi);
        lastA = lastB = -1;
        if (owner) {
            *reinterpret_cast<float4 *>(b1s + col) = *reinterpret_cast<const float4 *>(p.flat + kTableFloats + col);
            *reinterpret_cast<float4 *>(w2s + col) = *reinterpret_cast<const float4 *>(p.flat + kTableFloats + kHidden + col);
        }
        float b2 = p.flat[kTableFloats + 2 * kHidden], eb2 = 0.f;            // every thread keeps its own copy
        float4 eb1 = make_float4(0.f, 0.f, 0.f, 0.f), ew2 = eb1;             // owners only

        // this lane's rows in s_0, s_1, s_2 (0 beyond the game), and the first layer of step 0 against x(s_0) and x(s_1)
        const int r0 = (int)ring[meA.byte], r1 = (int)ring[32 + meA.byte], r2 = (int)ring[64 + meA.byte];   // (A and B look at the same byte)
        float xA0 = meA.value(r0, offtab), xA1 = T > 1 ? meA.value(r1, offtab) : 0.f, xA2 = T > 2 ? meA.value(r2, offtab) : 0.f;
        float xB0 = meB.value(r0, offtab), xB1 = T > 1 ? meB.value(r1, offtab) : 0.f, xB2 = T > 2 ? meB.value(r2, offtab) : 0.f;
        float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f), z1 = z0;
        for (int pass = 0; pass < 2; pass++) {
            uint32_t rows = __ballot_sync(kFull, pass ? (xB0 != 0.f || xB1 != 0.f) : (xA0 != 0.f || xA1 != 0.f));
            while (rows) {
                const int j = __ffs(rows) - 1;
                rows &= rows - 1;
                const float4 w = W4[__shfl_sync(kFull, pass ? idxB : idxA, j) + lane];
                z0 = fma4(__shfl_sync(kFull, pass ? xB0 : xA0, j), w, z0);
                z1 = fma4(__shfl_sync(kFull, pass ? xB1 : xA1, j), w, z1);
            }
        }

        if (kProf) pt = clock64();
        for (int t = 0; t < T; t++) {
            const bool terminal = t == T - 1;
            const float *w2c = w2s + (t & 1) * kHidden;
            // (1) class partials of both pre-activations -> shared memory, [class][unit][state]
            {
                float4 *zp = reinterpret_cast<float4 *>(zpart + (cls * kHidden + col) * 2);
                zp[0] = make_float4(z0.x, z1.x, z0.y, z1.y);
                zp[1] = make_float4(z0.z, z1.z, z0.w, z1.w);
            }
            TD_MARK(0);
            td_bar();
            TD_MARK(1);
            // (2) hidden layer: one unit per lane, both states; output partials of the warp's 32 units
            {
                const float2 *zp = reinterpret_cast<const float2 *>(zpart) + unit;
                const float2 za = zp[0 * kHidden], zb = zp[1 * kHidden], zc = zp[2 * kHidden], zd = zp[3 * kHidden];
                const float bias = b1s[unit], w2u = w2c[unit];
                const float h0 = sigmoid_f32(((za.x + zb.x) + (zc.x + zd.x)) + bias), h1 = sigmoid_f32(((za.y + zb.y) + (zc.y + zd.y)) + bias);
                reinterpret_cast<float2 *>(hs)[unit] = make_float2(h0, h1);
                float y0 = w2u * h0, y1 = w2u * h1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    y0 += __shfl_xor_sync(kFull, y0, o);
                    y1 += __shfl_xor_sync(kFull, y1, o);
                }
                if (lane == 0) {
                    red[((t & 1) * 2 + 0) * 4 + cls] = y0;
                    red[((t & 1) * 2 + 1) * 4 + cls] = y1;
                }
            }
            TD_MARK(2);
            // (3) the live rows of this class: the window moves on by one state, one new feature value per row
            if (t > 0) {
                const int rn = (int)ring[((t + 2) & (kTdRing - 1)) * 32 + meA.byte];
                xA0 = xA1; xA1 = xA2; xA2 = t + 2 < T ? meA.value(rn, offtab) : 0.f;
                xB0 = xB1; xB1 = xB2; xB2 = t + 2 < T ? meB.value(rn, offtab) : 0.f;
            }
            const bool onA = xA0 != 0.f || xA1 != 0.f || xA2 != 0.f, onB = xB0 != 0.f || xB1 != 0.f || xB2 != 0.f;
            const uint32_t liveA = __ballot_sync(kFull, onA), liveB = __ballot_sync(kFull, onB);
            rows_live += (unsigned)(__popc(liveA) + __popc(liveB));
            // rows that enter the window after a sleep catch up through step t-1
            const uint32_t lateA = __ballot_sync(kFull, onA && lastA >= 0 && lastA < t - 1);
            const uint32_t lateB = __ballot_sync(kFull, onB && lastB >= 0 && lastB < t - 1);
            if (lateA) replay_late(lateA, idxA, lastA, t - 1, lam);
            if (lateB) replay_late(lateB, idxB, lastB, t - 1, lam);
            if (onA) lastA = t;                              // (their update of step t follows in (6))
            if (onB) lastB = t;
            TD_MARK(3);
            if (kProf && lane == 0) td_arrive[cls] = clock64();
            td_bar();                                        // the step's second barrier
            TD_MARK(4);
            if (kProf) {
                long long la = 0;
                for (int w = 0; w < kTdWarps; w++) la = max(la, td_arrive[w]);
                const long long now_ = clock64();
                pc[11] += now_ - la;                         // barrier release after the last arrival
                pc[12] += la - td_arrive[cls];               // this warp's wait for the last arrival
                pt = clock64();
            }
            // (4) values, TD error: odd lanes evaluate s_t+1, even lanes s_t (one sigmoid stream per warp)
            const float4 ra = *reinterpret_cast<const float4 *>(red + ((t & 1) * 2 + (lane & 1)) * 4);
            const float v_mine = sigmoid_f32(((ra.x + ra.y) + (ra.z + ra.w)) + b2);
            const float v_cur = __shfl_sync(kFull, v_mine, 0);
            TD_MARK(8);
            float c;                                         // (float)(lr * delta), lr a double: train.py:147
            if (!terminal) {
                const float v_next = __shfl_sync(kFull, v_mine, 1);
                const float d = __fsub_rn(v_next, v_cur);                    // train.py:160
                if (tid == 0) {
                    const double delta = (double)d;
                    sq_sum += delta * delta;
                    if (p.sq_errors) p.sq_errors[t] = delta * delta;         // train.py:162
                }
                c = td_scale(lr_hi, lr_lo, d);
            } else {
                c = (float)(lr * ((p1_won ? 1.0 : 0.0) - (double)v_cur));    // train.py:168
            }
            if (tid == 0) chist[t] = c;
            TD_MARK(9);
            if (loader) {                                    // the record of s_(t+kTdAhead) starts its way into the ring; s_t+3 has arrived
                if (lane < 8) td_fetch(ring + ((t + kTdAhead) & (kTdRing - 1)) * 32 + lane * 4, traj + (size_t)min(t + kTdAhead, T - 1) * 32 + lane * 4);
                td_fetch_commit();
                td_fetch_wait<kTdAhead - 4>();
            }
            TD_MARK(10);
            // (5) gradients w.r.t. the pre-update weights
            const float gv = __fmul_rn(__fsub_rn(1.0f, v_cur), v_cur);
            const float4 ha = *reinterpret_cast<const float4 *>(hs + col * 2), hb = *reinterpret_cast<const float4 *>(hs + col * 2 + 4);
            const float4 h0 = make_float4(ha.x, ha.z, hb.x, hb.z);           // h(s_t) of col .. col + 3 (the odd entries are h(s_t+1))
            const float4 w2v = *reinterpret_cast<const float4 *>(w2c + col);
            const float4 gh = make_float4(__fmul_rn(__fmul_rn(__fmul_rn(gv, w2v.x), __fsub_rn(1.0f, h0.x)), h0.x),
                                          __fmul_rn(__fmul_rn(__fmul_rn(gv, w2v.y), __fsub_rn(1.0f, h0.y)), h0.y),
                                          __fmul_rn(__fmul_rn(__fmul_rn(gv, w2v.z), __fsub_rn(1.0f, h0.z)), h0.z),
                                          __fmul_rn(__fmul_rn(__fmul_rn(gv, w2v.w), __fsub_rn(1.0f, h0.w)), h0.w));
            // (6) one pass over the live rows (s_t+1 and s_t+2 have all their non-zero features among them)
            z0 = make_float4(0.f, 0.f, 0.f, 0.f); z1 = z0;
            TD_MARK(5);
            row_pass(liveA, idxA, xA0, xA1, xA2, lam, c, gh, z0, z1);
            row_pass(liveB, idxB, xB0, xB1, xB2, lam, c, gh, z0, z1);
            if (owner) {                                     // fc1.bias, fc2.weight: one thread per four units
                eb1.x = __fadd_rn(__fmul_rn(lam, eb1.x), gh.x); eb1.y = __fadd_rn(__fmul_rn(lam, eb1.y), gh.y);
                eb1.z = __fadd_rn(__fmul_rn(lam, eb1.z), gh.z); eb1.w = __fadd_rn(__fmul_rn(lam, eb1.w), gh.w);
                ew2.x = __fadd_rn(__fmul_rn(lam, ew2.x), __fmul_rn(gv, h0.x)); ew2.y = __fadd_rn(__fmul_rn(lam, ew2.y), __fmul_rn(gv, h0.y));
                ew2.z = __fadd_rn(__fmul_rn(lam, ew2.z), __fmul_rn(gv, h0.z)); ew2.w = __fadd_rn(__fmul_rn(lam, ew2.w), __fmul_rn(gv, h0.w));
                float4 b = *reinterpret_cast<float4 *>(b1s + col);
                b.x = __fadd_rn(b.x, __fmul_rn(c, eb1.x)); b.y = __fadd_rn(b.y, __fmul_rn(c, eb1.y));
                b.z = __fadd_rn(b.z, __fmul_rn(c, eb1.z)); b.w = __fadd_rn(b.w, __fmul_rn(c, eb1.w));
                *reinterpret_cast<float4 *>(b1s + col) = b;
                *reinterpret_cast<float4 *>(w2s + ((t + 1) & 1) * kHidden + col) =
                    make_float4(__fadd_rn(w2v.x, __fmul_rn(c, ew2.x)), __fadd_rn(w2v.y, __fmul_rn(c, ew2.y)),
                                __fadd_rn(w2v.z, __fmul_rn(c, ew2.z)), __fadd_rn(w2v.w, __fmul_rn(c, ew2.w)));
            }
            eb2 = __fadd_rn(__fmul_rn(lam, eb2), gv);
            b2 = __fadd_rn(b2, __fmul_rn(c, eb2));
            TD_MARK(6);
        }

        td_bar();                                            // game end: every c_k is visible; the ring and ctrl are free
        if (loader) next_game();                             // its global loads land while the rows are flushed
        // every row the game touched replays what it still misses, adds its change to the CTA's sum (feature-major
        // W1, then b1, w2, b2) and returns to the snapshot with zero traces
        const float *w2f = w2s + (T & 1) * kHidden;
        if (p.final_weights) {                               // single-game calls: the weights after the replay, state_dict order
            if (owner)
                for (int k = 0; k < 4; k++) {
                    p.final_weights[kTableFloats + col + k] = b1s[col + k];
                    p.final_weights[kTableFloats + kHidden + col + k] = w2f[col + k];
                }
            if (tid == 0) p.final_weights[kTableFloats + 2 * kHidden] = b2;
        }
        {
            float4 *acc = reinterpret_cast<float4 *>(mine);
            for (int pass = 0; pass < 2; pass++) {
                const int last = pass ? lastB : lastA, idx_lane = pass ? idxB : idxA, row_lane = pass ? meB.row : meA.row;
                uint32_t touched = __ballot_sync(kFull, last >= 0);
                if (p.final_weights) touched = __ballot_sync(kFull, row_lane >= 0);
                while (touched) {                            // two rows per trip: their global loads travel together
                    int idx[2], from[2], f[2];
                    bool ok[2];
                    float4 e[2], w[2], o[2], a[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        ok[i] = touched != 0;
                        const int j = ok[i] ? __ffs(touched) - 1 : 0;
                        touched &= touched - 1;
                        f[i] = __shfl_sync(kFull, row_lane, j);
                        idx[i] = __shfl_sync(kFull, idx_lane, j) + lane;
                        from[i] = __shfl_sync(kFull, last, j) + 1;
                    }
#pragma unroll
                    for (int i = 0; i < 2; i++)
                        if (ok[i]) { e[i] = E4[idx[i]]; w[i] = W4[idx[i]]; o[i] = wt4[idx[i]]; a[i] = acc[idx[i]]; }
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        if (!ok[i]) continue;
                        if (from[i] > 0 && from[i] <= T - 1) {
                            td_replay_row(e[i], w[i], from[i], T - 1, chist, lam);
                            lazy += (unsigned)(T - from[i]);
                        }
                        if (p.final_weights) {
                            p.final_weights[(col + 0) * kFeatures + f[i]] = w[i].x; p.final_weights[(col + 1) * kFeatures + f[i]] = w[i].y;
                            p.final_weights[(col + 2) * kFeatures + f[i]] = w[i].z; p.final_weights[(col + 3) * kFeatures + f[i]] = w[i].w;
                        }
                        a[i].x += w[i].x - o[i].x; a[i].y += w[i].y - o[i].y; a[i].z += w[i].z - o[i].z; a[i].w += w[i].w - o[i].w;
                        acc[idx[i]] = a[i];
                        W4[idx[i]] = o[i];
                        E4[idx[i]] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            if (owner)
                for (int k = 0; k < 4; k++) {
                    mine[kTableFloats + col + k] += b1s[col + k] - p.flat[kTableFloats + col + k];
                    mine[kTableFloats + kHidden + col + k] += w2f[col + k] - p.flat[kTableFloats + kHidden + col + k];
                }
            if (tid == 0) mine[kTableFloats + 2 * kHidden] += b2 - p.flat[kTableFloats + 2 * kHidden];
        }
        steps += (unsigned long long)T;
        games++;
        TD_MARK(7);
    }
    if (lane == 0) atomicAdd(p.stats + 8, rows_live);                         // row updates of the step passes, all four classes
    if (lane == 0) atomicAdd(p.stats + 7, lazy);                              // row-steps replayed lazily, all four classes
    if (tid == 0) {
        atomicAdd(p.stats + 3, games);
        atomicAdd(p.stats + 6, steps);
        atomicAdd(p.dstats, sq_sum);
    }
    if (kProf && blockIdx.x == 0 && lane == 0) {
        for (int i = 0; i < 14; i++) p.prof[cls * 16 + i] = (unsigned long long)pc[i];
        p.prof[cls * 16 + 15] = steps;
    }
#undef TD_MARK
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
