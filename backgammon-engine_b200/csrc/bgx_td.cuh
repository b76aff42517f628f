// bgx_td.cuh — exact online TD(lambda) replay, apply_td_updates (train.py:124-172).
//
// One CTA per game, weights AND eligibility traces resident in shared memory
// (2 x 25 601 fp32 = 204.8 KB of the 227 KB a CTA may own), so the 4 x 102 KB of
// read-modify-write traffic per TD step never leaves the SM.  The replay is the
// reference's, step for step: two forwards per step with the CURRENT weights, closed-form
// gradients of the 2-layer sigmoid net, e <- lambda*e + grad,
// p <- p + (lr*delta)*e with the same fp32 roundings torch produces (separate multiply
// and add, lr*delta formed in float64 then rounded to fp32).
#pragma once
#include "bgx_device.cuh"

namespace bgx {

constexpr int kTdThreads = 512;
constexpr int kTdListCap = 48;                     // >= 35 non-zero features of any legal position
// shared memory map (floats)
constexpr int kTdW = 0;                            // W  [198][128] feature-major
constexpr int kTdE = kTdW + kTableFloats;          // E  [198][128]
constexpr int kTdB1 = kTdE + kTableFloats;         // b1, eb1, w2, ew2: 128 each
constexpr int kTdEB1 = kTdB1 + kHidden;
constexpr int kTdW2 = kTdEB1 + kHidden;
constexpr int kTdEW2 = kTdW2 + kHidden;
constexpr int kTdX = kTdEW2 + kHidden;             // dense x of s_t: 198 (+2 pad)
constexpr int kTdPart = kTdX + 200;                // partial z: [2 states][2 halves][128]
constexpr int kTdH = kTdPart + 4 * kHidden;        // hidden activations [2][128]
constexpr int kTdList = kTdH + 2 * kHidden;        // feature lists: idx[2][48] (int) then val[2][48]
constexpr int kTdRed = kTdList + 4 * kTdListCap;   // y partials [2][4 warps] + scalars
constexpr int kTdFloats = kTdRed + 32;
constexpr int kTdSmem = kTdFloats * 4;
static_assert(kTdSmem <= 227 * 1024, "TD kernel shared memory");

struct TdParams {
    const int8_t *traj;        // [n_games][traj_cap][32] pre-move records (byte 28 = turn flag)
    const int8_t *slots;       // [n_games][32], byte 31 = 1 P1 won / 2 P2 won (others: skipped)
    const int32_t *ply;        // [n_games] number of recorded states
    long long n_games;
    int traj_cap;
    double lr;
    float lambda;
    const float *flat;         // snapshot, state_dict order (b1, w2, b2 are read from here)
    const float *wt;           // snapshot W1 transposed [198][128]
    float *partial;            // [gridDim.x][25604] per-CTA sum of (w_final - w_snapshot), feature-major W1
    float *final_weights;      // optional [25604] state_dict order (single-game calls)
    double *sq_errors;         // optional [T-1] (single-game calls)
    unsigned long long *stats; // [3] games replayed, [6] TD steps
    double *dstats;            // [0] sum of squared TD errors
};

// one warp turns one 32-byte record into (a) a compact list of non-zero features and
// (b) optionally the dense x[198] (model.py:111-144)
__device__ __forceinline__ int td_features(const int8_t *rec, int lane, int *idx, float *val, float *dense)
{
    const int b = (int)rec[lane];
    const int v = lane < 28 ? b : 0;
    const int turn = __shfl_sync(kFull, b, 28) ? 1 : 0;
    const int c = v < 0 ? -v : v;
    int nf = 0;
    if (lane < 24) nf = c < 4 ? c : 4;
    else if (lane < 28) nf = v != 0 ? 1 : 0;
    else if (lane == 28) nf = 1;
    int pos = nf;                                   // inclusive prefix sum over lanes
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(kFull, pos, s);
        if (lane >= s) pos += t;
    }
    const int total = __shfl_sync(kFull, pos, 31);
    pos -= nf;
    if (lane < 24) {
        const int base = 8 * lane + (v > 0 ? 0 : 4);
        for (int k = 0; k < nf && pos + k < kTdListCap; k++) {
            idx[pos + k] = base + k;
            val[pos + k] = k < 3 ? 1.0f : (float)(c - 3) * 0.5f;
        }
        if (dense) {
            float2 *d = reinterpret_cast<float2 *>(dense + 8 * lane);
            const float a = c >= 1 ? 1.f : 0.f, bb = c >= 2 ? 1.f : 0.f, cc = c >= 3 ? 1.f : 0.f;
            const float dd = c >= 4 ? (float)(c - 3) * 0.5f : 0.f;
            const bool p1 = v > 0;
            d[0] = p1 ? make_float2(a, bb) : make_float2(0.f, 0.f);
            d[1] = p1 ? make_float2(cc, dd) : make_float2(0.f, 0.f);
            d[2] = p1 ? make_float2(0.f, 0.f) : make_float2(a, bb);
            d[3] = p1 ? make_float2(0.f, 0.f) : make_float2(cc, dd);
        }
    } else if (lane < 28) {
        const float x = lane < 26 ? (float)v * 0.5f : off_feature(v);
        if (nf && pos < kTdListCap) { idx[pos] = 170 + lane; val[pos] = x; }
        if (dense) dense[170 + lane] = x;
    } else if (lane == 28) {
        if (pos < kTdListCap) { idx[pos] = 192 + turn; val[pos] = 1.0f; }
        if (dense) { dense[192] = turn == 0 ? 1.f : 0.f; dense[193] = turn == 0 ? 0.f : 1.f; }
    }
    return total < kTdListCap ? total : kTdListCap;
}

__global__ void __launch_bounds__(kTdThreads, 1) k_td_replay(TdParams p)
{
    extern __shared__ __align__(16) float sm[];
    float *W = sm + kTdW, *E = sm + kTdE;
    float *b1 = sm + kTdB1, *eb1 = sm + kTdEB1, *w2 = sm + kTdW2, *ew2 = sm + kTdEW2;
    float *x = sm + kTdX, *part = sm + kTdPart, *hs = sm + kTdH;
    int *lidx = reinterpret_cast<int *>(sm + kTdList);
    float *lval = sm + kTdList + 2 * kTdListCap;
    float *red = sm + kTdRed;                       // [0..7] y partials, [8] b2, [9] eb2, [10..11] list sizes, [12..13] v
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float lam = p.lambda;

    float *mine = p.partial + (size_t)blockIdx.x * BGX_NPARAMS_PADDED;
    for (int i = tid; i < BGX_NPARAMS_PADDED; i += kTdThreads) mine[i] = 0.f;

    unsigned long long steps = 0, games = 0;
    double sq_sum = 0.0;

    for (long long g = blockIdx.x; g < p.n_games; g += gridDim.x) {
        const int status = (int)p.slots[g * 32 + 31];
        if (status != kP1Won && status != kP2Won) continue;          // still running or truncated
        int T = p.ply[g];
        if (T > p.traj_cap) T = p.traj_cap;
        if (T <= 0) continue;
        const bool p1_won = status == kP1Won;
        const int8_t *traj = p.traj + (size_t)g * p.traj_cap * 32;

        // round snapshot -> shared memory; traces start at zero (train.py:539-540)
        __syncthreads();
        {
            const float4 *src = reinterpret_cast<const float4 *>(p.wt);
            float4 *dw = reinterpret_cast<float4 *>(W), *de = reinterpret_cast<float4 *>(E);
            for (int i = tid; i < kTableFloats / 4; i += kTdThreads) {
                dw[i] = src[i];
                de[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (tid < kHidden) {
                b1[tid] = p.flat[kTableFloats + tid];
                w2[tid] = p.flat[kTableFloats + kHidden + tid];
                eb1[tid] = 0.f;
                ew2[tid] = 0.f;
            }
            if (tid == 0) { red[8] = p.flat[kTableFloats + 2 * kHidden]; red[9] = 0.f; }
        }
        __syncthreads();

        for (int t = 0; t < T; t++) {
            const bool terminal = t == T - 1;
            // (1) features of s_t (list 0 + dense x) and s_{t+1} (list 1)
            if (warp == 0) {
                const int n = td_features(traj + (size_t)t * 32, lane, lidx, lval, x);
                if (lane == 0) red[10] = __int_as_float(n);
            } else if (warp == 1 && !terminal) {
                const int n = td_features(traj + (size_t)(t + 1) * 32, lane, lidx + kTdListCap, lval + kTdListCap, nullptr);
                if (lane == 0) red[11] = __int_as_float(n);
            }
            __syncthreads();
            // (2) both forwards at once: thread = (state s, half, hidden unit j)
            const int s = tid >> 8, half = (tid >> 7) & 1, j = tid & 127;
            if (s == 0 || !terminal) {
                const int n = __float_as_int(red[10 + s]);
                const int *li = lidx + s * kTdListCap;
                const float *lv = lval + s * kTdListCap;
                float z = 0.f;
                for (int k = half; k < n; k += 2) z += lv[k] * W[li[k] * kHidden + j];
                part[(s * 2 + half) * kHidden + j] = z;
            }
            __syncthreads();
            if (half == 0 && (s == 0 || !terminal)) {
                const float z = b1[j] + part[(s * 2) * kHidden + j] + part[(s * 2 + 1) * kHidden + j];
                const float h = sigmoid_f32(z);
                hs[s * kHidden + j] = h;
                float y = w2[j] * h;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(kFull, y, o);
                if (lane == 0) red[s * 4 + (j >> 5)] = y;
            }
            __syncthreads();
            const float v_cur = sigmoid_f32(red[0] + red[1] + red[2] + red[3] + red[8]);
            double delta;
            if (!terminal) {
                const float v_next = sigmoid_f32(red[4] + red[5] + red[6] + red[7] + red[8]);
                delta = (double)__fsub_rn(v_next, v_cur);                    // train.py:160
                if (tid == 0) {
                    sq_sum += delta * delta;
                    if (p.sq_errors) p.sq_errors[t] = delta * delta;         // train.py:162
                }
            } else {
                delta = (p1_won ? 1.0 : 0.0) - (double)v_cur;               // train.py:168
            }
            const float c = (float)(p.lr * delta);                           // train.py:147
            // (3) gradients w.r.t. the pre-update weights; this thread owns hidden units 4*lane..+3
            const float gv = __fmul_rn(__fsub_rn(1.0f, v_cur), v_cur);
            float gh[4], hh[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                hh[k] = hs[4 * lane + k];
                gh[k] = __fmul_rn(__fmul_rn(__fmul_rn(gv, w2[4 * lane + k]), __fsub_rn(1.0f, hh[k])), hh[k]);
            }
            __syncthreads();                                                 // everyone has read w2/h
            // (4) e <- lambda*e + grad ; p <- p + c*e   (train.py:141-147), all 25 601 parameters
            {
                float4 *W4 = reinterpret_cast<float4 *>(W), *E4 = reinterpret_cast<float4 *>(E);
                for (int f = warp; f < kFeatures; f += kTdThreads / 32) {
                    const float xf = x[f];
                    float4 e = E4[f * 32 + lane], w = W4[f * 32 + lane];
                    e.x = __fadd_rn(__fmul_rn(lam, e.x), __fmul_rn(gh[0], xf));
                    e.y = __fadd_rn(__fmul_rn(lam, e.y), __fmul_rn(gh[1], xf));
                    e.z = __fadd_rn(__fmul_rn(lam, e.z), __fmul_rn(gh[2], xf));
                    e.w = __fadd_rn(__fmul_rn(lam, e.w), __fmul_rn(gh[3], xf));
                    w.x = __fadd_rn(w.x, __fmul_rn(c, e.x));
                    w.y = __fadd_rn(w.y, __fmul_rn(c, e.y));
                    w.z = __fadd_rn(w.z, __fmul_rn(c, e.z));
                    w.w = __fadd_rn(w.w, __fmul_rn(c, e.w));
                    E4[f * 32 + lane] = e;
                    W4[f * 32 + lane] = w;
                }
                if (warp == 0) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int jj = 4 * lane + k;
                        const float e = __fadd_rn(__fmul_rn(lam, eb1[jj]), gh[k]);
                        eb1[jj] = e;
                        b1[jj] = __fadd_rn(b1[jj], __fmul_rn(c, e));
                    }
                } else if (warp == 1) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int jj = 4 * lane + k;
                        const float e = __fadd_rn(__fmul_rn(lam, ew2[jj]), __fmul_rn(gv, hh[k]));
                        ew2[jj] = e;
                        w2[jj] = __fadd_rn(w2[jj], __fmul_rn(c, e));
                    }
                } else if (tid == 64) {
                    const float e = __fadd_rn(__fmul_rn(lam, red[9]), gv);
                    red[9] = e;
                    red[8] = __fadd_rn(red[8], __fmul_rn(c, e));
                }
            }
            __syncthreads();
        }
        steps += (unsigned long long)T;
        games++;

        // this game's weight change, accumulated per CTA (feature-major W1, then b1, w2, b2)
        {
            const float4 *w0 = reinterpret_cast<const float4 *>(p.wt);
            const float4 *W4 = reinterpret_cast<const float4 *>(W);
            float4 *acc = reinterpret_cast<float4 *>(mine);
            for (int i = tid; i < kTableFloats / 4; i += kTdThreads) {
                float4 a = acc[i];
                const float4 w = W4[i], o = w0[i];
                a.x += w.x - o.x; a.y += w.y - o.y; a.z += w.z - o.z; a.w += w.w - o.w;
                acc[i] = a;
            }
            if (tid < kHidden) {
                mine[kTableFloats + tid] += b1[tid] - p.flat[kTableFloats + tid];
                mine[kTableFloats + kHidden + tid] += w2[tid] - p.flat[kTableFloats + kHidden + tid];
            }
            if (tid == 0) mine[kTableFloats + 2 * kHidden] += red[8] - p.flat[kTableFloats + 2 * kHidden];
        }
        if (p.final_weights) {
            for (int i = tid; i < kTableFloats; i += kTdThreads) {
                const int f = i >> 7, jj = i & 127;
                p.final_weights[jj * kFeatures + f] = W[i];
            }
            if (tid < kHidden) {
                p.final_weights[kTableFloats + tid] = b1[tid];
                p.final_weights[kTableFloats + kHidden + tid] = w2[tid];
            }
            if (tid == 0) p.final_weights[kTableFloats + 2 * kHidden] = red[8];
        }
    }
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
