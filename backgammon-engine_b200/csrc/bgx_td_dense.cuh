// bgx_td.cuh — exact online TD(lambda) replay, apply_td_updates (train.py:124-172).
//
// One CTA per game, 16 warps.  Weights AND eligibility traces of the 198 x 128 first layer live in
// REGISTERS: warp w owns the feature rows f = w, w+16, w+32, ... (13 rows), lane l the hidden units
// 4l..4l+3 of each, i.e. 13 float4 of W1 and 13 float4 of traces per thread (104 of the 128
// registers a thread of a 512-thread CTA may have).  The dense part of a TD step - e <- lambda*e +
// grad and p <- p + (lr*delta)*e over all 25,344 first-layer parameters - is then pure register
// arithmetic; only the two forward passes exchange data (per-warp partial pre-activations through
// shared memory, 16 KB).  The replay is the reference's, step for step: two forwards per step with
// the CURRENT weights, closed-form gradients of the 2-layer sigmoid net,
// lr*delta formed in float64 then rounded to fp32 as torch does.  Two __syncthreads per step.
//
// Blackwell packed fp32: the two forwards run on FFMA2 (fma.rn.f32x2, x stored as {x, x} pairs).  Measured dead ends
// (profiles/r1_td_replay_ncu_summary.md): the trace/weight update on FFMA2 with exact unfused roundings (fma(a,b,-0),
// fma(a,1,b)) is 4 % slower than scalar FMUL/FADD (the FUSED packed update below is 15 % faster), and visiting only the non-zero rows through a switch costs 25 %:
// the step is bound by its dependent phases and two barriers, not by issue slots.  One game per 2-CTA cluster (64 hidden units
// per CTA, output partials exchanged through DSMEM, halves of two games resident per SM) was built and measured too: 7 %
// slower - the overlap of two games' phases gains nothing, the cluster barrier costs ~380 cycles per step.  So was the
// transposed ownership (a warp owns 8 hidden units for all features, the first-layer sums completed inside the warp by a
// transposing shuffle butterfly, one sigmoid per lane, ONE barrier per step): 61.9 against 68.7 M TD steps/s - its chain
// of ~12 dependent shuffles per step is longer than the barrier it removes (gpurun_out/td2).
#pragma once
#include "bgx_td.cuh"

namespace bgx {

// Trace / weight update arithmetic.  1 (default): packed FMAs, e = fma(lambda, e, g*x), w = fma(c, e, w) - one rounding where
// torch's separate multiply and add have two.  0: torch's unfused roundings (__fmul_rn / __fadd_rn).  Neither is bit-identical
// to torch (the forward sums are ordered differently, and a TD error is a difference of two nearly equal values); measured
// against the oracle on the same games both give the same relative error of the weight change (tools/td_err_probe.py:
// 2e-6 .. 1.3e-5 of max|dw| over 12 games either way), and the fused form needs 6 packed instructions per row instead of
// 20 scalar ones: 59.6 -> 68.7 M TD steps/s (gpurun_out/tdfma).
#ifndef BGX_TD_FMA
#define BGX_TD_FMA 1
#endif
// Sum of the 16 per-warp partial pre-activations of a hidden unit.  0: all in float64 (16 F2F + 16 DADD per thread, on the
// step's critical path).  1 (default): four fp32 chains of four, the four chain sums and b1 added in float64 (4 F2F).
// 2: all fp32.  Measured (gpurun_out/td3): 68.4 / 71.6 / 72.3 M TD steps/s; worst |dw - dw_ref| / tolerance over the five
// reference-played golden games 0.75 / 0.75 / 0.85.  Forming c = (float)(lr * delta) without float64 (lr split in two floats,
// one FMA for the exact product error) is bit-identical and 1.5 % slower: the conversions are not what the step waits for.
// One sigmoid instruction stream per warp for the two output values (even lanes: s_t, odd lanes: s_t+1, then two broadcasts)
// instead of two: bit-identical, 72.8 -> 77.9 M TD steps/s - with 4 warps per scheduler the SFU (MUFU.EX2 + MUFU.RCP, quarter
// rate) is what phase (3) queues on.
#ifndef BGX_TD_LANESIG
#define BGX_TD_LANESIG 1
#endif
#ifndef BGX_TD_SUM
#define BGX_TD_SUM 1
#endif
constexpr int kTdDThreads = 512;
constexpr int kTdDWarps = kTdDThreads / 32;
constexpr int kTdDRows = (kFeatures + kTdDWarps - 1) / kTdDWarps;      // 13 feature rows per warp
constexpr int kTdDXStride = 400;                    // dense x of one state, every entry twice ({x, x}: an FFMA2 operand), 198 + pad pairs
// shared memory map (floats)
constexpr int kTdDX = 0;                            // x of three consecutive states, rotating
constexpr int kTdDPart = kTdDX + 3 * kTdDXStride;     // partial pre-activations [16 warps][2 states][128]
constexpr int kTdDH = kTdDPart + kTdDWarps * 2 * kHidden;   // hidden activations [2][128]
constexpr int kTdDB1 = kTdDH + 2 * kHidden;          // b1, eb1: 128 each
constexpr int kTdDEB1 = kTdDB1 + kHidden;
constexpr int kTdDW2 = kTdDEB1 + kHidden;            // w2 double-buffered [2][128] (read and rewritten in the same phase)
constexpr int kTdDEW2 = kTdDW2 + 2 * kHidden;
constexpr int kTdDRed = kTdDEW2 + kHidden;           // [0..7] output partials of the two states, [8..9] b2 (double-buffered), [10] eb2
constexpr int kTdDFloats = kTdDRed + 16;
constexpr int kTdDSmem = kTdDFloats * 4;

// one warp turns one 32-byte record into the dense x[198] (model.py:111-144), every entry stored twice; all are written
__device__ __forceinline__ void td_features(int b, int lane, float *dense)
{
    const int v = lane < 28 ? b : 0;
    const int turn = __shfl_sync(kFull, b, 28) ? 1 : 0;
    const int c = v < 0 ? -v : v;
    float4 *d4 = reinterpret_cast<float4 *>(dense);             // d4[i] = {x[2i], x[2i], x[2i+1], x[2i+1]}
    if (lane < 24) {
        const float a = c >= 1 ? 1.f : 0.f, bb = c >= 2 ? 1.f : 0.f, cc = c >= 3 ? 1.f : 0.f;
        const float dd = c >= 4 ? (float)(c - 3) * 0.5f : 0.f;
        const bool p1 = v > 0;
        const float4 lo = make_float4(a, a, bb, bb), hi = make_float4(cc, cc, dd, dd), z = make_float4(0.f, 0.f, 0.f, 0.f);
        d4[4 * lane + 0] = p1 ? lo : z;
        d4[4 * lane + 1] = p1 ? hi : z;
        d4[4 * lane + 2] = p1 ? z : lo;
        d4[4 * lane + 3] = p1 ? z : hi;
    } else if (lane < 28) {
        const float x = lane < 26 ? (float)v * 0.5f : off_feature(v);
        reinterpret_cast<float2 *>(dense)[170 + lane] = make_float2(x, x);
    } else if (lane == 28) {
        d4[96] = turn == 0 ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 1.f, 1.f);
    }
}

__global__ void __launch_bounds__(kTdDThreads, 1) k_td_replay_dense(TdParams p)
{
    extern __shared__ __align__(16) float sm[];
    float *xs = sm + kTdDX, *part = sm + kTdDPart, *hs = sm + kTdDH;
    float *b1 = sm + kTdDB1, *eb1 = sm + kTdDEB1, *w2 = sm + kTdDW2, *ew2 = sm + kTdDEW2, *red = sm + kTdDRed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float lam = p.lambda;
    const float4 *wt4 = reinterpret_cast<const float4 *>(p.wt);

    float *mine = p.partial + (size_t)blockIdx.x * BGX_NPARAMS_PADDED;
    for (int i = tid; i < BGX_NPARAMS_PADDED; i += kTdDThreads) mine[i] = 0.f;

    unsigned long long steps = 0, games = 0;
    double sq_sum = 0.0;
    float2 W[kTdDRows][2], E[kTdDRows][2];            // this thread's slice of W1 and of its traces (hidden 4l,4l+1 | 4l+2,4l+3)

    for (long long g = blockIdx.x; g < p.n_games; g += gridDim.x) {
        const int status = (int)p.slots[g * 32 + 31];
        if (status != kP1Won && status != kP2Won) continue;          // still running or truncated
        int T = p.ply[g];
        if (T > p.traj_cap) T = p.traj_cap;
        if (T <= 0) continue;
        const bool p1_won = status == kP1Won;
        const int8_t *traj = p.traj + (size_t)g * p.traj_cap * 32;

        // round snapshot -> registers / shared memory; traces start at zero (train.py:539-540)
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kTdDRows; r++) {
            const int f = warp + kTdDWarps * r;
            const float4 w4 = f < kFeatures ? wt4[f * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            W[r][0] = make_float2(w4.x, w4.y); W[r][1] = make_float2(w4.z, w4.w);
            E[r][0] = E[r][1] = make_float2(0.f, 0.f);
        }
        if (tid < kHidden) {
            b1[tid] = p.flat[kTableFloats + tid];
            w2[tid] = p.flat[kTableFloats + kHidden + tid];
            eb1[tid] = 0.f;
            ew2[tid] = 0.f;
        }
        if (tid == 0) { red[8] = p.flat[kTableFloats + 2 * kHidden]; red[10] = 0.f; }
        if (warp == 2) td_features((int)traj[lane], lane, xs);
        if (warp == 3 && T > 1) td_features((int)traj[32 + lane], lane, xs + kTdDXStride);
        int ahead = (warp == 8 && T > 2) ? (int)traj[2 * 32 + lane] : 0;     // warp 8 keeps one record in flight
        __syncthreads();

        for (int t = 0; t < T; t++) {
            const bool terminal = t == T - 1;
            const float *xc = xs + (t % 3) * kTdDXStride, *xn = xs + ((t + 1) % 3) * kTdDXStride;
            const float *w2c = w2 + (t & 1) * kHidden;
            float *w2n = w2 + ((t + 1) & 1) * kHidden;
            // (1) both forwards, first layer: this warp's rows against x(s_t) and x(s_t+1), two hidden units per FFMA2
            {
                float2 z[2][2] = {{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}};
#pragma unroll
                for (int r = 0; r < kTdDRows; r++) {
                    const int f = warp + kTdDWarps * r;
                    if (f < kFeatures) {
                        const float2 xa = *reinterpret_cast<const float2 *>(xc + 2 * f);
                        const float2 xb = terminal ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2 *>(xn + 2 * f);
                        z[0][0] = fma2(xa, W[r][0], z[0][0]); z[0][1] = fma2(xa, W[r][1], z[0][1]);
                        z[1][0] = fma2(xb, W[r][0], z[1][0]); z[1][1] = fma2(xb, W[r][1], z[1][1]);
                    }
                }
                reinterpret_cast<float4 *>(part + (warp * 2 + 0) * kHidden)[lane] = make_float4(z[0][0].x, z[0][0].y, z[0][1].x, z[0][1].y);
                reinterpret_cast<float4 *>(part + (warp * 2 + 1) * kHidden)[lane] = make_float4(z[1][0].x, z[1][0].y, z[1][1].x, z[1][1].y);
            }
            __syncthreads();
            // (2) hidden layer and output partials: thread = (state s, hidden unit j); meanwhile warp 8 encodes s_t+2
            if (tid < 2 * kHidden) {
                const int s = tid >> 7, j = tid & 127;
                if (s == 0 || !terminal) {
#if BGX_TD_SUM == 0
                    double za4[4] = {0.0, 0.0, 0.0, 0.0};         // few-term fp32 partials, summed in float64 (exact): four short chains
#pragma unroll
                    for (int w = 0; w < kTdDWarps; w++) za4[w & 3] += (double)part[(w * 2 + s) * kHidden + j];
                    const double zd = (za4[0] + za4[1]) + (za4[2] + za4[3]);
                    const float h = sigmoid_f32((float)(zd + (double)b1[j]));
#elif BGX_TD_SUM == 1
                    float zf4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int w = 0; w < kTdDWarps; w++) zf4[w & 3] += part[(w * 2 + s) * kHidden + j];
                    const double zd = ((double)zf4[0] + (double)zf4[1]) + ((double)zf4[2] + (double)zf4[3]);
                    const float h = sigmoid_f32((float)(zd + (double)b1[j]));
#else
                    float zf4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int w = 0; w < kTdDWarps; w++) zf4[w & 3] += part[(w * 2 + s) * kHidden + j];
                    const float h = sigmoid_f32(((zf4[0] + zf4[1]) + (zf4[2] + zf4[3])) + b1[j]);
#endif
                    hs[s * kHidden + j] = h;
                    float y = w2c[j] * h;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(kFull, y, o);
                    if (lane == 0) red[s * 4 + (j >> 5)] = y;
                }
            } else if (warp == 8 && t + 2 < T) {
                td_features(ahead, lane, xs + ((t + 2) % 3) * kTdDXStride);
                if (t + 3 < T) ahead = (int)traj[(size_t)(t + 3) * 32 + lane];      // lands during the next step
            }
            __syncthreads();
            // (3) TD error, gradients w.r.t. the pre-update weights
            const float b2c = red[8 + (t & 1)];
#if BGX_TD_LANESIG
            // odd lanes evaluate s_t+1, even lanes s_t: one sigmoid instruction stream per warp instead of two
            const float *rs = red + 4 * (lane & 1);
            const float v_mine = sigmoid_f32(rs[0] + rs[1] + rs[2] + rs[3] + b2c);
            const float v_cur = __shfl_sync(kFull, v_mine, 0);
#else
            const float v_cur = sigmoid_f32(red[0] + red[1] + red[2] + red[3] + b2c);
#endif
            float c;                                                         // (float)(lr * delta), lr a double: train.py:147
            if (!terminal) {
#if BGX_TD_LANESIG
                const float v_next = __shfl_sync(kFull, v_mine, 1);
#else
                const float v_next = sigmoid_f32(red[4] + red[5] + red[6] + red[7] + b2c);
#endif
                const float d = __fsub_rn(v_next, v_cur);                    // train.py:160
                if (tid == 0) {
                    const double delta = (double)d;
                    sq_sum += delta * delta;
                    if (p.sq_errors) p.sq_errors[t] = delta * delta;         // train.py:162
                }
                c = (float)(p.lr * (double)d);
            } else {
                c = (float)(p.lr * ((p1_won ? 1.0 : 0.0) - (double)v_cur));  // train.py:168
            }
            const float gv = __fmul_rn(__fsub_rn(1.0f, v_cur), v_cur);
            float gh[4], hh[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                hh[k] = hs[4 * lane + k];
                gh[k] = __fmul_rn(__fmul_rn(__fmul_rn(gv, w2c[4 * lane + k]), __fsub_rn(1.0f, hh[k])), hh[k]);
            }
            // (4) e <- lambda*e + grad ; p <- p + c*e   (train.py:141-147), all 25 601 parameters
#if BGX_TD_FMA
            {
                const float2 lam2 = make_float2(lam, lam), c2 = make_float2(c, c);
                const float2 g01 = make_float2(gh[0], gh[1]), g23 = make_float2(gh[2], gh[3]);
#pragma unroll
                for (int r = 0; r < kTdDRows; r++) {
                    const int f = warp + kTdDWarps * r;
                    if (f < kFeatures) {
                        const float2 x2 = *reinterpret_cast<const float2 *>(xc + 2 * f);
                        E[r][0] = fma2(lam2, E[r][0], mul2(g01, x2));
                        E[r][1] = fma2(lam2, E[r][1], mul2(g23, x2));
                        W[r][0] = fma2(c2, E[r][0], W[r][0]);
                        W[r][1] = fma2(c2, E[r][1], W[r][1]);
                    }
                }
            }
#else
#pragma unroll
            for (int r = 0; r < kTdDRows; r++) {
                const int f = warp + kTdDWarps * r;
                if (f < kFeatures) {
                    const float xf = xc[2 * f];
                    E[r][0].x = __fadd_rn(__fmul_rn(lam, E[r][0].x), __fmul_rn(gh[0], xf));
                    E[r][0].y = __fadd_rn(__fmul_rn(lam, E[r][0].y), __fmul_rn(gh[1], xf));
                    E[r][1].x = __fadd_rn(__fmul_rn(lam, E[r][1].x), __fmul_rn(gh[2], xf));
                    E[r][1].y = __fadd_rn(__fmul_rn(lam, E[r][1].y), __fmul_rn(gh[3], xf));
                    W[r][0].x = __fadd_rn(W[r][0].x, __fmul_rn(c, E[r][0].x));
                    W[r][0].y = __fadd_rn(W[r][0].y, __fmul_rn(c, E[r][0].y));
                    W[r][1].x = __fadd_rn(W[r][1].x, __fmul_rn(c, E[r][1].x));
                    W[r][1].y = __fadd_rn(W[r][1].y, __fmul_rn(c, E[r][1].y));
                }
            }
#endif
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
                    w2n[jj] = __fadd_rn(w2c[jj], __fmul_rn(c, e));
                }
            } else if (tid == 64) {
                const float e = __fadd_rn(__fmul_rn(lam, red[10]), gv);
                red[10] = e;
                red[8 + ((t + 1) & 1)] = __fadd_rn(b2c, __fmul_rn(c, e));
            }
            // no barrier here: the next step's phase (1) touches only registers, x and `part`
        }
        __syncthreads();
        steps += (unsigned long long)T;
        games++;
        const float *w2f = w2 + (T & 1) * kHidden;
        const float b2f = red[8 + (T & 1)];

        // this game's weight change, accumulated per CTA (feature-major W1, then b1, w2, b2)
        {
            float4 *acc = reinterpret_cast<float4 *>(mine);
#pragma unroll
            for (int r = 0; r < kTdDRows; r++) {
                const int f = warp + kTdDWarps * r;
                if (f < kFeatures) {
                    float4 a = acc[f * 32 + lane];
                    const float4 o = wt4[f * 32 + lane];
                    a.x += W[r][0].x - o.x; a.y += W[r][0].y - o.y; a.z += W[r][1].x - o.z; a.w += W[r][1].y - o.w;
                    acc[f * 32 + lane] = a;
                }
            }
            if (tid < kHidden) {
                mine[kTableFloats + tid] += b1[tid] - p.flat[kTableFloats + tid];
                mine[kTableFloats + kHidden + tid] += w2f[tid] - p.flat[kTableFloats + kHidden + tid];
            }
            if (tid == 0) mine[kTableFloats + 2 * kHidden] += b2f - p.flat[kTableFloats + 2 * kHidden];
        }
        if (p.final_weights) {
#pragma unroll
            for (int r = 0; r < kTdDRows; r++) {
                const int f = warp + kTdDWarps * r;
                if (f < kFeatures) {
                    p.final_weights[(4 * lane + 0) * kFeatures + f] = W[r][0].x;
                    p.final_weights[(4 * lane + 1) * kFeatures + f] = W[r][0].y;
                    p.final_weights[(4 * lane + 2) * kFeatures + f] = W[r][1].x;
                    p.final_weights[(4 * lane + 3) * kFeatures + f] = W[r][1].y;
                }
            }
            if (tid < kHidden) {
                p.final_weights[kTableFloats + tid] = b1[tid];
                p.final_weights[kTableFloats + kHidden + tid] = w2f[tid];
            }
            if (tid == 0) p.final_weights[kTableFloats + 2 * kHidden] = b2f;
        }
    }
    if (tid == 0) {
        atomicAdd(p.stats + 3, games);
        atomicAdd(p.stats + 6, steps);
        atomicAdd(p.dstats, sq_sum);
    }
}

} // namespace bgx
