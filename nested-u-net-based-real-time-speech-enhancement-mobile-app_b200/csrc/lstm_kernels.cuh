// LSTM bottleneck of one nested sub-U-Net in ONE kernel (models/proposed.py:305-309, objects :26-63):
//     Reshape [T, F_b*C] -> LSTM(21, return_sequences=True) -> Dense(F_b*C) -> Reshape [T, F_b, C]
// Keras gate order i, f, c, o in kernel (D,84) / recurrent_kernel (21,84) / bias (84); sigmoid for i, f, o, tanh for the
// candidate and the output; the one-frame graph passes initial_state=[h, c] and returns the new pair
// (converter_proposed.py:235, :457).
//
// One CTA (256 threads, two CTAs per SM) per clip / stream walks the clip in blocks of LSTM_TB frames as a four-stage
// software pipeline with one block barrier per block; in pipeline step i
//   warp 4     stage  block i:    x rows (sh16 or fp32) -> fp32 in shared memory          (global-load latency lives here)
//   warps 6-7  dense  block i-3:  y[t][n] = bd[n] + sum_j h[t][j] Wd[j][n], written as sh16 (or fp32) frame rows
//   warps 1,2,3,5 project block i-1 (none of them shares warp 0's scheduler: a warp with a stream of independent FMAs would
//                                 take the issue slots the recurrence's dependent chain needs every few cycles): xw[t][g] = bias[g] + sum_k x[t][k] Wk[k][g]; one thread = 4 frames x 3 gates, so the weight
//                                 matrix is read from L1 four times per block instead of sixteen (that traffic, not the
//                                 FMAs, bounds this stage)
//   warp 0     recur  block i-2:  lane u < 21 owns unit u, keeps its 4 x 21 recurrent weights in registers (as fp32x2 pairs:
//                                 gates i|f and c|o advance with one FFMA2 each) and gets h[k] of the previous step by warp
//                                 shuffle -- no barrier inside the time loop, whose dependent chain is what bounds the kernel:
//                                 three partial sums per gate, sigmoid / tanh on the special-function unit (ex2.approx +
//                                 rcp.approx, |error| < 3e-7: inside the fp32 noise of the surrounding layers), next step's
//                                 projected inputs prefetched.
// A streaming step (T = 1) is one launch instead of three.  Every output depends only on its own clip's frames in time
// order, so results do not depend on how a clip is cut into time chunks (carried h / c) or where it sits in the batch.
#pragma once
#include "sh16_kernels.cuh"

namespace nunet {

constexpr int LSTM_TB = 16;        // frames per pipeline block
constexpr int LSTM_THREADS = 256;
constexpr int LSTM_STAGE_THREADS = 32;   // warp 4
constexpr int LSTM_DENSE_THREADS = 64;   // warps 6-7

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// sigmoid / tanh on the special-function unit.  Both saturate cleanly (e^x -> inf gives 0 resp. 1 - 0 = 1).
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f); }

__device__ __forceinline__ void lstm_step_barrier() {   // pipeline step: all 8 warps (role loops run in lock-step)
    asm volatile("bar.sync 0, %0;" ::"n"(LSTM_THREADS) : "memory");
}

// SH: x and y are sh16 frame rows [F_b][C] (k = f*C + c); else fp32 rows [D].
// Wk4: the input kernel re-packed on the host as [D/4][84][4] (four consecutive k of one gate = one 16-byte load).
template <bool SH>
__global__ void __launch_bounds__(LSTM_THREADS, 2) lstm_block_kernel(const void* __restrict__ xv, const float* __restrict__ Wk4,
                                                                    const float* __restrict__ Wr /*[21][84]*/, const float* __restrict__ bk /*[84]*/,
                                                                    const float* __restrict__ Wd /*[21][D]*/, const float* __restrict__ bd /*[D]*/,
                                                                    float* __restrict__ h_state, float* __restrict__ c_state /*[B][21] or null*/,
                                                                    int zero_init, void* __restrict__ yv, int T, int D, int C) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                  // [2][D][LSTM_TB]  (k-major: the four frames of a thread are one 16-byte word)
    float* xw = xs + 2 * LSTM_TB * D;                  // [2][LSTM_TB][84]
    float* hs = xw + 2 * LSTM_TB * LSTM_GATES;         // [2][LSTM_TB][24]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Fb = D / C, D8 = D >> 3, C8 = C >> 3;
    const uint8_t* xsh = reinterpret_cast<const uint8_t*>(xv) + (size_t)b * T * D * 4;
    const float* xf = reinterpret_cast<const float*>(xv) + (size_t)b * T * D;
    const int nblk = (T + LSTM_TB - 1) / LSTM_TB;
    const int nstep = nblk + 3;

    if (warp == 0) {
        // ------------------------------------------------------------------ recurrence warp: block i-2 in step i
        const int u = (lane < LSTM_UNITS) ? lane : 0;
        float2 wif[LSTM_UNITS], wco[LSTM_UNITS];       // (i, f) and (c, o) recurrent weights of unit u per source unit k
#pragma unroll
        for (int k = 0; k < LSTM_UNITS; ++k) {
            wif[k] = make_float2(__ldg(Wr + k * LSTM_GATES + u), __ldg(Wr + k * LSTM_GATES + LSTM_UNITS + u));
            wco[k] = make_float2(__ldg(Wr + k * LSTM_GATES + 2 * LSTM_UNITS + u), __ldg(Wr + k * LSTM_GATES + 3 * LSTM_UNITS + u));
        }
        float h = 0.0f, c = 0.0f;
        if (!zero_init && h_state && lane < LSTM_UNITS) {
            h = h_state[(size_t)b * LSTM_UNITS + lane];
            c = c_state[(size_t)b * LSTM_UNITS + lane];
        }
        for (int i = 0; i < nstep; ++i) {
            const int blk = i - 2;
            if (blk >= 0 && blk < nblk) {
                const int nt = min(LSTM_TB, T - blk * LSTM_TB);
                const float* xwb = xw + (blk & 1) * LSTM_TB * LSTM_GATES;
                float* hsb = hs + (blk & 1) * LSTM_TB * 24;
                float x0 = xwb[u], x1 = xwb[LSTM_UNITS + u], x2 = xwb[2 * LSTM_UNITS + u], x3 = xwb[3 * LSTM_UNITS + u];
                for (int t = 0; t < nt; ++t) {
                    float2 zif[3], zco[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) zif[j] = zco[j] = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int k = 0; k < LSTM_UNITS; ++k) {
                        const float hk = __shfl_sync(0xffffffffu, h, k);
                        const float2 hh = make_float2(hk, hk);
                        zif[k % 3] = __ffma2_rn(hh, wif[k], zif[k % 3]);
                        zco[k % 3] = __ffma2_rn(hh, wco[k], zco[k % 3]);
                    }
                    const float2 sif = __fadd2_rn(__fadd2_rn(zif[0], zif[1]), zif[2]), sco = __fadd2_rn(__fadd2_rn(zco[0], zco[1]), zco[2]);
                    const float z0 = x0 + sif.x, z1 = x1 + sif.y, z2 = x2 + sco.x, z3 = x3 + sco.y;
                    if (t + 1 < nt) {
                        const float* xn = xwb + (t + 1) * LSTM_GATES + u;
                        x0 = xn[0]; x1 = xn[LSTM_UNITS]; x2 = xn[2 * LSTM_UNITS]; x3 = xn[3 * LSTM_UNITS];
                    }
                    const float gi = fast_sigmoid(z0), gf = fast_sigmoid(z1), gc = fast_tanh(z2), go = fast_sigmoid(z3);
                    c = fmaf(gf, c, gi * gc);
                    h = go * fast_tanh(c);
                    if (lane < LSTM_UNITS) hsb[t * 24 + lane] = h;
                }
            }
            lstm_step_barrier();
        }
        if (h_state && lane < LSTM_UNITS) {
            h_state[(size_t)b * LSTM_UNITS + lane] = h;
            c_state[(size_t)b * LSTM_UNITS + lane] = c;
        }
    } else if (warp != 4 && warp < 6) {
        // ------------------------------------------------------------------ projection warps 1, 2, 3, 5: block i-1 in step i
        const int pt = ((warp < 4) ? warp - 1 : 3) * 32 + lane;
        const int gg = pt % 28, rq = pt / 28;          // gate triple, frame quad (rq < 4 for the 112 working threads)
        for (int i = 0; i < nstep; ++i) {
            const int blk = i - 1;
            if (blk >= 0 && blk < nblk && rq < LSTM_TB / 4) {
                const int nt = min(LSTM_TB, T - blk * LSTM_TB);
                if (rq * 4 < nt) {
                    // acc[j][p]: gate 3 gg + j, frame pair p (frames rq*4 + 2p, +1) -- packed fp32x2 arithmetic
                    float2 acc[3][2];
                    const float b0 = __ldg(bk + 3 * gg), b1 = __ldg(bk + 3 * gg + 1), b2 = __ldg(bk + 3 * gg + 2);
                    acc[0][0] = acc[0][1] = make_float2(b0, b0);
                    acc[1][0] = acc[1][1] = make_float2(b1, b1);
                    acc[2][0] = acc[2][1] = make_float2(b2, b2);
                    const float4* xr = reinterpret_cast<const float4*>(xs + (blk & 1) * LSTM_TB * D) + rq;    // [k][4 frame quads]
                    const float4* wp = reinterpret_cast<const float4*>(Wk4) + 3 * gg;
                    const int D4 = D >> 2;
#pragma unroll 2
                    for (int k4 = 0; k4 < D4; ++k4) {          // every sum runs k ascending
                        const float4 w0 = __ldg(wp + k4 * LSTM_GATES), w1 = __ldg(wp + k4 * LSTM_GATES + 1), w2 = __ldg(wp + k4 * LSTM_GATES + 2);
                        const float wv[3][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}, {w2.x, w2.y, w2.z, w2.w}};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const float4 x4 = xr[(k4 * 4 + kk) * (LSTM_TB / 4)];   // frames past nt hold stale data: their results are never read
                            const float2 xa = make_float2(x4.x, x4.y), xb = make_float2(x4.z, x4.w);
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const float2 ww = make_float2(wv[j][kk], wv[j][kk]);
                                acc[j][0] = __ffma2_rn(xa, ww, acc[j][0]);
                                acc[j][1] = __ffma2_rn(xb, ww, acc[j][1]);
                            }
                        }
                    }
                    float* dst = xw + ((blk & 1) * LSTM_TB + rq * 4) * LSTM_GATES + 3 * gg;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        dst[j] = acc[j][0].x;
                        dst[LSTM_GATES + j] = acc[j][0].y;
                        dst[2 * LSTM_GATES + j] = acc[j][1].x;
                        dst[3 * LSTM_GATES + j] = acc[j][1].y;
                    }
                }
            }
            lstm_step_barrier();
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ staging warp: block i in step i
        for (int i = 0; i < nstep; ++i) {
            if (i < nblk) {
                const int t0 = i * LSTM_TB, nt = min(LSTM_TB, T - t0);
                float* xsb = xs + (i & 1) * LSTM_TB * D;
                if (SH) {
                    // batches of 8 items: all loads of a batch first, then the joins and stores (one global-memory round trip
                    // per batch; D = 128 is one batch per block)
                    for (int base = 0; base < LSTM_TB * D8; base += 8 * LSTM_STAGE_THREADS) {
                        uint4 hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int it = base + lane + j * LSTM_STAGE_THREADS;
                            const int r = it % LSTM_TB, k8 = it / LSTM_TB;
                            if (k8 < D8 && r < nt) {
                                const int f = k8 / C8, c8 = k8 - f * C8;
                                const uint8_t* row = xsh + (size_t)(t0 + r) * D * 4;
                                hi[j] = __ldg(reinterpret_cast<const uint4*>(row + ((size_t)c8 * Fb + f) * 16));
                                lo[j] = __ldg(reinterpret_cast<const uint4*>(row + ((size_t)(C8 + c8) * Fb + f) * 16));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int it = base + lane + j * LSTM_STAGE_THREADS;
                            const int r = it % LSTM_TB, k8 = it / LSTM_TB;
                            if (k8 < D8 && r < nt) {
                                const int f = k8 / C8, c8 = k8 - f * C8;
                                float v[8];
                                join8(hi[j], lo[j], v);
                                float* dst = xsb + (f * C + c8 * 8) * LSTM_TB + r;
#pragma unroll
                                for (int e = 0; e < 8; ++e) dst[e * LSTM_TB] = v[e];
                            }
                        }
                    }
                } else {
                    for (int it = lane; it < LSTM_TB * (D >> 2); it += LSTM_STAGE_THREADS) {
                        const int r = it % LSTM_TB, k4 = it / LSTM_TB;
                        if (r >= nt) continue;
                        const float4 v = __ldg(reinterpret_cast<const float4*>(xf + (size_t)(t0 + r) * D) + k4);
                        float* dst = xsb + (k4 * 4) * LSTM_TB + r;
                        dst[0] = v.x; dst[LSTM_TB] = v.y; dst[2 * LSTM_TB] = v.z; dst[3 * LSTM_TB] = v.w;
                    }
                }
            }
            lstm_step_barrier();
        }
    } else {
        // ------------------------------------------------------------------ dense warps 6-7: block i-3 in step i
        const int wt = (warp - 6) * 32 + lane;
        for (int i = 0; i < nstep; ++i) {
            const int blk = i - 3;
            if (blk >= 0) {
                // item = 8 consecutive outputs of one frame
                const int t0 = blk * LSTM_TB, nt = min(LSTM_TB, T - t0);
                const float* hsb = hs + (blk & 1) * LSTM_TB * 24;
                for (int it = wt; it < nt * D8; it += LSTM_DENSE_THREADS) {
                    const int r = it / D8, k8 = it - r * D8;
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = __ldg(bd + k8 * 8 + e);
                    const float* hr = hsb + r * 24;
#pragma unroll
                    for (int j = 0; j < LSTM_UNITS; ++j) {
                        const float hj = hr[j];
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)j * D + k8 * 8));
                        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)j * D + k8 * 8) + 1);
                        o[0] = fmaf(hj, w0.x, o[0]); o[1] = fmaf(hj, w0.y, o[1]); o[2] = fmaf(hj, w0.z, o[2]); o[3] = fmaf(hj, w0.w, o[3]);
                        o[4] = fmaf(hj, w1.x, o[4]); o[5] = fmaf(hj, w1.y, o[5]); o[6] = fmaf(hj, w1.z, o[6]); o[7] = fmaf(hj, w1.w, o[7]);
                    }
                    if (SH) {
                        const int f = k8 / C8, c8 = k8 - f * C8;
                        sh16_store8(reinterpret_cast<uint8_t*>(yv) + ((size_t)b * T + t0 + r) * D * 4, Fb, C, f, c8, o);
                    } else {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(yv) + ((size_t)b * T + t0 + r) * D + k8 * 8);
                        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            lstm_step_barrier();
        }
    }
}

// Streaming step (T = 1): one CTA per group of LSTM_TB STREAMS.  lstm_block_kernel with T = 1 is one CTA per stream, each
// reading the whole input kernel (D x 84 floats: 1024 streams x 43..86 KB from L2 per launch) for 28 working threads; here the
// sixteen streams of a group share every weight read and all 256 threads work in each of the four phases (stage, project,
// one recurrence step, Dense).  Every sum is formed exactly as in lstm_block_kernel (k ascending from the bias; three partial
// sums per gate in the recurrence; j ascending in the Dense), so a stream's frames equal the offline kernel's bit for bit.
template <bool SH>
__global__ void __launch_bounds__(LSTM_THREADS) lstm_stream_kernel(const void* __restrict__ xv, const float* __restrict__ Wk4,
                                                                  const float* __restrict__ Wr /*[21][84]*/, const float* __restrict__ bk /*[84]*/,
                                                                  const float* __restrict__ Wd /*[21][D]*/, const float* __restrict__ bd /*[D]*/,
                                                                  float* __restrict__ h_state, float* __restrict__ c_state /*[S][21]*/,
                                                                  void* __restrict__ yv, int S, int D, int C) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                                  // [D][LSTM_TB]  (k-major)
    float* xw = xs + LSTM_TB * D;                      // [LSTM_TB][84]
    float* hs = xw + LSTM_TB * LSTM_GATES;             // [LSTM_TB][24]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = blockIdx.x * LSTM_TB, ns = min(LSTM_TB, S - s0);
    const int Fb = D / C, D8 = D >> 3, C8 = C >> 3;
    // ---- stage: x rows of the group -> fp32, k-major
    if (SH) {
        const uint8_t* xsh = reinterpret_cast<const uint8_t*>(xv) + (size_t)s0 * D * 4;
        for (int it = tid; it < LSTM_TB * D8; it += LSTM_THREADS) {
            const int r = it % LSTM_TB, k8 = it / LSTM_TB;
            const int f = k8 / C8, c8 = k8 - f * C8;
            float v[8];
            if (r < ns) {
                sh16_load8(xsh + (size_t)r * D * 4, Fb, C, f, c8, v);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.0f;
            }
            float* dst = xs + (f * C + c8 * 8) * LSTM_TB + r;
#pragma unroll
            for (int e = 0; e < 8; ++e) dst[e * LSTM_TB] = v[e];
        }
    } else {
        const float* xf = reinterpret_cast<const float*>(xv) + (size_t)s0 * D;
        for (int it = tid; it < LSTM_TB * (D >> 2); it += LSTM_THREADS) {
            const int r = it % LSTM_TB, k4 = it / LSTM_TB;
            const float4 v = (r < ns) ? __ldg(reinterpret_cast<const float4*>(xf + (size_t)r * D) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float* dst = xs + (k4 * 4) * LSTM_TB + r;
            dst[0] = v.x; dst[LSTM_TB] = v.y; dst[2 * LSTM_TB] = v.z; dst[3 * LSTM_TB] = v.w;
        }
    }
    __syncthreads();
    // ---- project: thread = (gate triple gg, stream pair sp), the pair packed in fp32x2
    if (tid < 28 * (LSTM_TB / 2)) {
        const int gg = tid % 28, sp = tid / 28;
        float2 acc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float bj = __ldg(bk + 3 * gg + j);
            acc[j] = make_float2(bj, bj);
        }
        const float2* xr = reinterpret_cast<const float2*>(xs) + sp;                  // [k][8 stream pairs]
        const float4* wp = reinterpret_cast<const float4*>(Wk4) + 3 * gg;
        const int D4 = D >> 2;
#pragma unroll 2
        for (int k4 = 0; k4 < D4; ++k4) {
            const float4 w0 = __ldg(wp + k4 * LSTM_GATES), w1 = __ldg(wp + k4 * LSTM_GATES + 1), w2 = __ldg(wp + k4 * LSTM_GATES + 2);
            const float wv[3][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}, {w2.x, w2.y, w2.z, w2.w}};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float2 x2 = xr[(k4 * 4 + kk) * (LSTM_TB / 2)];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[j] = __ffma2_rn(x2, make_float2(wv[j][kk], wv[j][kk]), acc[j]);
            }
        }
        float* dst = xw + (2 * sp) * LSTM_GATES + 3 * gg;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dst[j] = acc[j].x;
            dst[LSTM_GATES + j] = acc[j].y;
        }
    }
    __syncthreads();
    // ---- one recurrence step: warp w takes streams 2w and 2w + 1, lane u < 21 owns unit u
    {
        const int u = (lane < LSTM_UNITS) ? lane : 0;
        float2 wif[LSTM_UNITS], wco[LSTM_UNITS];
#pragma unroll
        for (int k = 0; k < LSTM_UNITS; ++k) {
            wif[k] = make_float2(__ldg(Wr + k * LSTM_GATES + u), __ldg(Wr + k * LSTM_GATES + LSTM_UNITS + u));
            wco[k] = make_float2(__ldg(Wr + k * LSTM_GATES + 2 * LSTM_UNITS + u), __ldg(Wr + k * LSTM_GATES + 3 * LSTM_UNITS + u));
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int r = 2 * warp + q;
            const bool live = r < ns && lane < LSTM_UNITS;
            float h = 0.0f, c = 0.0f;
            if (live) {
                h = h_state[(size_t)(s0 + r) * LSTM_UNITS + lane];
                c = c_state[(size_t)(s0 + r) * LSTM_UNITS + lane];
            }
            float2 zif[3], zco[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) zif[j] = zco[j] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int k = 0; k < LSTM_UNITS; ++k) {
                const float hk = __shfl_sync(0xffffffffu, h, k);
                const float2 hh = make_float2(hk, hk);
                zif[k % 3] = __ffma2_rn(hh, wif[k], zif[k % 3]);
                zco[k % 3] = __ffma2_rn(hh, wco[k], zco[k % 3]);
            }
            const float2 sif = __fadd2_rn(__fadd2_rn(zif[0], zif[1]), zif[2]), sco = __fadd2_rn(__fadd2_rn(zco[0], zco[1]), zco[2]);
            const float* xg = xw + r * LSTM_GATES + u;
            const float z0 = xg[0] + sif.x, z1 = xg[LSTM_UNITS] + sif.y, z2 = xg[2 * LSTM_UNITS] + sco.x, z3 = xg[3 * LSTM_UNITS] + sco.y;
            const float gi = fast_sigmoid(z0), gf = fast_sigmoid(z1), gc = fast_tanh(z2), go = fast_sigmoid(z3);
            c = fmaf(gf, c, gi * gc);
            h = go * fast_tanh(c);
            if (live) {
                hs[r * 24 + lane] = h;
                h_state[(size_t)(s0 + r) * LSTM_UNITS + lane] = h;
                c_state[(size_t)(s0 + r) * LSTM_UNITS + lane] = c;
            }
        }
    }
    __syncthreads();
    // ---- Dense: item = 8 consecutive outputs of one stream
    for (int it = tid; it < ns * D8; it += LSTM_THREADS) {
        const int r = it / D8, k8 = it - r * D8;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __ldg(bd + k8 * 8 + e);
        const float* hr = hs + r * 24;
#pragma unroll
        for (int j = 0; j < LSTM_UNITS; ++j) {
            const float hj = hr[j];
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)j * D + k8 * 8));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)j * D + k8 * 8) + 1);
            o[0] = fmaf(hj, w0.x, o[0]); o[1] = fmaf(hj, w0.y, o[1]); o[2] = fmaf(hj, w0.z, o[2]); o[3] = fmaf(hj, w0.w, o[3]);
            o[4] = fmaf(hj, w1.x, o[4]); o[5] = fmaf(hj, w1.y, o[5]); o[6] = fmaf(hj, w1.z, o[6]); o[7] = fmaf(hj, w1.w, o[7]);
        }
        if (SH) {
            const int f = k8 / C8, c8 = k8 - f * C8;
            sh16_store8(reinterpret_cast<uint8_t*>(yv) + (size_t)(s0 + r) * D * 4, Fb, C, f, c8, o);
        } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(yv) + (size_t)(s0 + r) * D + k8 * 8);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

}  // namespace nunet
