// The non-convolution units of the offline path over "sh16" activations (conv_tc3.cuh): a tensor [frame][F][C]
// keeps every frame row planar, row[hi|lo][chunk C/8][f][8 halves], x ~= hi + lo.  Same arithmetic as their fp32
// twins in misc_kernels.cuh (which the streaming plan keeps using); only the loads join hi + lo and the stores split.
// Thread mappings put consecutive bins f on consecutive lanes, so every 16-byte access of a warp is contiguous.
#pragma once
#include "conv_tc3.cuh"
#include "misc_kernels.cuh"

namespace nunet {

// 8 channels (chunk c8) of bin f of a frame row `row` (pointer to the row's first byte)
__device__ __forceinline__ void sh16_load8(const uint8_t* row, int F, int C, int f, int c8, float* v) {
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(row + ((size_t)c8 * F + f) * 16));
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(row + ((size_t)((C >> 3) + c8) * F + f) * 16));
    join8(hi, lo, v);
}
__device__ __forceinline__ void sh16_store8(uint8_t* row, int F, int C, int f, int c8, const float* v) {
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(row + ((size_t)c8 * F + f) * 16) = hi;
    *reinterpret_cast<uint4*>(row + ((size_t)((C >> 3) + c8) * F + f) * 16) = lo;
}

// input_layer = inconv(64) on the 1-channel magnitude (models/proposed.py:293): one thread per pixel, all 64
// channels in registers (LayerNorm is thread-local), 16 coalesced 16-byte stores.
__global__ void __launch_bounds__(128) input_layer_sh_kernel(const float* __restrict__ mag, const float* __restrict__ w,
                                                            const float* __restrict__ b, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ alpha,
                                                            uint8_t* __restrict__ out, long long npix, int F) {
    __shared__ __align__(16) float par[4 * 64];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        par[i] = __ldg(w + i);
        par[64 + i] = __ldg(b + i);
        par[128 + i] = __ldg(gamma + i);
        par[192 + i] = __ldg(beta + i);
    }
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const float x = __ldg(mag + pix);
    float2 v2[32];
    float* v = reinterpret_cast<float*>(v2);
#pragma unroll
    for (int c = 0; c < 64; ++c) v[c] = fmaf(x, par[c], par[64 + c]);
    ln_prelu_s<64>(v2, par + 128, par + 192, __ldg(alpha));
    const long long frame = pix / F;
    const int f = (int)(pix - frame * F);
    uint8_t* row = out + frame * ((long long)F * 256);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) sh16_store8(row, F, 64, f, c8, v + 8 * c8);
}

// out_conv: Conv2D 1x1, 64 -> 1 (models/proposed.py:615).  One thread per pixel.
__global__ void __launch_bounds__(128) out_conv_sh_kernel(const uint8_t* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ out, long long npix,
                                                         int F, int out_stride, int out_off) {
    __shared__ float ws[64];
    if (threadIdx.x < 64) ws[threadIdx.x] = __ldg(w + threadIdx.x);
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long frame = pix / F;
    const int f = (int)(pix - frame * F);
    const uint8_t* row = x + frame * ((long long)F * 256);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        float v[8];
        sh16_load8(row, F, 64, f, c8, v);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            s0 = fmaf(v[e], ws[c8 * 8 + e], s0);
            s1 = fmaf(v[e + 1], ws[c8 * 8 + e + 1], s1);
        }
    }
    out[frame * out_stride + out_off + f] = (s0 + s1) + __ldg(b);
}

// The two-layer MLP of a CTFA attention (64 -> 16 relu -> 64 sigmoid) for CTFA_FPB frames at once, weights staged in
// shared memory by the caller: thread (fr = tid / 64, c = tid % 64).  v_s [CTFA_FPB][64] input, h_s [CTFA_FPB][16] scratch.
constexpr int CTFA_FPB = 4;   // frames per CTA (256 threads)
struct MlpSmem {
    float k0[64 * 16], b0[16], k1[16 * 64], b1[64];
};
__device__ __forceinline__ void mlp_stage(MlpSmem& m, const MlpW& w) {
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
        m.k0[i] = __ldg(w.k0 + i);
        m.k1[i] = __ldg(w.k1 + i);
    }
    if (threadIdx.x < 16) m.b0[threadIdx.x] = __ldg(w.b0 + threadIdx.x);
    if (threadIdx.x < 64) m.b1[threadIdx.x] = __ldg(w.b1 + threadIdx.x);
}
// call with all 256 threads after a __syncthreads() that made v_s and the staged weights visible
__device__ __forceinline__ float mlp_apply(const MlpSmem& m, const float (*v_s)[64], float (*h_s)[16]) {
    const int fr = threadIdx.x >> 6, c = threadIdx.x & 63;
    if (c < 16) {
        float a = m.b0[c];
#pragma unroll 16
        for (int k = 0; k < 64; ++k) a = fmaf(v_s[fr][k], m.k0[k * 16 + c], a);
        h_s[fr][c] = fmaxf(a, 0.0f);
    }
    __syncthreads();
    float o = m.b1[c];
#pragma unroll
    for (int j = 0; j < 16; ++j) o = fmaf(h_s[fr][j], m.k1[j * 64 + c], o);
    return sigmoidf_(o);
}

// CTFA stage 1 (models/proposed.py:125): TA[frame, c] = MLP(mean_f x[frame, f, c]).  One CTA per CTFA_FPB frames:
// warp w sums chunk w (8 channels) over the bins of each frame, lanes striding f; then all 256 threads run the MLP.
__global__ void __launch_bounds__(256) ctfa_ta_sh_kernel(const uint8_t* __restrict__ x, MlpW ta, float* __restrict__ ta_out, int F,
                                                        long long frames) {
    __shared__ MlpSmem w_s;
    __shared__ float mean_s[CTFA_FPB][64];
    __shared__ float h_s[CTFA_FPB][16];
    mlp_stage(w_s, ta);
    const long long frame0 = (long long)blockIdx.x * CTFA_FPB;
    const int c8 = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int fr = 0; fr < CTFA_FPB; ++fr) {
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (frame0 + fr < frames) {
            const uint8_t* row = x + (size_t)(frame0 + fr) * F * 256;
#pragma unroll 4
            for (int f = lane; f < F; f += 32) {
                float v[8];
                sh16_load8(row, F, 64, f, c8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] += v[e];
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) s[e] += __shfl_xor_sync(0xffffffffu, s[e], m);
        }
        if (lane == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) mean_s[fr][c8 * 8 + e] = ((F & (F - 1)) == 0) ? s[e] * (1.0f / (float)F) : s[e] / (float)F;
        }
    }
    __syncthreads();
    const float t = mlp_apply(w_s, mean_s, h_s);
    const long long frame = frame0 + (threadIdx.x >> 6);
    if (frame < frames) ta_out[frame * 64 + (threadIdx.x & 63)] = t;
}

// The CTFA MLP (64 -> 16 relu -> 64, pre-sigmoid) of FOUR frames by one warp, as packed fp32x2 FMAs over frame pairs: layer 1 on
// lane = (hidden unit j, frame pair p), layer 2 on lane = channels 2 lane, 2 lane + 1 of all four frames.  v_s [64][4] holds the
// inputs (channel-major), h_s [16][4] is scratch; every sum runs in the order of mlp_apply / ctfa_mlp (k and j ascending from the
// bias), so the results equal the one-frame kernels' bit for bit.  ox / oy [4]: outputs of the two channels per frame.
__device__ __forceinline__ void mlp4_warp(const MlpSmem& m, const float (*v_s)[4], float (*h_s)[4], int lane, float* ox, float* oy) {
    __syncwarp();
    {
        const int j = lane & 15, p = lane >> 4;
        float2 acc = make_float2(m.b0[j], m.b0[j]);
#pragma unroll 16
        for (int k = 0; k < 64; ++k) {
            const float2 av = *reinterpret_cast<const float2*>(&v_s[k][2 * p]);
            const float w = m.k0[k * 16 + j];
            acc = __ffma2_rn(av, make_float2(w, w), acc);
        }
        *reinterpret_cast<float2*>(&h_s[j][2 * p]) = make_float2(fmaxf(acc.x, 0.0f), fmaxf(acc.y, 0.0f));
    }
    __syncwarp();
    const float2 b1 = *reinterpret_cast<const float2*>(&m.b1[2 * lane]);
    float2 x01 = make_float2(b1.x, b1.x), x23 = x01, y01 = make_float2(b1.y, b1.y), y23 = y01;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 hv = *reinterpret_cast<const float4*>(&h_s[j][0]);
        const float2 w2 = *reinterpret_cast<const float2*>(&m.k1[j * 64 + 2 * lane]);
        x01 = __ffma2_rn(make_float2(hv.x, hv.y), make_float2(w2.x, w2.x), x01);
        x23 = __ffma2_rn(make_float2(hv.z, hv.w), make_float2(w2.x, w2.x), x23);
        y01 = __ffma2_rn(make_float2(hv.x, hv.y), make_float2(w2.y, w2.y), y01);
        y23 = __ffma2_rn(make_float2(hv.z, hv.w), make_float2(w2.y, w2.y), y23);
    }
    ox[0] = x01.x; ox[1] = x01.y; ox[2] = x23.x; ox[3] = x23.y;
    oy[0] = y01.x; oy[1] = y01.y; oy[2] = y23.x; oy[3] = y23.y;
}

// CTFA stage 1 for blocks with few bins (F <= 64): one WARP per group of four frames, grid-stride.  Lane (c4 = lane >> 3,
// fq = lane & 7) sums chunks c4 and c4 + 4 over the bins f = fq, fq + 8, ..: one load instruction of the warp reads eight
// neighbouring positions (128 contiguous bytes) of four planes; the four frames of a group go one after the other, then their
// MLPs run together (mlp4_warp: a per-frame MLP keeps 16 lanes busy and was a third of this kernel's instructions).  The
// CTA-per-four-frames kernel above spends ~10 us of fixed latency per CTA, which is all there is at F <= 32; this one does not.
__global__ void __launch_bounds__(256) ctfa_ta_warp_sh_kernel(const uint8_t* __restrict__ x, MlpW ta, float* __restrict__ ta_out, int F,
                                                             long long frames) {
    __shared__ MlpSmem w_s;
    __shared__ __align__(16) float mean_s[8][64][4];
    __shared__ __align__(16) float h_s[8][16][4];
    mlp_stage(w_s, ta);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c4 = lane >> 3, fq = lane & 7;
    // F is a power of two in every block of the topology: multiplying by the (exact) reciprocal is the same rounding as the
    // division of the reference's mean and a tenth of its instructions
    const bool pow2F = (F & (F - 1)) == 0;
    const float invF = 1.0f / (float)F;
    const long long groups = (frames + 3) >> 2;
    for (long long g = (long long)blockIdx.x * 8 + warp; g < groups; g += (long long)gridDim.x * 8) {
        const long long frame0 = g * 4;
        const int nf = (int)min((long long)4, frames - frame0);
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            float s[2][8];
#pragma unroll
            for (int e = 0; e < 8; ++e) s[0][e] = s[1][e] = 0.0f;
            if (i < nf) {
                const uint8_t* row = x + (size_t)(frame0 + i) * F * 256;
#pragma unroll 2
                for (int f = fq; f < F; f += 8) {
                    float v0[8], v1[8];
                    sh16_load8(row, F, 64, f, c4, v0);
                    sh16_load8(row, F, 64, f, c4 + 4, v1);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        s[0][e] += v0[e];
                        s[1][e] += v1[e];
                    }
                }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s[hh][e] += __shfl_xor_sync(0xffffffffu, s[hh][e], 1);
                    s[hh][e] += __shfl_xor_sync(0xffffffffu, s[hh][e], 2);
                    s[hh][e] += __shfl_xor_sync(0xffffffffu, s[hh][e], 4);
                }
            if (fq == 0) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        mean_s[warp][(c4 + 4 * hh) * 8 + e][i] = pow2F ? s[hh][e] * invF : s[hh][e] / (float)F;
            }
        }
        float ox[4], oy[4];
        mlp4_warp(w_s, mean_s[warp], h_s[warp], lane, ox, oy);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nf) *reinterpret_cast<float2*>(ta_out + (frame0 + i) * 64 + 2 * lane) = make_float2(sigmoidf_(ox[i]), sigmoidf_(oy[i]));
        __syncwarp();
    }
}

// CTFA stage 2 (see ctfa_gate_kernel) with one WARP per frame and the MLP weights staged once per CTA: lane l owns channels
// l and l + 32.  Same summation order (oldest row first) and MLP operation order as ctfa_gate_kernel, so the results are
// bit-identical; no ring (offline plans and the reference's one-frame rule only).
// Time-chunked offline calls: `hist` [clip][31][64] holds the TA rows of the 31 frames in front of this chunk (oldest first) and
// t0 is the clip-relative index of the chunk's first frame; hist == nullptr means the chunk starts the clip (zeros before it).
__global__ void __launch_bounds__(256) ctfa_gate_warp_kernel(const float* __restrict__ ta, MlpW fa, float* __restrict__ gate, int T,
                                                            int mode_div32, long long frames, const float* __restrict__ hist, int t0) {
    __shared__ MlpSmem w_s;
    __shared__ float avg_s[8][64];
    __shared__ float h_s[8][16];
    mlp_stage(w_s, fa);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long frame = (long long)blockIdx.x * 8 + warp; frame < frames; frame += (long long)gridDim.x * 8) {
        const float tv0 = ta[frame * 64 + lane], tv1 = ta[frame * 64 + 32 + lane];
        float a0, a1;
        if (mode_div32) {
            a0 = tv0 * (1.0f / CTFA_WINDOW);
            a1 = tv1 * (1.0f / CTFA_WINDOW);
        } else {
            const int t = (int)(frame % T);
            const float* hrow = hist ? hist + ((frame / T) * (CTFA_WINDOW - 1) + (CTFA_WINDOW - 1)) * 64 : nullptr;   // row of frame "t = 0"
            const int n = min((hist ? t0 : 0) + t + 1, CTFA_WINDOW);
            // oldest row first: the order of the reference's pooling window (and of the streaming ring)
            float s0 = 0.0f, s1 = 0.0f;
            for (int d = n - 1; d >= 0; --d) {
                const float* src = (d <= t) ? ta + (frame - d) * 64 : hrow + (long long)(t - d) * 64;
                s0 += src[lane];
                s1 += src[32 + lane];
            }
            a0 = s0 * (1.0f / CTFA_WINDOW);
            a1 = s1 * (1.0f / CTFA_WINDOW);
        }
        avg_s[warp][lane] = a0;
        avg_s[warp][32 + lane] = a1;
        __syncwarp();
        if (lane < 16) {
            float a = w_s.b0[lane];
#pragma unroll 8
            for (int k = 0; k < 64; ++k) a = fmaf(avg_s[warp][k], w_s.k0[k * 16 + lane], a);
            h_s[warp][lane] = fmaxf(a, 0.0f);
        }
        __syncwarp();
        float o0 = w_s.b1[lane], o1 = w_s.b1[32 + lane];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            o0 = fmaf(h_s[warp][j], w_s.k1[j * 64 + lane], o0);
            o1 = fmaf(h_s[warp][j], w_s.k1[j * 64 + 32 + lane], o1);
        }
        gate[frame * 64 + lane] = sigmoidf_(o0) * tv0;
        gate[frame * 64 + 32 + lane] = sigmoidf_(o1) * tv1;
        __syncwarp();
    }
}

// The same stage, FOUR consecutive frames per warp iteration (the default).  ctfa_gate_warp_kernel is instruction-bound
// (ncu: ~860 warp instructions per frame, two thirds of them address arithmetic of the window loop, and an MLP that keeps 16
// lanes busy); here lane l owns channels 2l, 2l+1 (one 8-byte load per TA row), the 35 rows that the four windows share are
// read once and added into each window oldest row first -- the order of ctfa_gate_kernel, so results stay bit-identical --
// and the MLP runs as packed fp32x2 FMAs over frame pairs: layer 1 on (hidden unit j, frame pair) = all 32 lanes, layer 2 on
// (2 channels x 4 frames) per lane.  Groups that straddle two clips (and the last, partial one) take the per-frame sums below.
__device__ __forceinline__ float2 ctfa_window_sum2(const float* __restrict__ ta, const float* __restrict__ hist, long long frame, int T,
                                                   int t0, int lane) {
    const int t = (int)(frame % T);
    const float* hrow = hist ? hist + ((frame / T) * (CTFA_WINDOW - 1) + (CTFA_WINDOW - 1)) * 64 : nullptr;   // row of frame "t = 0"
    const int n = min((hist ? t0 : 0) + t + 1, CTFA_WINDOW);
    // all 32 candidate rows, eight loads in flight at a time; rows in front of the window (d >= n, the OLDEST ones) count as +0,
    // which leaves every partial sum what the n-row loop of ctfa_gate_warp_kernel makes it
    float2 s = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int db = CTFA_WINDOW - 1; db >= 0; db -= 8) {
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int d = db - u;
            const float* src = (d <= t) ? ta + (frame - d) * 64 : hrow + (long long)(t - d) * 64;
            v[u] = (d < n) ? *reinterpret_cast<const float2*>(src + 2 * lane) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __fadd2_rn(s, v[u]);
    }
    return s;
}
__global__ void __launch_bounds__(256) ctfa_gate_warp4_kernel(const float* __restrict__ ta, MlpW fa, float* __restrict__ gate, int T,
                                                             int mode_div32, long long frames, const float* __restrict__ hist, int t0) {
    __shared__ MlpSmem w_s;
    __shared__ __align__(16) float avg_s[8][64][4];   // [warp][channel][frame of the group]
    __shared__ __align__(16) float h_s[8][16][4];
    mlp_stage(w_s, fa);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long groups = (frames + 3) >> 2;
    for (long long g = (long long)blockIdx.x * 8 + warp; g < groups; g += (long long)gridDim.x * 8) {
        const long long frame0 = g * 4;
        const int nf = (int)min((long long)4, frames - frame0);
        float2 tv[4], a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            tv[i] = (i < nf) ? *reinterpret_cast<const float2*>(ta + (frame0 + i) * 64 + 2 * lane) : make_float2(0.0f, 0.0f);
        const int tf = (int)(frame0 % T);
        if (mode_div32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = make_float2(tv[i].x * (1.0f / CTFA_WINDOW), tv[i].y * (1.0f / CTFA_WINDOW));
        } else if (nf == 4 && tf + 3 < T) {
            // four frames of one clip: their windows share the 35 rows t = tf - 31 .. tf + 3.  Rows in front of the clip come from
            // the carried history (time-chunked calls) or, in front of what exists (t < -t0), count as +0 -- they are the OLDEST
            // terms of a window, so every partial sum is what the n-row loop makes it.
            const int base = hist ? t0 : 0;
            const float* hrow = hist ? hist + ((frame0 / T) * (CTFA_WINDOW - 1) + (CTFA_WINDOW - 1)) * 64 : nullptr;   // row of "t = 0"
            const float* trow = ta + (frame0 - tf) * 64;                                                             // row of t = 0
            float2 s[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int rb = 0; rb < CTFA_WINDOW + 3; rb += 7) {
                float2 v[7];
#pragma unroll
                for (int u = 0; u < 7; ++u) {
                    const int tr = tf - (CTFA_WINDOW - 1) + rb + u;
                    const float* src = (tr >= 0 ? trow : hrow) + (long long)tr * 64;
                    v[u] = (tr >= -base) ? *reinterpret_cast<const float2*>(src + 2 * lane) : make_float2(0.0f, 0.0f);
                }
#pragma unroll
                for (int u = 0; u < 7; ++u)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (rb + u >= i && rb + u <= i + CTFA_WINDOW - 1) s[i] = __fadd2_rn(s[i], v[u]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = make_float2(s[i].x * (1.0f / CTFA_WINDOW), s[i].y * (1.0f / CTFA_WINDOW));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 s = make_float2(0.0f, 0.0f);
                if (i < nf) s = ctfa_window_sum2(ta, hist, frame0 + i, T, t0, lane);
                a[i] = make_float2(s.x * (1.0f / CTFA_WINDOW), s.y * (1.0f / CTFA_WINDOW));
            }
        }
        *reinterpret_cast<float4*>(&avg_s[warp][2 * lane][0]) = make_float4(a[0].x, a[1].x, a[2].x, a[3].x);
        *reinterpret_cast<float4*>(&avg_s[warp][2 * lane + 1][0]) = make_float4(a[0].y, a[1].y, a[2].y, a[3].y);
        __syncwarp();
        float ox[4], oy[4];
        mlp4_warp(w_s, avg_s[warp], h_s[warp], lane, ox, oy);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nf)
                *reinterpret_cast<float2*>(gate + (frame0 + i) * 64 + 2 * lane) = make_float2(sigmoidf_(ox[i]) * tv[i].x, sigmoidf_(oy[i]) * tv[i].y);
        __syncwarp();
    }
}

// storage position of bin f inside a plane: natural order, or [even bins | odd bins]
__device__ __forceinline__ int sh16_pos(int f, int F, int eo) { return eo ? (f & 1) * (F >> 1) + (f >> 1) : f; }

// CTFA stage 3 + residual (models/proposed.py:319): out = x * gate[frame] + res.  Item = (frame, chunk, output
// position): 8 channels.  x and res are stored [even | odd] (in_eo), the output as the consumer wants it (out_eo).
__global__ void __launch_bounds__(256) gate_residual_sh_kernel(const uint8_t* __restrict__ x, const uint8_t* __restrict__ res,
                                                              const float* __restrict__ gate /*[frames][64]*/,
                                                              uint8_t* __restrict__ out, long long n8, int F, int in_eo, int out_eo) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    // F is a power of two in every block of the topology: a shift instead of a 64-bit division per item (the kernel runs at the
    // HBM rate either way, but the division was two thirds of its instructions -- and the step is power-capped)
    const bool pow2F = (F & (F - 1)) == 0;
    const int lgF = 31 - __clz(F);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const long long fc = pow2F ? (i >> lgF) : i / F;      // frame * 8 + chunk
        const int po = (int)(i - fc * F);           // output storage position
        const int f = out_eo ? ((po < (F >> 1)) ? 2 * po : 2 * (po - (F >> 1)) + 1) : po;   // its bin
        const int pi = sh16_pos(f, F, in_eo);
        const long long frame = fc >> 3;
        const int c8 = (int)(fc & 7);
        const long long rowoff = frame * ((long long)F * 256);
        float xv[8], rv[8];
        sh16_load8(x + rowoff, F, 64, pi, c8, xv);
        sh16_load8(res + rowoff, F, 64, pi, c8, rv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + frame * 64 + c8 * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + frame * 64 + c8 * 8) + 1);
        float o[8];
        o[0] = fmaf(xv[0], g0.x, rv[0]); o[1] = fmaf(xv[1], g0.y, rv[1]);
        o[2] = fmaf(xv[2], g0.z, rv[2]); o[3] = fmaf(xv[3], g0.w, rv[3]);
        o[4] = fmaf(xv[4], g1.x, rv[4]); o[5] = fmaf(xv[5], g1.y, rv[5]);
        o[6] = fmaf(xv[6], g1.z, rv[6]); o[7] = fmaf(xv[7], g1.w, rv[7]);
        sh16_store8(out + rowoff, F, 64, po, c8, o);
    }
}

// Streaming step: the whole CTFA of one block -- frequency mean, TA MLP, FA MLP on TA / 32 (the one-frame graph's rule,
// models/proposed.py:162-196), gate and `x * gate + residual` (:319) -- for one stream per CTA (256 threads), one launch instead of
// three (ctfa_ta*, ctfa_gate*, gate_residual_sh).  Operation order of every sum as in those kernels (ctfa_ta_sh_kernel's reduction,
// ctfa_mlp, gate_residual_sh_kernel's fused multiply-add).
__global__ void __launch_bounds__(256) ctfa_stream_sh_kernel(const uint8_t* __restrict__ x, const uint8_t* __restrict__ res, MlpW ta, MlpW fa,
                                                            uint8_t* __restrict__ out, int F, int in_eo, int out_eo) {
    __shared__ float v_s[64], h_s[16], t_s[64], g_s[64];
    const long long s = blockIdx.x;
    const int tid = threadIdx.x, c8 = tid >> 5, lane = tid & 31;
    const uint8_t* xrow = x + s * (long long)F * 256;
    float sum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int f = lane; f < F; f += 32) {           // storage positions: the mean does not care about the bin order
        float v[8];
        sh16_load8(xrow, F, 64, f, c8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum[e] += v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) sum[e] += __shfl_xor_sync(0xffffffffu, sum[e], m);
    }
    if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v_s[c8 * 8 + e] = ((F & (F - 1)) == 0) ? sum[e] * (1.0f / (float)F) : sum[e] / (float)F;
    }
    __syncthreads();
    if (tid < 64) {
        const float t = ctfa_mlp(v_s, h_s, ta, tid, 1);
        t_s[tid] = t;
        asm volatile("bar.sync 1, 64;" ::: "memory");      // everyone is done reading v_s / h_s
        v_s[tid] = t * (1.0f / CTFA_WINDOW);                // 31 zero rows + this frame, average-pooled over 32
        asm volatile("bar.sync 1, 64;" ::: "memory");
        const float g = ctfa_mlp(v_s, h_s, fa, tid, 1);
        g_s[tid] = g * t;
    }
    __syncthreads();
    const uint8_t* rrow = res + s * (long long)F * 256;
    uint8_t* orow = out + s * (long long)F * 256;
    for (int i = tid; i < F * 8; i += 256) {
        const int c = i / F, po = i - c * F;                 // chunk, output storage position
        const int f = out_eo ? ((po < (F >> 1)) ? 2 * po : 2 * (po - (F >> 1)) + 1) : po;
        const int pi = sh16_pos(f, F, in_eo);
        float xv[8], rv[8], o[8];
        sh16_load8(xrow, F, 64, pi, c, xv);
        sh16_load8(rrow, F, 64, pi, c, rv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(xv[e], g_s[c * 8 + e], rv[e]);
        sh16_store8(orow, F, 64, po, c, o);
    }
}

// Last decoder block: out_conv (Conv2D 1x1, 64 -> 1, models/proposed.py:615) applied to x * gate + res without ever
// writing the 64-channel block output.  One thread per bin; lanes = consecutive bins, so every 16-byte load of a warp
// is contiguous.  x and res share the [even | odd] order (in_eo); the estimate goes to bin order.
__global__ void __launch_bounds__(128) gate_residual_out_conv_sh_kernel(const uint8_t* __restrict__ x, const uint8_t* __restrict__ res,
                                                                       const float* __restrict__ gate, const float* __restrict__ w,
                                                                       const float* __restrict__ b, float* __restrict__ out,
                                                                       long long npix, int F, int in_eo, int out_stride, int out_off) {
    __shared__ float ws[64];
    if (threadIdx.x < 64) ws[threadIdx.x] = __ldg(w + threadIdx.x);
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long frame = pix / F;
    const int po = (int)(pix - frame * F);                       // storage position (coalesced loads)
    const int f = in_eo ? ((po < (F >> 1)) ? 2 * po : 2 * (po - (F >> 1)) + 1) : po;
    const long long rowoff = frame * ((long long)F * 256);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        float xv[8], rv[8];
        sh16_load8(x + rowoff, F, 64, po, c8, xv);
        sh16_load8(res + rowoff, F, 64, po, c8, rv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + frame * 64 + c8 * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + frame * 64 + c8 * 8) + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            // the 64-channel value is rounded through the sh16 split exactly like the unfused path stores it
            float y0 = fmaf(xv[e], gg[e], rv[e]), y1 = fmaf(xv[e + 1], gg[e + 1], rv[e + 1]);
            const __half2 hh = __floats2half2_rn(y0, y1);
            const float2 hf = __half22float2(hh);
            const float2 lf = __half22float2(__floats2half2_rn(y0 - hf.x, y1 - hf.y));
            s0 = fmaf(hf.x + lf.x, ws[c8 * 8 + e], s0);
            s1 = fmaf(hf.y + lf.y, ws[c8 * 8 + e + 1], s1);
        }
    }
    out[frame * out_stride + out_off + f] = (s0 + s1) + __ldg(b);
}

// half index of element (f, c) of part `part` inside a planar frame row [F][C]
__device__ __forceinline__ int sh16_half_index(int F, int C, int part, int f, int c) {
    return (((part * (C >> 3) + (c >> 3)) * F + f) << 3) + (c & 7);
}

// ---- carried state of time-chunked offline calls -------------------------------------------------------------------
// After a chunk of T frames: hist[clip][j] (j < 31) <- row T + j of the concatenation [old hist (31 rows) | ta (T rows)],
// i.e. the TA rows of the 31 frames in front of the next chunk.  One CTA (64 threads) per clip; all reads before any write.
__global__ void __launch_bounds__(64) ctfa_hist_update_kernel(const float* __restrict__ ta, float* __restrict__ hist, int T, int have_hist) {
    const long long b = blockIdx.x;
    const int c = threadIdx.x;
    float v[CTFA_WINDOW - 1];
    float* h = hist + b * (CTFA_WINDOW - 1) * 64;
#pragma unroll
    for (int j = 0; j < CTFA_WINDOW - 1; ++j) {
        const int i = T + j;                                   // index into the concatenation
        v[j] = (i < CTFA_WINDOW - 1) ? (have_hist ? h[i * 64 + c] : 0.0f) : ta[(b * T + (i - (CTFA_WINDOW - 1))) * 64 + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CTFA_WINDOW - 1; ++j) h[j * 64 + c] = v[j];
}

// The last frame row of every tensor that feeds a causal (two time tap) conv of one nested sub-U-Net is kept for the next
// chunk, where it is the row "t = -1" (conv_tc3: prev0 / prev1).  grid = (entries, clips).
constexpr int CARRY_MAX = 32;
struct CarrySave {
    const uint8_t* src[CARRY_MAX];   // tensor base of this chunk: [clip][T][row_bytes]
    uint8_t* dst[CARRY_MAX];         // carry base: [clip][row_bytes]
    int row16[CARRY_MAX];            // row_bytes / 16
    int n;
};
__global__ void __launch_bounds__(128) carry_save_kernel(const __grid_constant__ CarrySave cs, int T) {
    const int e = blockIdx.x;
    const long long b = blockIdx.y;
    const int n16 = cs.row16[e];
    const uint4* s = reinterpret_cast<const uint4*>(cs.src[e]) + (b * T + (T - 1)) * (long long)n16;
    uint4* d = reinterpret_cast<uint4*>(cs.dst[e]) + b * (long long)n16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d[i] = __ldg(s + i);
}

}  // namespace nunet
