// Fused causal convolution unit, FP32 SIMT path (exact fp32 products, fp32 accumulation).
//
// One kernel covers every convolutional layer factory of the reference (models/proposed.py):
//   conv / conv_valid        :198/:208  ZeroPad((1,0),(1,1)) + Conv2D(2,3) stride (1,2) + LN + PReLU
//   spconv / spconv_valid    :227/:240  same conv stride 1 to 2C channels + sub-pixel shuffle + LN + PReLU
//   inconv                   :218       Conv2D 1x1 + LN + PReLU
//   down_sampling            :253       Conv2D (1,3) stride (1,2) 'same', bias only
//   up_sampling -> inconv    :260,:218  Conv2DTranspose (1,3) stride (1,2) composed on the host with the
//                                       1x1 conv that always follows it (no non-linearity in between)
// as an implicit GEMM: M = output pixels (b, t, f), N = all conv output channels (needed so that the
// LayerNorm over channels stays inside the CTA), K = taps x input channels.
//
// Data layout: activations are NHWC [frame = b*T + t][F][C] fp32.  The channel concatenations of the
// reference graph (`Concatenate(axis=3)`) are never materialised: a unit reads up to two source tensors.
// The time tap kt=0 is the previous frame: inside a clip it is the same tensor one frame earlier; at the
// first frame it is zero (offline, ZeroPadding2D) or the carried history row (streaming,
// converter_proposed.py:226 `Concatenate(axis=1)([prev, cur])`).
//
// Tiling: a CTA owns P = 128 output pixels = G clips x TT frames x FT bins and all COUT channels.
// Per source tensor it stages the (TT+KT-1) x (stride*(FT-1)+KF) x C input patch in shared memory once
// (cp.async, zero-filled halo) and streams the weights through a double-buffered 32-row shared tile.
// Thread (tx, py) accumulates PM pixels x CN contiguous channels in registers; LayerNorm is a two-pass
// (mean, centred variance) reduction over the TX lanes that share a pixel, by warp shuffle.
#pragma once
#include "common.cuh"

namespace nunet {

enum Epi { EPI_LN = 0, EPI_BIAS = 1, EPI_SHUF32 = 2, EPI_SHUF64 = 3 };

struct ConvParams {
    const float* a_cur;   // source A [frames][F_in][CA]
    const float* a_prev;  // history row of A per clip/stream [B][F_in][CA] (has_prev only)
    const float* b_cur;   // source B (CB == 0: absent)
    const float* b_prev;
    const float* w;       // [KT*KF][CA+CB][COUT], columns permuted for the thread mapping (pack_conv_w)
    const float* bias;    // [COUT] in logical conv-channel order
    const float* gamma;   // LayerNorm scale / offset indexed by OUTPUT channel (after the shuffle)
    const float* beta;
    const float* alpha;   // PReLU slope (one scalar per layer, shared_axes=[1,2,3])
    float* out;           // [frames][F_out * (shuffle ? 2 : 1)][C_out]
    int CA, CB;
    int B;                // clips (offline) or streams (streaming)
    int T;                // frames per clip in this call (streaming: 1)
    int has_prev;         // 1: the row at t = -1 is *_prev[b]; 0: zeros
    int F_in, F_out;      // input bins; conv output bins (before the sub-pixel shuffle)
    int KT, KF, padl, stride;
    int lFT, lTT, lG;     // log2 of the tile geometry (bins, frames, clips)
};

constexpr int CONV_KC = 32;   // weight rows per shared-memory stage
constexpr int CONV_P = 128;   // output pixels per CTA

__host__ __device__ inline int conv_tile_floats(int G, int TT, int FT, int KT, int KF, int stride, int Cmax) {
    return G * (TT + KT - 1) * (stride * (FT - 1) + KF) * (Cmax + 4);
}

template <int COUT, int CN, int PM, int NT, int EPI>
__global__ void __launch_bounds__(NT) conv_unit_kernel(const ConvParams p) {
    constexpr int TX = COUT / CN;   // lanes across channels
    constexpr int PG = NT / TX;     // pixel groups
    static_assert(PG * PM == CONV_P, "tile must hold 128 pixels");
    static_assert(CN % 4 == 0 && TX <= 32 && (32 % TX) == 0, "bad thread mapping");
    constexpr int KC = CONV_KC;

    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                    // [2][KC][COUT]
    float* tile = smem + 2 * KC * COUT;  // [G*R][FW][C+4]

    const int tid = threadIdx.x;
    const int tx = tid % TX;
    const int py = tid / TX;
    const int FT = 1 << p.lFT, TT = 1 << p.lTT, G = 1 << p.lG;
    const int R = TT + p.KT - 1;
    const int FW = p.stride * (FT - 1) + p.KF;
    const int ntaps = p.KT * p.KF;
    const int Cin = p.CA + p.CB;

    const int nfb = p.F_out >> p.lFT;
    const int ntb = (p.T + TT - 1) >> p.lTT;
    int bid = blockIdx.x;
    const int fblk = bid % nfb;
    bid /= nfb;
    const int tblk = bid % ntb;
    const int bgrp = bid / ntb;
    const int f0 = fblk << p.lFT, t0 = tblk << p.lTT, b0 = bgrp << p.lG;

    // this thread's pixels: pix = py + PG*i  ->  (g, tt, fl)
    int rowpix[PM];   // (g*R + tt)*FW + stride*fl, in input pixels
    bool pvalid[PM];
#pragma unroll
    for (int i = 0; i < PM; ++i) {
        const int pix = py + PG * i;
        const int fl = pix & (FT - 1);
        const int tt = (pix >> p.lFT) & (TT - 1);
        const int g = pix >> (p.lFT + p.lTT);
        pvalid[i] = (g < G) && (b0 + g < p.B) && (t0 + tt < p.T);
        rowpix[i] = (g < G) ? ((g * R + tt) * FW + p.stride * fl) : 0;
    }

    float acc[PM][CN];
#pragma unroll
    for (int i = 0; i < PM; ++i)
#pragma unroll
        for (int e = 0; e < CN; ++e) acc[i][e] = 0.0f;

    const int nchA = ntaps * (p.CA / KC);
    const int nch = nchA + ntaps * (p.CB / KC);

    auto load_w = [&](int buf, int ci) {
        int row;
        if (ci < nchA) {
            const int per = p.CA / KC;
            row = (ci / per) * Cin + (ci % per) * KC;
        } else {
            const int cj = ci - nchA;
            const int per = p.CB / KC;
            row = (cj / per) * Cin + p.CA + (cj % per) * KC;
        }
        const float* src = p.w + (size_t)row * COUT;
        float* dst = Ws + buf * (KC * COUT);
        for (int i = tid; i < KC * COUT / 4; i += NT) cp_async16(dst + i * 4, src + i * 4, 16);
    };

    auto load_tile = [&](const float* cur, const float* prev, int C) {
        const int pitch = C + 4;
        const int C4 = C >> 2;
        const size_t frame_elems = (size_t)p.F_in * C;
        for (int row = 0; row < G * R; ++row) {
            const int g = row / R, r = row - g * R;
            const int b = b0 + g;
            const int t = t0 + r - (p.KT - 1);
            const float* base = nullptr;
            if (b < p.B && t < p.T) {
                if (t >= 0) base = cur + ((size_t)b * p.T + t) * frame_elems;
                else if (p.has_prev) base = prev + (size_t)b * frame_elems;
            }
            float* drow = tile + (size_t)row * FW * pitch;
            for (int i = tid; i < FW * C4; i += NT) {
                const int px = i / C4, c4 = i - px * C4;
                const int fi = p.stride * f0 - p.padl + px;
                const bool ok = (base != nullptr) && (fi >= 0) && (fi < p.F_in);
                const float* src = ok ? (base + (size_t)fi * C + c4 * 4) : cur;
                cp_async16(drow + px * pitch + c4 * 4, src, ok ? 16 : 0);
            }
        }
    };

    load_w(0, 0);
    cp_async_commit();

    int ci = 0;
    const int nsrc = p.CB > 0 ? 2 : 1;
    for (int s = 0; s < nsrc; ++s) {
        const int C = s ? p.CB : p.CA;
        const int pitch = C + 4;
        load_tile(s ? p.b_cur : p.a_cur, s ? p.b_prev : p.a_prev, C);
        cp_async_commit();
        int poff[PM];
#pragma unroll
        for (int i = 0; i < PM; ++i) poff[i] = rowpix[i] * pitch;

        for (int tap = 0; tap < ntaps; ++tap) {
            const int kt = tap / p.KF, kf = tap - kt * p.KF;
            const int tapoff = (kt * FW + kf) * pitch;
            for (int cc = 0; cc < C; cc += KC) {
                if (ci + 1 < nch) load_w((ci + 1) & 1, ci + 1);
                cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                const float* Wb = Ws + (ci & 1) * (KC * COUT) + tx * 4;
                const float* At = tile + tapoff + cc;
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    float4 a[PM];
#pragma unroll
                    for (int i = 0; i < PM; ++i)
                        a[i] = *reinterpret_cast<const float4*>(At + poff[i] + k4 * 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        float4 w[CN / 4];
#pragma unroll
                        for (int j = 0; j < CN / 4; ++j)
                            w[j] = *reinterpret_cast<const float4*>(Wb + (k4 * 4 + kk) * COUT + j * (4 * TX));
#pragma unroll
                        for (int i = 0; i < PM; ++i) {
                            const float av = f4_get(a[i], kk);
#pragma unroll
                            for (int j = 0; j < CN / 4; ++j) {
                                acc[i][j * 4 + 0] = fmaf(av, w[j].x, acc[i][j * 4 + 0]);
                                acc[i][j * 4 + 1] = fmaf(av, w[j].y, acc[i][j * 4 + 1]);
                                acc[i][j * 4 + 2] = fmaf(av, w[j].z, acc[i][j * 4 + 2]);
                                acc[i][j * 4 + 3] = fmaf(av, w[j].w, acc[i][j * 4 + 3]);
                            }
                        }
                    }
                }
                __syncthreads();
                ++ci;
            }
        }
    }
    cp_async_wait<0>();

    // ------------------------------------------------------------------ epilogue
    const int c0 = tx * CN;   // first logical conv channel of this thread
    float bias[CN];
#pragma unroll
    for (int e = 0; e < CN; ++e) bias[e] = __ldg(p.bias + c0 + e);
    const float alpha = (EPI == EPI_BIAS) ? 0.0f : __ldg(p.alpha);

#pragma unroll
    for (int i = 0; i < PM; ++i) {
        const int pix = py + PG * i;
        const int fl = pix & (FT - 1);
        const int tt = (pix >> p.lFT) & (TT - 1);
        const int g = pix >> (p.lFT + p.lTT);
        const size_t frame = (size_t)(b0 + g) * p.T + (t0 + tt);
        const int f = f0 + fl;
        float v[CN];
#pragma unroll
        for (int e = 0; e < CN; ++e) v[e] = acc[i][e] + bias[e];

        if (EPI == EPI_BIAS) {
            if (pvalid[i]) {
                float* o = p.out + (frame * p.F_out + f) * COUT + c0;
#pragma unroll
                for (int j = 0; j < CN / 4; ++j)
                    *reinterpret_cast<float4*>(o + j * 4) = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
            }
            continue;
        }

        // LayerNorm groups: EPI_LN: all COUT channels; EPI_SHUF32: even / odd conv channels (32 each);
        // EPI_SHUF64: the half (64 channels) this thread's channels fall in.
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int e = 0; e < CN; e += 2) { s0 += v[e]; s1 += v[e + 1]; }
        constexpr int RED = (EPI == EPI_SHUF64) ? TX / 2 : TX;          // lanes per LN group
        constexpr float INVN = (EPI == EPI_LN) ? 1.0f / COUT : (EPI == EPI_SHUF32 ? 2.0f / COUT : 2.0f / COUT);
        if (EPI != EPI_SHUF32) { s0 += s1; s1 = 0.0f; }
#pragma unroll
        for (int m = 1; m < RED; m <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, m);
            if (EPI == EPI_SHUF32) s1 += __shfl_xor_sync(0xffffffffu, s1, m);
        }
        const float mean0 = s0 * INVN;
        const float mean1 = (EPI == EPI_SHUF32) ? s1 * INVN : mean0;
        float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
        for (int e = 0; e < CN; e += 2) {
            const float d0 = v[e] - mean0, d1 = v[e + 1] - mean1;
            q0 = fmaf(d0, d0, q0);
            q1 = fmaf(d1, d1, q1);
        }
        if (EPI != EPI_SHUF32) { q0 += q1; q1 = 0.0f; }
#pragma unroll
        for (int m = 1; m < RED; m <<= 1) {
            q0 += __shfl_xor_sync(0xffffffffu, q0, m);
            if (EPI == EPI_SHUF32) q1 += __shfl_xor_sync(0xffffffffu, q1, m);
        }
        const float inv0 = rsqrtf(q0 * INVN + LN_EPS);
        const float inv1 = (EPI == EPI_SHUF32) ? rsqrtf(q1 * INVN + LN_EPS) : inv0;

        if (!pvalid[i]) continue;

        if (EPI == EPI_LN) {
            float* o = p.out + (frame * p.F_out + f) * COUT + c0;
#pragma unroll
            for (int j = 0; j < CN / 4; ++j) {
                float r[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + j * 4 + e;
                    const float sc = inv0 * __ldg(p.gamma + c);
                    float y = fmaf(v[j * 4 + e], sc, __ldg(p.beta + c) - mean0 * sc);
                    r[e] = y >= 0.0f ? y : alpha * y;
                }
                *reinterpret_cast<float4*>(o + j * 4) = make_float4(r[0], r[1], r[2], r[3]);
            }
        } else if (EPI == EPI_SHUF32) {
            // out[frame, 2f+j, i] = y[frame, f, 2i+j]   (C_out = COUT/2)
            constexpr int CO = COUT / 2;
            const int i0 = c0 / 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float mean = j ? mean1 : mean0, inv = j ? inv1 : inv0;
                float* o = p.out + ((frame * p.F_out + f) * 2 + j) * CO + i0;
#pragma unroll
                for (int q = 0; q < CN / 8; ++q) {
                    float r[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ii = i0 + q * 4 + e;
                        const float sc = inv * __ldg(p.gamma + ii);
                        float y = fmaf(v[(q * 4 + e) * 2 + j], sc, __ldg(p.beta + ii) - mean * sc);
                        r[e] = y >= 0.0f ? y : alpha * y;
                    }
                    *reinterpret_cast<float4*>(o + q * 4) = make_float4(r[0], r[1], r[2], r[3]);
                }
            }
        } else {  // EPI_SHUF64: out[frame, 2f+h, 32j+i] = y[frame, f, 64h+2i+j]   (C_out = COUT/2 = 64)
            constexpr int CO = COUT / 2;
            const int h = c0 / CO;
            const int i0 = (c0 % CO) / 2;
            float* obase = p.out + ((frame * p.F_out + f) * 2 + h) * CO;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int q = 0; q < CN / 8; ++q) {
                    float r[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int oc = (CO / 2) * j + i0 + q * 4 + e;
                        const float sc = inv0 * __ldg(p.gamma + oc);
                        float y = fmaf(v[(q * 4 + e) * 2 + j], sc, __ldg(p.beta + oc) - mean0 * sc);
                        r[e] = y >= 0.0f ? y : alpha * y;
                    }
                    *reinterpret_cast<float4*>(obase + (CO / 2) * j + i0 + q * 4) = make_float4(r[0], r[1], r[2], r[3]);
                }
            }
        }
    }
}

// Column permutation of the packed weights: shared-memory column q holds logical conv channel c(q) so
// that lane tx reads its CN contiguous logical channels as CN/4 conflict-free float4s.
__host__ inline int conv_col_to_channel(int q, int COUT, int CN) {
    const int TX = COUT / CN;
    return ((q % (4 * TX)) / 4) * CN + (q / (4 * TX)) * 4 + (q % 4);
}

}  // namespace nunet
