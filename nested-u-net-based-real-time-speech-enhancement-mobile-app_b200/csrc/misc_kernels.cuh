// Small fused units of the NUNet-TLS path that are not convolutions over (t, f) patches:
// the Cin=1 input layer, the 64->1 output layer, CTFA (causal time-frequency attention), the LSTM
// bottleneck (input projection, recurrence, Dense) and the gate+residual application.
#pragma once
#include "common.cuh"

namespace nunet {

// ---------------------------------------------------------------------------------------------------
// input_layer = inconv(64) on a 1-channel input (models/proposed.py:293, factory :218):
// y[c] = PReLU(LN_c(x * w[c] + b[c])).  8 lanes per pixel, 8 channels per lane.
__global__ void __launch_bounds__(256) input_layer_kernel(const float* __restrict__ mag,   // [npix]
                                                         const float* __restrict__ w,     // [64]
                                                         const float* __restrict__ b,     // [64]
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ alpha, float* __restrict__ out,  // [npix][64]
                                                         long long npix) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt >> 3;
    const int l = (int)(gt & 7);
    const bool ok = pix < npix;
    const float x = ok ? __ldg(mag + pix) : 0.0f;
    float v[8];
    float s = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        v[e] = fmaf(x, __ldg(w + l * 8 + e), __ldg(b + l * 8 + e));
        s += v[e];
    }
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    const float mean = s * (1.0f / 64.0f);
    float q = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float d = v[e] - mean;
        q = fmaf(d, d, q);
    }
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) q += __shfl_xor_sync(0xffffffffu, q, m);
    const float inv = rsqrtf(q * (1.0f / 64.0f) + LN_EPS);
    if (!ok) return;
    const float a = __ldg(alpha);
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = l * 8 + e;
        const float sc = inv * __ldg(gamma + c);
        const float y = fmaf(v[e], sc, __ldg(beta + c) - mean * sc);
        r[e] = y >= 0.0f ? y : a * y;
    }
    float4* o = reinterpret_cast<float4*>(out + pix * 64 + l * 8);
    o[0] = make_float4(r[0], r[1], r[2], r[3]);
    o[1] = make_float4(r[4], r[5], r[6], r[7]);
}

// ---------------------------------------------------------------------------------------------------
// out_conv: Conv2D 1x1, 64 -> 1, bias, no activation (models/proposed.py:615).  16 lanes per pixel.
// `out_stride`/`out_off` let the result land directly in a [.., 257] spectrogram with the DC bin left
// for the caller (zero pad, models/proposed.py:617).
__global__ void __launch_bounds__(256) out_conv_kernel(const float* __restrict__ x,   // [npix][64]
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      float* __restrict__ out, long long npix, int F, int out_stride,
                                                      int out_off) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt >> 4;
    const int l = (int)(gt & 15);
    float s = 0.0f;
    if (pix < npix) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + pix * 64) + l);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + l);
        s = xv.x * wv.x;
        s = fmaf(xv.y, wv.y, s);
        s = fmaf(xv.z, wv.z, s);
        s = fmaf(xv.w, wv.w, s);
    }
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (pix < npix && l == 0) {
        const long long frame = pix / F;
        const int f = (int)(pix - frame * F);
        out[frame * out_stride + out_off + f] = s + __ldg(b);
    }
}

// ---------------------------------------------------------------------------------------------------
// CTFA, models/proposed.py:125 (`ctfa`) and :162 (`ctfa_rt`).
// Stage 1: TA[frame, c] = sigmoid(W1 . relu(W0 . mean_f x[frame, :, c] + b0) + b1)      (64 -> 16 -> 64)
// One CTA (256 threads = 4 bin groups x 64 channels) per frame.
struct MlpW {
    const float* k0;  // [64][16]
    const float* b0;  // [16]
    const float* k1;  // [16][64]
    const float* b1;  // [64]
};

__device__ __forceinline__ float ctfa_mlp(const float* v_s /*[64] smem*/, float* h_s /*[16] smem*/, const MlpW& m,
                                          int c /*0..63, threads 0..63 participate*/, int sync_id) {
    // caller guarantees v_s is visible; uses named barrier over 64 threads (the first two warps)
    if (c < 16) {
        float a = __ldg(m.b0 + c);
#pragma unroll 8
        for (int k = 0; k < 64; ++k) a = fmaf(v_s[k], __ldg(m.k0 + k * 16 + c), a);
        h_s[c] = fmaxf(a, 0.0f);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(sync_id) : "memory");
    float o = __ldg(m.b1 + c);
#pragma unroll
    for (int j = 0; j < 16; ++j) o = fmaf(h_s[j], __ldg(m.k1 + j * 64 + c), o);
    return sigmoidf_(o);
}

__global__ void __launch_bounds__(256) ctfa_ta_kernel(const float* __restrict__ x,  // [frames][F][64]
                                                     MlpW ta, float* __restrict__ ta_out /*[frames][64]*/, int F) {
    __shared__ float part[4][64];
    __shared__ float mean_s[64];
    __shared__ float h_s[16];
    const int frame = blockIdx.x;
    const int c = threadIdx.x & 63, ry = threadIdx.x >> 6;
    const float* xf = x + (size_t)frame * F * 64;
    float s = 0.0f;
    for (int f = ry; f < F; f += 4) s += __ldg(xf + f * 64 + c);
    part[ry][c] = s;
    __syncthreads();
    if (threadIdx.x < 64) {
        mean_s[c] = (part[0][c] + part[1][c] + part[2][c] + part[3][c]) / (float)F;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        const float t = ctfa_mlp(mean_s, h_s, ta, c, 1);
        ta_out[(size_t)frame * 64 + c] = t;
    }
}

// Stage 2: gate[frame, c] = TA * sigmoid(V1 . relu(V0 . avg + c0) + c1), avg = mean of the last 32 TA of the
// clip (zeros before the clip start, always / 32: ZeroPadding2D((31,0)) + AveragePooling1D(32)) or, in the
// one-frame graph, TA / 32.  FA's input is TA broadcast over frequency, so FA does not depend on f.
// ring != nullptr (streaming extension): avg over the per-stream ring of the last 32 TA, slot `ring_pos`.
__global__ void __launch_bounds__(64) ctfa_gate_kernel(const float* __restrict__ ta,  // [frames][64]
                                                      MlpW fa, float* __restrict__ gate, int T, int mode_div32,
                                                      float* __restrict__ ring /*[B][32][64] or null*/, int ring_pos) {
    __shared__ float avg_s[64];
    __shared__ float h_s[16];
    const int frame = blockIdx.x;
    const int c = threadIdx.x;
    const float tv = ta[(size_t)frame * 64 + c];
    float avg;
    if (ring != nullptr) {
        float* rg = ring + (size_t)frame * CTFA_WINDOW * 64;
        rg[ring_pos * 64 + c] = tv;
        float s = 0.0f;
        // oldest -> newest so the summation order matches the offline window
        for (int d = 1; d <= CTFA_WINDOW; ++d) s += rg[((ring_pos + d) & (CTFA_WINDOW - 1)) * 64 + c];
        avg = s * (1.0f / CTFA_WINDOW);
    } else if (mode_div32) {
        avg = tv * (1.0f / CTFA_WINDOW);
    } else {
        const int t = frame % T;
        const int n = min(t + 1, CTFA_WINDOW);
        float s = 0.0f;
        for (int d = n - 1; d >= 0; --d) s += ta[(size_t)(frame - d) * 64 + c];
        avg = s * (1.0f / CTFA_WINDOW);
    }
    avg_s[c] = avg;
    __syncthreads();
    const float f = ctfa_mlp(avg_s, h_s, fa, c, 1);
    gate[(size_t)frame * 64 + c] = f * tv;
}

// Stage 3: out = x * gate (broadcast over f) + residual   (models/proposed.py:319 `ctfa(...) + en_in`)
__global__ void __launch_bounds__(256) gate_residual_kernel(const float4* __restrict__ x, const float4* __restrict__ res,
                                                           const float4* __restrict__ gate /*[frames][16] float4*/,
                                                           float4* __restrict__ out, long long n4, int F) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const long long frame = i / ((long long)F * 16);
        const int c4 = (int)(i & 15);
        const float4 xv = __ldg(x + i), rv = __ldg(res + i), gv = __ldg(gate + frame * 16 + c4);
        out[i] = make_float4(fmaf(xv.x, gv.x, rv.x), fmaf(xv.y, gv.y, rv.y), fmaf(xv.z, gv.z, rv.z),
                             fmaf(xv.w, gv.w, rv.w));
    }
}

// (the LSTM bottleneck -- input projection, recurrence, Dense -- lives in lstm_kernels.cuh)

}  // namespace nunet
