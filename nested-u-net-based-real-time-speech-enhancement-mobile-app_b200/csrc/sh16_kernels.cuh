// The non-convolution units of the offline path over "sh16" activations (conv_tc3.cuh: one 4*C-byte record per
// pixel, [C halves hi][C halves lo], x ~= hi + lo).  Same arithmetic as their fp32 twins in misc_kernels.cuh (which
// the streaming plan keeps using); only the loads join hi + lo and the stores split.
#pragma once
#include "conv_tc3.cuh"
#include "misc_kernels.cuh"

namespace nunet {

__device__ __forceinline__ void sh16_load8(const uint8_t* rec, int C, int c8 /*chunk of 8 channels*/, float* v) {
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(rec + c8 * 16));
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(rec + C * 2 + c8 * 16));
    join8(hi, lo, v);
}
__device__ __forceinline__ void sh16_store8(uint8_t* rec, int C, int c8, const float* v) {
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(rec + c8 * 16) = hi;
    *reinterpret_cast<uint4*>(rec + C * 2 + c8 * 16) = lo;
}

// input_layer = inconv(64) on the 1-channel magnitude (models/proposed.py:293): 8 lanes per pixel, 8 channels per lane.
__global__ void __launch_bounds__(256) input_layer_sh_kernel(const float* __restrict__ mag, const float* __restrict__ w,
                                                            const float* __restrict__ b, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ alpha,
                                                            uint8_t* __restrict__ out, long long npix) {
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
    sh16_store8(out + pix * 256, 64, l, r);
}

// out_conv: Conv2D 1x1, 64 -> 1 (models/proposed.py:615).  8 lanes per pixel.
__global__ void __launch_bounds__(256) out_conv_sh_kernel(const uint8_t* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ out, long long npix,
                                                         int F, int out_stride, int out_off) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt >> 3;
    const int l = (int)(gt & 7);
    float s = 0.0f;
    if (pix < npix) {
        float v[8];
        sh16_load8(x + pix * 256, 64, l, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(v[e], __ldg(w + l * 8 + e), s);
    }
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (pix < npix && l == 0) {
        const long long frame = pix / F;
        const int f = (int)(pix - frame * F);
        out[frame * out_stride + out_off + f] = s + __ldg(b);
    }
}

// CTFA stage 1 (models/proposed.py:125): TA[frame, c] = MLP(mean_f x[frame, f, c]).  One CTA per frame:
// 256 threads = 32 bin rows x 8 channel chunks.
__global__ void __launch_bounds__(256) ctfa_ta_sh_kernel(const uint8_t* __restrict__ x, MlpW ta, float* __restrict__ ta_out, int F) {
    __shared__ float part[32][65];
    __shared__ float mean_s[64];
    __shared__ float h_s[16];
    const int frame = blockIdx.x;
    const int c8 = threadIdx.x & 7, ry = threadIdx.x >> 3;
    const uint8_t* xf = x + (size_t)frame * F * 256;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int f = ry; f < F; f += 32) {
        float v[8];
        sh16_load8(xf + (size_t)f * 256, 64, c8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[ry][c8 * 8 + e] = s[e];
    __syncthreads();
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        float a = 0.0f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) a += part[r][c];
        mean_s[c] = a / (float)F;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        const float t = ctfa_mlp(mean_s, h_s, ta, c, 1);
        ta_out[(size_t)frame * 64 + c] = t;
    }
}

// CTFA stage 3 + residual (models/proposed.py:319): out = x * gate[frame] + res, 8 channels per thread.
__global__ void __launch_bounds__(256) gate_residual_sh_kernel(const uint8_t* __restrict__ x, const uint8_t* __restrict__ res,
                                                              const float* __restrict__ gate /*[frames][64]*/,
                                                              uint8_t* __restrict__ out, long long n8, int F) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const long long pix = i >> 3;
        const int c8 = (int)(i & 7);
        const long long frame = pix / F;
        float xv[8], rv[8];
        sh16_load8(x + pix * 256, 64, c8, xv);
        sh16_load8(res + pix * 256, 64, c8, rv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + frame * 64 + c8 * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + frame * 64 + c8 * 8) + 1);
        float o[8];
        o[0] = fmaf(xv[0], g0.x, rv[0]); o[1] = fmaf(xv[1], g0.y, rv[1]);
        o[2] = fmaf(xv[2], g0.z, rv[2]); o[3] = fmaf(xv[3], g0.w, rv[3]);
        o[4] = fmaf(xv[4], g1.x, rv[4]); o[5] = fmaf(xv[5], g1.y, rv[5]);
        o[6] = fmaf(xv[6], g1.z, rv[6]); o[7] = fmaf(xv[7], g1.w, rv[7]);
        sh16_store8(out + pix * 256, 64, c8, o);
    }
}

// Small dense Y[r][n] = b[n] + sum_k X[r][k] W[k][n] (LSTM input projection / Dense after the LSTM,
// models/proposed.py:305-309) where X and / or Y rows are sh16 tensors [F_b][C] flattened as k = f*C + c.
template <bool IN_SH, bool OUT_SH>
__global__ void __launch_bounds__(128) dense_rows_sh_kernel(const void* __restrict__ Xv, const float* __restrict__ W,
                                                           const float* __restrict__ bias, void* __restrict__ Yv, long long rows,
                                                           int K, int N, int C /*channels per sh16 pixel*/) {
    extern __shared__ float xs[];   // [DENSE_RB][K]
    const long long r0 = (long long)blockIdx.x * DENSE_RB;
    const int nr = (int)min((long long)DENSE_RB, rows - r0);
    if (IN_SH) {
        const __half* X = reinterpret_cast<const __half*>(Xv);
        for (int i = threadIdx.x; i < DENSE_RB * K; i += blockDim.x) {
            const int r = i / K, k = i - r * K;
            const int f = k / C, c = k - f * C;
            float v = 0.0f;
            if (r < nr) {
                const __half* rec = X + ((r0 + r) * (long long)K + (long long)f * C) * 2;
                v = __half2float(rec[c]) + __half2float(rec[C + c]);
            }
            xs[i] = v;
        }
    } else {
        const float* X = reinterpret_cast<const float*>(Xv);
        for (int i = threadIdx.x; i < DENSE_RB * K; i += blockDim.x) {
            const int r = i / K;
            xs[i] = (r < nr) ? __ldg(X + r0 * K + i) : 0.0f;
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc[DENSE_RB];
        const float bv = __ldg(bias + n);
#pragma unroll
        for (int r = 0; r < DENSE_RB; ++r) acc[r] = bv;
        for (int k = 0; k < K; ++k) {
            const float wv = __ldg(W + (size_t)k * N + n);
#pragma unroll
            for (int r = 0; r < DENSE_RB; ++r) acc[r] = fmaf(xs[r * K + k], wv, acc[r]);
        }
        if (OUT_SH) {
            __half* Y = reinterpret_cast<__half*>(Yv);
            const int f = n / C, c = n - f * C;
            for (int r = 0; r < nr; ++r) {
                __half* rec = Y + ((r0 + r) * (long long)N + (long long)f * C) * 2;
                const __half h = __float2half_rn(acc[r]);
                rec[c] = h;
                rec[C + c] = __float2half_rn(acc[r] - __half2float(h));
            }
        } else {
            float* Y = reinterpret_cast<float*>(Yv);
            for (int r = 0; r < nr; ++r) Y[(r0 + r) * N + n] = acc[r];
        }
    }
}

}  // namespace nunet
