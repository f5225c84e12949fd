// Dilated dense block (DDB) bottleneck of the NUNet-TLS baseline, offline form (models/nunet_tls.py:190-272; use
// :383-410 inside every nested sub-U-Net and :678-700 for the main bottleneck):
//   in    ZeroPad((1,0),(1,1)) + Conv2D(C/2,(2,3)) + PReLU
//   k=1..6  d = 2^(k-1): ZeroPad((d,0),(d,d)) + grouped Conv2D(C/2,(2,3), dilation d in time AND frequency, groups = C/2)
//           over cat[out_{k-1}, .., out_0] (group g reads channels [g k, (g+1) k)) -> Conv2D 1x1 -> LayerNorm -> PReLU
//   out   ZeroPad((1,0),(1,1)) + Conv2D(C,(2,3)) + PReLU
// The block sits at F_b in {1,2,4} bins, 0.4 % of the model's MACs and K = 6..36 per group: nothing here fills an MMA
// fragment, so these are plain FP32 kernels, one lane per output channel.  The seven intermediate tensors out_0..out_6
// are fp32 [frame][F_b][C/2]; the block input / output are the plan's activation tensors (sh16 planar or fp32).
#pragma once
#include "sh16_kernels.cuh"

namespace nunet {

// element (frame row, bin f, channel c) of an activation tensor with F bins and C channels
template <bool SH>
__device__ __forceinline__ float act_load(const void* base, long long frame, int F, int C, int f, int c) {
    if (SH) {
        const __half* row = reinterpret_cast<const __half*>(base) + frame * (long long)F * C * 2;
        return __half2float(row[sh16_half_index(F, C, 0, f, c)]) + __half2float(row[sh16_half_index(F, C, 1, f, c)]);
    }
    return __ldg(reinterpret_cast<const float*>(base) + (frame * F + f) * (long long)C + c);
}
template <bool SH>
__device__ __forceinline__ void act_store(void* base, long long frame, int F, int C, int f, int c, float v) {
    if (SH) {
        __half* row = reinterpret_cast<__half*>(base) + frame * (long long)F * C * 2;
        const __half h = __float2half_rn(v);
        row[sh16_half_index(F, C, 0, f, c)] = h;
        row[sh16_half_index(F, C, 1, f, c)] = __float2half_rn(v - __half2float(h));
    } else {
        reinterpret_cast<float*>(base)[(frame * F + f) * (long long)C + c] = v;
    }
}

// `in`: x [frame][F][C] -> out0 [frame][F][h], causal (2,3) conv + bias + PReLU.  One thread per (pixel, co).
template <bool SH>
__global__ void __launch_bounds__(128) ddb_in_kernel(const void* __restrict__ x, const float* __restrict__ w /*[2][3][C][h]*/,
                                                    const float* __restrict__ b, const float* __restrict__ alpha,
                                                    float* __restrict__ out0, long long frames, int T, int F, int C) {
    const int h = C >> 1;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt / h;
    const int co = (int)(gt - pix * h);
    if (pix >= frames * F) return;
    const long long frame = pix / F;
    const int f = (int)(pix - frame * F);
    const int t = (int)(frame % T);
    float acc = __ldg(b + co);
    for (int kt = 0; kt < 2; ++kt) {
        if (t - 1 + kt < 0) continue;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f - 1 + kf;
            if (ff < 0 || ff >= F) continue;
            const float* wk = w + (size_t)((kt * 3 + kf) * C) * h + co;
            for (int ci = 0; ci < C; ++ci) acc = fmaf(act_load<SH>(x, frame - 1 + kt, F, C, ff, ci), __ldg(wk + (size_t)ci * h), acc);
        }
    }
    const float a = __ldg(alpha);
    out0[pix * h + co] = acc >= 0.f ? acc : a * acc;
}

struct DdbOuts {
    const float* o[6];   // out_0 .. out_5 (layer k reads o[0..k-1])
};

// layer k: H = C/2 lanes per pixel (16 or 32), lane = output channel g.
template <int H>
__global__ void __launch_bounds__(128) ddb_layer_kernel(DdbOuts src, int k, int d, const float* __restrict__ w0 /*[2][3][k][H]*/,
                                                       const float* __restrict__ b0, const float* __restrict__ w1 /*[H][H]*/,
                                                       const float* __restrict__ b1, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ alpha,
                                                       float* __restrict__ outk, long long frames, int T, int F) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long pix = gt / H;
    const int g = (int)(gt - pix * H);
    const bool ok = pix < frames * F;
    if (!ok) pix = frames * F - 1;            // keep the lane alive for the shuffles
    const long long frame = pix / F;
    const int f = (int)(pix - frame * F);
    const int t = (int)(frame % T);
    float z = __ldg(b0 + g);
    for (int kt = 0; kt < 2; ++kt) {
        const int back = d * (1 - kt);
        if (t - back < 0) continue;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f + (kf - 1) * d;
            if (ff < 0 || ff >= F) continue;
            const long long p2 = ((frame - back) * F + ff) * H;
            for (int j = 0; j < k; ++j) {
                const int c = g * k + j;                 // channel of cat[out_{k-1}, .., out_0]
                const int m = c / H, ch = c - m * H;
                z = fmaf(__ldg(src.o[k - 1 - m] + p2 + ch), __ldg(w0 + (size_t)((kt * 3 + kf) * k + j) * H + g), z);
            }
        }
    }
    float y = __ldg(b1 + g);
#pragma unroll
    for (int gg = 0; gg < H; ++gg) y = fmaf(__shfl_sync(0xffffffffu, z, gg, H), __ldg(w1 + gg * H + g), y);
    float s = y;
#pragma unroll
    for (int m = H / 2; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m, H);
    const float mean = s * (1.0f / H);
    const float dv = y - mean;
    float q = dv * dv;
#pragma unroll
    for (int m = H / 2; m >= 1; m >>= 1) q += __shfl_xor_sync(0xffffffffu, q, m, H);
    const float inv = rsqrtf(q * (1.0f / H) + LN_EPS) * __ldg(gamma + g);
    const float r = fmaf(y, inv, __ldg(beta + g) - mean * inv);
    const float a = __ldg(alpha);
    if (ok) outk[pix * H + g] = r >= 0.f ? r : a * r;
}

// `out`: out6 [frame][F][h] -> y [frame][F][C] (activation tensor), causal (2,3) conv + bias + PReLU.
template <bool SH>
__global__ void __launch_bounds__(128) ddb_out_kernel(const float* __restrict__ o6, const float* __restrict__ w /*[2][3][h][C]*/,
                                                     const float* __restrict__ b, const float* __restrict__ alpha,
                                                     void* __restrict__ y, long long frames, int T, int F, int C) {
    const int h = C >> 1;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt / C;
    const int co = (int)(gt - pix * C);
    if (pix >= frames * F) return;
    const long long frame = pix / F;
    const int f = (int)(pix - frame * F);
    const int t = (int)(frame % T);
    float acc = __ldg(b + co);
    for (int kt = 0; kt < 2; ++kt) {
        if (t - 1 + kt < 0) continue;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f - 1 + kf;
            if (ff < 0 || ff >= F) continue;
            const float* src = o6 + ((frame - 1 + kt) * F + ff) * h;
            const float* wk = w + (size_t)((kt * 3 + kf) * h) * C + co;
            for (int ci = 0; ci < h; ++ci) acc = fmaf(__ldg(src + ci), __ldg(wk + (size_t)ci * C), acc);
        }
    }
    const float a = __ldg(alpha);
    act_store<SH>(y, frame, F, C, f, co, acc >= 0.f ? acc : a * acc);
}

}  // namespace nunet
