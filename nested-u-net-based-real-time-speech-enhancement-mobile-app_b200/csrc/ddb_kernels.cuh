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

// Where "the same unit, `back` steps earlier" lives.  Offline: unit = frame b*T+t of a dense [frame][F][h] tensor, the
// earlier value is `back` frames before (zero when that leaves the clip).  Streaming (converter_nunet_tls.py:373-411:
// layer k keeps the last d rows of its input): unit = stream, every intermediate tensor is a per-stream ring of
// DDB_RING steps [stream][DDB_RING][F][h], `slot` = current step; rings start as zeros = the reference's zero history.
constexpr int DDB_RING = 64;   // > the deepest look-back (32) so that writing step s never clobbers step s - 32
struct DdbGeom {
    int streaming;
    int slot;      // streaming: current step & (DDB_RING - 1)
    int T;         // offline: frames per clip
};
__device__ __forceinline__ bool ddb_back_ok(const DdbGeom& g, long long unit, int back) {
    return g.streaming || (int)(unit % g.T) - back >= 0;
}
// element offset (in floats) of (unit, back, bin ff) in an intermediate tensor with F bins x H channels
__device__ __forceinline__ long long ddb_off(const DdbGeom& g, long long unit, int back, int F, int H, int ff) {
    if (g.streaming) return ((unit * DDB_RING + ((g.slot - back) & (DDB_RING - 1))) * F + ff) * (long long)H;
    return ((unit - back) * F + ff) * (long long)H;
}

// `in`: x [unit][F][C] -> out0, causal (2,3) conv + bias + PReLU.  One thread per (pixel, co).  The previous row of x is
// the same tensor one frame earlier (offline) or the other parity's buffer `x_prev` (streaming).
template <bool SH>
__global__ void __launch_bounds__(128) ddb_in_kernel(const void* __restrict__ x, const void* __restrict__ x_prev,
                                                    const float* __restrict__ w /*[2][3][C][h]*/, const float* __restrict__ b,
                                                    const float* __restrict__ alpha, float* __restrict__ out0, long long units,
                                                    DdbGeom g, int F, int C) {
    const int h = C >> 1;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt / h;
    const int co = (int)(gt - pix * h);
    if (pix >= units * F) return;
    const long long unit = pix / F;
    const int f = (int)(pix - unit * F);
    float acc = __ldg(b + co);
    for (int kt = 0; kt < 2; ++kt) {
        if (kt == 0 && !ddb_back_ok(g, unit, 1)) continue;
        const void* src = (kt == 0 && g.streaming) ? x_prev : x;
        const long long u = (kt == 0 && !g.streaming) ? unit - 1 : unit;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f - 1 + kf;
            if (ff < 0 || ff >= F) continue;
            const float* wk = w + (size_t)((kt * 3 + kf) * C) * h + co;
            for (int ci = 0; ci < C; ++ci) acc = fmaf(act_load<SH>(src, u, F, C, ff, ci), __ldg(wk + (size_t)ci * h), acc);
        }
    }
    const float a = __ldg(alpha);
    out0[ddb_off(g, unit, 0, F, h, f) + co] = acc >= 0.f ? acc : a * acc;
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
                                                       float* __restrict__ outk, int outk_is_ring, long long units, DdbGeom geo, int F) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long pix = gt / H;
    const int g = (int)(gt - pix * H);
    const bool ok = pix < units * F;
    if (!ok) pix = units * F - 1;            // keep the lane alive for the shuffles
    const long long unit = pix / F;
    const int f = (int)(pix - unit * F);
    float z = __ldg(b0 + g);
    for (int kt = 0; kt < 2; ++kt) {
        const int back = d * (1 - kt);
        if (!ddb_back_ok(geo, unit, back)) continue;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f + (kf - 1) * d;
            if (ff < 0 || ff >= F) continue;
            const long long p2 = ddb_off(geo, unit, back, F, H, ff);
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
    // out_1..out_5 are rings when streaming; out_6 is a plain (ping-ponged) tensor
    const long long oo = outk_is_ring ? ddb_off(geo, unit, 0, F, H, f) : pix * H;
    if (ok) outk[oo + g] = r >= 0.f ? r : a * r;
}

// `out`: out6 [frame][F][h] -> y [frame][F][C] (activation tensor), causal (2,3) conv + bias + PReLU.
template <bool SH>
__global__ void __launch_bounds__(128) ddb_out_kernel(const float* __restrict__ o6, const float* __restrict__ o6_prev,
                                                     const float* __restrict__ w /*[2][3][h][C]*/, const float* __restrict__ b,
                                                     const float* __restrict__ alpha, void* __restrict__ y, long long units,
                                                     DdbGeom g, int F, int C) {
    const int h = C >> 1;
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pix = gt / C;
    const int co = (int)(gt - pix * C);
    if (pix >= units * F) return;
    const long long frame = pix / F;
    const int f = (int)(pix - frame * F);
    float acc = __ldg(b + co);
    for (int kt = 0; kt < 2; ++kt) {
        if (kt == 0 && !ddb_back_ok(g, frame, 1)) continue;
        const float* base = (kt == 0 && g.streaming) ? o6_prev : o6;
        const long long u = (kt == 0 && !g.streaming) ? frame - 1 : frame;
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f - 1 + kf;
            if (ff < 0 || ff >= F) continue;
            const float* src = base + (u * F + ff) * h;
            const float* wk = w + (size_t)((kt * 3 + kf) * h) * C + co;
            for (int ci = 0; ci < h; ++ci) acc = fmaf(__ldg(src + ci), __ldg(wk + (size_t)ci * C), acc);
        }
    }
    const float a = __ldg(alpha);
    act_store<SH>(y, frame, F, C, f, co, acc >= 0.f ? acc : a * acc);
}

}  // namespace nunet
