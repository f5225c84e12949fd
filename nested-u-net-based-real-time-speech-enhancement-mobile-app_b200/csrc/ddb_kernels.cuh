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

// Causal (2,3) convolution + bias + PReLU at small F -- the `in` (C -> C/2, activation tensor -> intermediate) and
// `out` (C/2 -> C, intermediate -> activation tensor) layers of the block.  Register-blocked FP32: a CTA of 128
// threads owns 32 pixels x COUT channels; the weights [6][CIN][COUT] and the input patch [6][CIN][32 pixels] are staged
// in shared memory once, then every thread accumulates 4 pixels x (COUT/16) channels with one 16-byte and one
// 8/16-byte shared load per 4*(COUT/16) FMAs.  The previous row is the same tensor one unit earlier (offline) or the
// other parity's buffer `x_prev` (streaming).  IN_ACT / OUT_ACT: the tensor on that side is a plan activation tensor
// (sh16 planar when SH, else fp32 [unit][F][C]); otherwise it is an fp32 intermediate addressed through DdbGeom.
template <int CIN, int COUT, bool IN_ACT, bool OUT_ACT, bool SH>
__global__ void __launch_bounds__(128) ddb_conv23_kernel(const void* __restrict__ x, const void* __restrict__ x_prev,
                                                        const float* __restrict__ w /*[2][3][CIN][COUT]*/,
                                                        const float* __restrict__ b, const float* __restrict__ alpha,
                                                        void* __restrict__ y, long long units, DdbGeom g, int F) {
    constexpr int PX = 32, CPT = COUT / 16;          // pixels per CTA, channels per thread
    extern __shared__ __align__(16) float sm[];
    float* ws = sm;                                  // [6][CIN][COUT]
    float* xs = sm + 6 * CIN * COUT;                 // [6][CIN][PX]
    const long long pix0 = (long long)blockIdx.x * PX;
    const long long npix = units * F;
    for (int i = threadIdx.x; i < 6 * CIN * COUT / 4; i += 128)
        reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    // input patch: item = (pixel, tap, 8-channel chunk)
    for (int i = threadIdx.x; i < PX * 6 * (CIN / 8); i += 128) {
        const int px = i % PX, rest = i / PX;
        const int c8 = rest % (CIN / 8), tap = rest / (CIN / 8);
        const int kt = tap / 3, kf = tap - kt * 3;
        const long long pix = pix0 + px;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (pix < npix) {
            const long long unit = pix / F;
            const int f = (int)(pix - unit * F), ff = f - 1 + kf;
            if (ff >= 0 && ff < F && (kt == 1 || ddb_back_ok(g, unit, 1))) {
                if (IN_ACT) {
                    const void* src = (kt == 0 && g.streaming) ? x_prev : x;
                    const long long u = (kt == 0 && !g.streaming) ? unit - 1 : unit;
                    if (SH) {
                        sh16_load8(reinterpret_cast<const uint8_t*>(src) + u * ((long long)F * CIN * 4), F, CIN, ff, c8, v);
                    } else {
                        const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (u * F + ff) * (long long)CIN + c8 * 8);
                        const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1);
                        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                    }
                } else {   // intermediate: plain tensor, previous row = other parity (streaming) or previous unit
                    const float* src = reinterpret_cast<const float*>((kt == 0 && g.streaming) ? x_prev : x);
                    const long long u = (kt == 0 && !g.streaming) ? unit - 1 : unit;
                    const float4* p4 = reinterpret_cast<const float4*>(src + (u * F + ff) * (long long)CIN + c8 * 8);
                    const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1);
                    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) xs[(tap * CIN + c8 * 8 + e) * PX + px] = v[e];
    }
    __syncthreads();
    const int cg = threadIdx.x & 15, pg = threadIdx.x >> 4;     // channels cg*CPT.., pixels pg*4..
    float acc[4][CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const float bv = __ldg(b + cg * CPT + j);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = bv;
    }
#pragma unroll 4
    for (int k = 0; k < 6 * CIN; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + k * PX + pg * 4);
        float wv[CPT];
        if (CPT == 4) {
            const float4 t = *reinterpret_cast<const float4*>(ws + k * COUT + cg * 4);
            wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
        } else if (CPT == 2) {
            const float2 t = *reinterpret_cast<const float2*>(ws + k * COUT + cg * 2);
            wv[0] = t.x; wv[1] = t.y;
        } else {
            wv[0] = ws[k * COUT + cg];
        }
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(xa[i], wv[j], acc[i][j]);
    }
    const float a = __ldg(alpha);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long pix = pix0 + pg * 4 + i;
        if (pix >= npix) continue;
        const long long unit = pix / F;
        const int f = (int)(pix - unit * F);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const float r = acc[i][j] >= 0.f ? acc[i][j] : a * acc[i][j];
            const int co = cg * CPT + j;
            if (OUT_ACT) act_store<SH>(y, unit, F, COUT, f, co, r);
            else reinterpret_cast<float*>(y)[ddb_off(g, unit, 0, F, COUT, f) + co] = r;
        }
    }
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

}  // namespace nunet
