// One kernel per dilated dense block (models/nunet_tls.py:190-272, :383-410; one-frame form converter_nunet_tls.py:373-411):
// `in` conv, the six dilated grouped layers (+ pointwise, LayerNorm, PReLU) and the `out` conv run back to back in one CTA
// per clip (offline) or per group of streams (streaming), separated by block barriers.  A dilated layer looks d = 1..32
// frames back into the lower layers' outputs of the SAME clip, so a clip never needs another CTA's results; the seven
// intermediate tensors stay fp32 in global memory (L2-resident, read back with plain loads: they were written by this
// kernel).  Replaces 8 launches per block (in, 6 layers, out: 104 per step) by one (13 per step); the arithmetic of every
// output element is the former kernels' (ddb_kernels.cuh), operation for operation.
//
// Time-chunked offline calls: `hist` tensors [clip][32][F][h] hold the last 32 frames of out_0..out_5 of the previous chunk
// (and one row of the block input / of out_6 for the two causal (2,3) convs); frames before the chunk are read from there.
#pragma once
#include "ddb_kernels.cuh"

namespace nunet {

constexpr int DDBF_THREADS = 512;
constexpr int DDB_HIST = 32;     // frames of history a chunk needs (deepest dilation)

struct DdbFused {
    const void* x;            // block input (activation tensor: sh16 or fp32) [unit][F][C]
    const void* x_prev;       // streaming: the other parity's buffer; offline: null
    void* y;                  // block output (activation tensor)
    float* mid[7];            // out_0..out_6 (fp32; streaming: out_0..out_5 are per-stream rings of DDB_RING steps)
    const float* mid6_prev;   // streaming: other parity of out_6
    // offline time chunks (null = the chunk starts its clips): last rows of the previous chunk, per clip
    const float* hist_mid[6]; // [clip][DDB_HIST][F][h]   frames -32..-1 of out_0..out_5
    const float* hist_x;      // [clip][F][C] fp32         frame -1 of the block input
    const float* hist_mid6;   // [clip][F][h]              frame -1 of out_6
    const float *w_in, *b_in, *a_in, *w_out, *b_out, *a_out;
    const float *w0[6], *b0[6], *w1[6], *b1[6], *gamma[6], *beta[6], *alpha[6];
    DdbGeom g;
    int F;
    long long units;          // all units of the launch
    int units_per_cta;        // offline: T (one clip per CTA); streaming: streams per CTA
};

// value of an intermediate (fp32 [unit][F][H]) `back` steps before `unit`; handles rings (streaming) and chunk history
__device__ __forceinline__ float ddbf_mid(const DdbFused& p, const float* t, const float* hist, long long unit, int back, int F, int H, int ff,
                                          int ch, bool* ok) {
    *ok = true;
    if (p.g.streaming) return t[ddb_off(p.g, unit, back, F, H, ff) + ch];
    const int tt = (int)(unit % p.g.T) - back;
    if (tt >= 0) return t[((unit - back) * F + ff) * (long long)H + ch];
    if (hist) return hist[(((unit / p.g.T) * DDB_HIST + (DDB_HIST + tt)) * F + ff) * (long long)H + ch];
    *ok = false;
    return 0.0f;
}

// Causal (2,3) conv + bias + PReLU over the pixels [pix_lo, pix_hi) of this CTA: NG groups of 128 threads take 32-pixel tiles
// in turn (register-blocked as ddb_conv23_kernel).  Weights must already be staged in ws.
template <int CIN, int COUT, bool IN_ACT, bool OUT_ACT, bool SH, int NG>
__device__ __forceinline__ void ddbf_conv23(const DdbFused& p, const float* ws, float* xs_all, long long pix_lo, long long pix_hi) {
    constexpr int PX = 32, CPT = COUT / 16;
    const int grp = threadIdx.x >> 7, lt = threadIdx.x & 127;
    const int F = p.F;
    const float* bias = IN_ACT ? p.b_in : p.b_out;
    const float a = __ldg(IN_ACT ? p.a_in : p.a_out);
    float* xs = xs_all + (size_t)grp * 6 * CIN * PX;
    const long long ntile = (pix_hi - pix_lo + PX - 1) / PX;
    for (long long tb = 0; tb < ntile; tb += NG) {
        const long long tile = tb + grp;
        const long long pix0 = pix_lo + tile * PX;
        const bool active = grp < NG && tile < ntile;
        if (active) {
            for (int i = lt; i < PX * 6 * (CIN / 8); i += 128) {
                const int px = i % PX, rest = i / PX;
                const int c8 = rest % (CIN / 8), tap = rest / (CIN / 8);
                const int kt = tap / 3, kf = tap - kt * 3;
                const long long pix = pix0 + px;
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (pix < pix_hi) {
                    const long long unit = pix / F;
                    const int f = (int)(pix - unit * F), ff = f - 1 + kf;
                    if (ff >= 0 && ff < F) {
                        const bool stream_prev = (kt == 0 && p.g.streaming);
                        const bool in_clip = p.g.streaming || kt == 1 || (int)(unit % p.g.T) >= 1;
                        if (IN_ACT) {
                            if (in_clip) {
                                const void* src = stream_prev ? p.x_prev : p.x;
                                const long long u = (kt == 0 && !p.g.streaming) ? unit - 1 : unit;
                                if (SH) {
                                    sh16_load8(reinterpret_cast<const uint8_t*>(src) + u * ((long long)F * CIN * 4), F, CIN, ff, c8, v);
                                } else {
                                    const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (u * F + ff) * (long long)CIN + c8 * 8);
                                    const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1);
                                    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                                }
                            } else if (p.hist_x) {
                                const float4* p4 = reinterpret_cast<const float4*>(p.hist_x + ((unit / p.g.T) * F + ff) * (long long)CIN + c8 * 8);
                                const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1);
                                v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                            }
                        } else {   // intermediate out_6: plain loads (written by this kernel)
                            const float* src = nullptr;
                            if (in_clip) src = (stream_prev ? p.mid6_prev : p.mid[6]) + (((kt == 0 && !p.g.streaming) ? unit - 1 : unit) * F + ff) * (long long)CIN + c8 * 8;
                            else if (p.hist_mid6) src = p.hist_mid6 + ((unit / p.g.T) * F + ff) * (long long)CIN + c8 * 8;
                            if (src) {
                                const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *(reinterpret_cast<const float4*>(src) + 1);
                                v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                            }
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) xs[(tap * CIN + c8 * 8 + e) * PX + px] = v[e];
            }
        }
        __syncthreads();
        if (active) {
            const int cg = lt & 15, pg = lt >> 4;
            float acc[4][CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const float bv = __ldg(bias + cg * CPT + j);
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
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long pix = pix0 + pg * 4 + i;
                if (pix >= pix_hi) continue;
                const long long unit = pix / F;
                const int f = (int)(pix - unit * F);
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const float r = acc[i][j] >= 0.f ? acc[i][j] : a * acc[i][j];
                    const int co = cg * CPT + j;
                    if (OUT_ACT) act_store<SH>(p.y, unit, F, COUT, f, co, r);
                    else p.mid[0][ddb_off(p.g, unit, 0, F, COUT, f) + co] = r;
                }
            }
        }
        __syncthreads();
    }
}

// dilated layer K over the pixels [pix_lo, pix_hi): H lanes per pixel (lane = output channel), as ddb_layer_kernel.  All 6 K input
// values of an output are fetched before the first multiply-add (addresses and validity are computed branch-free), so the loads
// -- L2 hits on what this kernel wrote a moment ago -- overlap instead of queueing behind one another; the sum itself runs in the
// former kernel's order.
template <int H, int K>
__device__ __forceinline__ void ddbf_layer(const DdbFused& p, long long pix_lo, long long pix_hi) {
    constexpr int d = 1 << (K - 1);
    const int F = p.F;
    const int g = threadIdx.x % H;
    constexpr int per = DDBF_THREADS / H;
    const float *w0 = p.w0[K - 1], *w1 = p.w1[K - 1];
    const float b0 = __ldg(p.b0[K - 1] + g), b1 = __ldg(p.b1[K - 1] + g), gm = __ldg(p.gamma[K - 1] + g), bt = __ldg(p.beta[K - 1] + g);
    const float a = __ldg(p.alpha[K - 1]);
    float* outk = p.mid[K];
    const bool ring_out = p.g.streaming && K < 6;
    float wv[6 * K];
#pragma unroll
    for (int i = 0; i < 6 * K; ++i) wv[i] = __ldg(w0 + (size_t)i * H + g);
    // source tensor / channel of input j of this lane's group: channel c = g K + j of cat[out_{K-1}, .., out_0]
    const float* src[K];
    const float* hsrc[K];
    int ch[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int c = g * K + j, m = c / H;
        ch[j] = c - m * H;
        src[j] = p.mid[K - 1 - m];
        hsrc[j] = p.hist_mid[K - 1 - m];
    }
    for (long long base = pix_lo; base < pix_hi; base += per) {
        long long pix = base + threadIdx.x / H;
        const bool ok = pix < pix_hi;
        if (!ok) pix = pix_hi - 1;            // keep the lane alive for the shuffles
        const long long unit = pix / F;
        const int f = (int)(pix - unit * F);
        const int tclip = p.g.streaming ? 0 : (int)(unit % p.g.T);
        const long long clip = p.g.streaming ? 0 : unit / p.g.T;
        float v[6 * K];
        unsigned long long have_mask = 0ull;          // bit i: input i exists (inside the clip / its history, inside the bins)
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const int back = d * (1 - kt);
            const int tt = tclip - back;
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                const int ff = f + (kf - 1) * d;
                const bool fin = ff >= 0 && ff < F;
                const int ffc = fin ? ff : f;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int i = (kt * 3 + kf) * K + j;
                    const float* ptr = src[j] + (unit * F + f) * (long long)H + ch[j];          // always a valid address (value unused)
                    bool have = fin;
                    if (p.g.streaming) {
                        ptr = src[j] + ddb_off(p.g, unit, back, F, H, ffc) + ch[j];
                    } else if (tt >= 0) {
                        ptr = src[j] + ((unit - back) * F + ffc) * (long long)H + ch[j];
                    } else if (hsrc[j]) {
                        ptr = hsrc[j] + ((clip * DDB_HIST + (DDB_HIST + tt)) * F + ffc) * (long long)H + ch[j];
                    } else {
                        have = false;
                    }
                    v[i] = *ptr;
                    have_mask |= (unsigned long long)(have ? 1 : 0) << i;
                }
            }
        }
        float z = b0;
#pragma unroll
        for (int i = 0; i < 6 * K; ++i)
            if ((have_mask >> i) & 1ull) z = fmaf(v[i], wv[i], z);
        float y = b1;
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
        const float inv = rsqrtf(q * (1.0f / H) + LN_EPS) * gm;
        const float r = fmaf(y, inv, bt - mean * inv);
        const long long oo = ring_out ? ddb_off(p.g, unit, 0, F, H, f) : pix * H;
        if (ok) outk[oo + g] = r >= 0.f ? r : a * r;
    }
}

template <int C, bool SH>
__global__ void __launch_bounds__(DDBF_THREADS, 1) ddb_block_kernel(const __grid_constant__ DdbFused p) {
    constexpr int H = C / 2;
    constexpr int NG = (C == 32) ? 4 : 2;          // 32-pixel tiles in flight per conv round (shared-memory budget)
    extern __shared__ __align__(16) float sm[];
    float* ws = sm;                                 // [6][CIN][COUT] of the conv being run (6 C H floats either way)
    float* xs = sm + 6 * C * H;                     // NG input patches
    const long long u0 = (long long)blockIdx.x * p.units_per_cta;
    const long long u1 = min(p.units, u0 + p.units_per_cta);
    if (u0 >= u1) return;
    const long long pix_lo = u0 * p.F, pix_hi = u1 * p.F;
    for (int i = threadIdx.x; i < 6 * C * H / 4; i += DDBF_THREADS) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(p.w_in) + i);
    __syncthreads();
    ddbf_conv23<C, H, true, false, SH, NG>(p, ws, xs, pix_lo, pix_hi);
    for (int i = threadIdx.x; i < 6 * C * H / 4; i += DDBF_THREADS) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(p.w_out) + i);
    ddbf_layer<H, 1>(p, pix_lo, pix_hi);
    __syncthreads();
    ddbf_layer<H, 2>(p, pix_lo, pix_hi);
    __syncthreads();
    ddbf_layer<H, 3>(p, pix_lo, pix_hi);
    __syncthreads();
    ddbf_layer<H, 4>(p, pix_lo, pix_hi);
    __syncthreads();
    ddbf_layer<H, 5>(p, pix_lo, pix_hi);
    __syncthreads();
    ddbf_layer<H, 6>(p, pix_lo, pix_hi);
    __syncthreads();
    ddbf_conv23<H, C, false, true, SH, NG>(p, ws, xs, pix_lo, pix_hi);
}

// After a chunk of T frames: keep the last DDB_HIST frames of out_0..out_5, and frame T - 1 of the block input (as fp32) and of
// out_6, per clip, for the next chunk.  hist[clip][j] <- row T + j of the concatenation [old hist (32 rows) | this chunk (T rows)].
struct DdbHistW {
    float* mid[6];
    float* x;
    float* mid6;
};
template <bool SH>
__global__ void __launch_bounds__(256) ddb_hist_update_kernel(const __grid_constant__ DdbFused p, DdbHistW hw, int C, int have_hist) {
    const long long b = blockIdx.x;
    const int T = p.g.T, F = p.F, H = C / 2;
    const int m = blockIdx.y;                       // 0..5: out_m; 6: x row and out_6 row
    if (m < 6) {
        float* h = hw.mid[m] + b * (long long)DDB_HIST * F * H;
        const float* t = p.mid[m] + b * (long long)T * F * H;
        const int row = F * H;
        // when T < 32 the rows that stay come from the old history: read everything first (registers), then write
        for (int e = threadIdx.x; e < row; e += blockDim.x) {
            float v[DDB_HIST];
#pragma unroll
            for (int j = 0; j < DDB_HIST; ++j) {
                const int i = T + j;                // index into the concatenation
                v[j] = (i < DDB_HIST) ? (have_hist ? h[(long long)i * row + e] : 0.0f) : t[(long long)(i - DDB_HIST) * row + e];
            }
#pragma unroll
            for (int j = 0; j < DDB_HIST; ++j) h[(long long)j * row + e] = v[j];
        }
    } else {
        for (int e = threadIdx.x; e < F * C; e += blockDim.x) {
            const int f = e / C, c = e - f * C;
            hw.x[b * (long long)F * C + e] = act_load<SH>(p.x, b * T + (T - 1), F, C, f, c);
        }
        for (int e = threadIdx.x; e < F * H; e += blockDim.x) hw.mid6[b * (long long)F * H + e] = p.mid[6][(b * T + (T - 1)) * (long long)F * H + e];
    }
}

}  // namespace nunet
