// The deep, tiny part of a nested sub-U-Net in ONE kernel for the streaming plan: the causal convs whose
// input has <= 32 bins, the LSTM bottleneck and the first sub-pixel convs (output <= 32 bins) -- and, for the blocks that
// start at <= 32 bins, every conv of the block.  At one frame per stream these layers are far too small to fill the GPU: as
// separate launches each costs ~13 us of launch / prologue / drain latency whatever the stream count, which is why a step of
// 1024 streams was latency-bound (174 launches, 2.55 ms against a 0.9 ms HBM floor).
//
// Streams are independent, so a CTA owns a group of G streams through the whole chain: layer l + 1 starts after one block
// barrier, reading what this CTA itself wrote (global memory, still in L1 / L2) plus last step's rows from the other parity's
// buffers (converter_proposed.py:226: Concatenate(axis=1)([prev, cur])).  Every layer output is a state tensor of the next step
// anyway, so nothing extra is written.
//
// Arithmetic: the same split-half scheme as conv_tc3.cuh (x = hi + lo fp16, three products a_hi b_hi + a_hi b_lo + a_lo b_hi,
// fp32 accumulation) on warp-level mma.sync.m16n8k16 -- for M = G x F <= 128 rows per layer a tcgen05 pipeline (TMEM allocation,
// descriptors, mbarrier ring) would spend longer setting up than computing.  A fragments are read straight from the planar sh16
// rows (the 16 bytes of a (row, 8-channel chunk) are exactly what the four lanes of a quad need), B fragments from weights
// pre-packed on the host in fragment order (one 16-byte load per lane, k-step and 8-column tile: b_hi and b_lo), accumulators
// stay in registers, LayerNorm (two-pass) runs over the quad with two shuffles.  One warp = 16 rows x one LayerNorm group of
// columns; the next layer's weights are prefetched into L1 while the current layer computes.
// Results differ from the unfused conv_tc3 path only in summation order (~1e-6).
#pragma once
#include "lstm_kernels.cuh"
#include "tcgen05_ptx.cuh"   // mbarrier / bulk-copy wrappers

namespace nunet {

constexpr int FZ_MAXL = 16;
constexpr int FZ_WARPS = 16;
constexpr int FZ_CTAS_PER_SM = 1;
constexpr int FZ_WBUF = 98304;     // one weight chunk: a whole layer, or one 64-column half of a 128-column sub-pixel unit
constexpr int FZ_SMEM = 2 * FZ_WBUF + 1024;
constexpr int FZ_THREADS = 32 * FZ_WARPS;
enum { FZ_CONV = 0, FZ_LSTM = 1 };
enum { FZE_LN = 0, FZE_SHUF32 = 1, FZE_SHUF64 = 2, FZE_BIAS = 3 };

struct FzLayer {
    int kind;
    // ---- conv: out[s][f'][c] = epi(sum_taps sum_ci in[tap row][f * stride - padl + kf][ci] W[tap][ci][c])
    const uint8_t *a_cur, *a_prev, *b_cur, *b_prev;   // sh16 [streams][F_in][Ca | Cb]; prev = other parity (kt = 0), null when KT == 1
    uint8_t* out;                                     // sh16 [streams][F_out][PC]
    const uint4* w;                                   // [LN group][k-step][n-tile of the group][lane]: {b_hi k 0-7, b_hi k 8-15, b_lo ..}
    const float *bias, *gamma, *beta, *alpha;         // packed-column order (bias), output-channel order (gamma / beta)
    float wscale_inv;
    int F_in, Ca, Cb, F_conv, N, KT, KF, padl, stride, epi, in_eo, out_eo;
    // ---- lstm: x = a_cur [streams][D] as sh16 rows [Fb][C], y = out
    const float *wk, *wr, *bk, *wd, *bd;              // [D][84], [21][84], [84], [21][D], [D]
    float *h, *c;                                     // [streams][21]
    int D, C;
};
struct FzParams {
    FzLayer L[FZ_MAXL];
    int nl, S, G;      // layers, streams of this launch, streams per CTA
    int dbg;           // experiments: 1 = no conv tiles, 2 = no LSTM, 4 = no weight copies, 8 = no prefetches
};

__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ld_b32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }   // plain load: sees this CTA's writes

// One work item = 16 rows x one LayerNorm group of columns (NT tiles of 8: the PC channels of ONE output pixel, or all N
// columns of a bias-only unit), so that a 64- or 128-column sub-pixel unit is shared by two warps per row tile.
template <int NT>
__device__ __forceinline__ void fz_conv_tile(const FzLayer& L, const uint4* wsm /*this group's weights in shared memory*/, int s0, int ns, int mt,
                                             int px, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const int M = ns * L.F_conv;
    const int Ct = L.Ca + L.Cb;
    const int cgn = Ct >> 4;                      // 16-channel groups per tap
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
    // the two rows of this thread: (stream, conv bin)
    int rs[2], rf[2];
    bool rv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = mt * 16 + g + 8 * r;
        rv[r] = row < M;
        rs[r] = rv[r] ? row / L.F_conv : 0;
        rf[r] = rv[r] ? row - rs[r] * L.F_conv : 0;
    }
    // Lean inner loop: per (tap, source) the two row pointers are set up once; a k-step is then 8 loads at fixed offsets from
    // them, NT 16-byte shared-memory loads of weights and 3 NT MMAs.  k order = tap-major, channels of [a | b] ascending: the
    // order the weights are packed in.
    const int plane = L.F_in * 16;                                 // bytes between two 8-channel planes of a source row
    const int taps = L.KT * L.KF;
    const uint4* wp = wsm + lane;
    for (int tap = 0; tap < taps; ++tap) {
        const int kt = tap / L.KF, kf = tap - kt * L.KF;
        const bool use_prev = (L.KT == 2 && kt == 0);
        int poff[2];
        bool ok[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int fi = rf[r] * L.stride - L.padl + kf;
            ok[r] = rv[r] && fi >= 0 && fi < L.F_in;
            const int pos = L.in_eo ? (fi & 1) * (L.F_in >> 1) + (fi >> 1) : fi;
            poff[r] = ok[r] ? pos * 16 + q * 4 : 0;
        }
#pragma unroll 1
        for (int seg = 0; seg < 2; ++seg) {
            const int Cs = seg ? L.Cb : L.Ca;
            if (!Cs) continue;
            const uint8_t* base = seg ? (use_prev ? L.b_prev : L.b_cur) : (use_prev ? L.a_prev : L.a_cur);
            const int lo_off = (Cs >> 3) * plane;                  // hi planes [0, Cs/8), lo planes [Cs/8, Cs/4)
            const uint8_t* p0 = base + (long long)(s0 + rs[0]) * (L.F_in * Cs * 4) + poff[0];
            const uint8_t* p1 = base + (long long)(s0 + rs[1]) * (L.F_in * Cs * 4) + poff[1];
#pragma unroll 2
            for (int cgl = 0; cgl < (Cs >> 4); ++cgl) {
                uint32_t ah[4] = {0u, 0u, 0u, 0u}, al[4] = {0u, 0u, 0u, 0u};
                if (ok[0]) {
                    ah[0] = ld_b32(p0); ah[2] = ld_b32(p0 + plane); al[0] = ld_b32(p0 + lo_off); al[2] = ld_b32(p0 + lo_off + plane);
                }
                if (ok[1]) {
                    ah[1] = ld_b32(p1); ah[3] = ld_b32(p1 + plane); al[1] = ld_b32(p1 + lo_off); al[3] = ld_b32(p1 + lo_off + plane);
                }
                uint4 b[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) b[nt] = wp[nt * 32];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma_f16(acc[nt], ah, b[nt].x, b[nt].y);     // a_hi b_hi
                    mma_f16(acc[nt], ah, b[nt].z, b[nt].w);     // a_hi b_lo
                    mma_f16(acc[nt], al, b[nt].x, b[nt].y);     // a_lo b_hi
                }
                wp += NT * 32;
                p0 += 2 * plane;
                p1 += 2 * plane;
            }
        }
    }
    // ---- epilogue: columns 2q, 2q+1 of every 8-column tile for rows g and g + 8; this warp's columns are one LN group
    constexpr int PC = NT * 8;
    const bool ln = L.epi != FZE_BIAS;
    const int NPX = L.N / PC;
    const float alpha = ln ? __ldg(L.alpha) : 0.0f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float2 bv = __ldg(reinterpret_cast<const float2*>(L.bias + px * PC + nt * 8 + 2 * q));
        acc[nt][0] = fmaf(acc[nt][0], L.wscale_inv, bv.x);
        acc[nt][1] = fmaf(acc[nt][1], L.wscale_inv, bv.y);
        acc[nt][2] = fmaf(acc[nt][2], L.wscale_inv, bv.x);
        acc[nt][3] = fmaf(acc[nt][3], L.wscale_inv, bv.y);
    }
    float mean[2] = {0.f, 0.f}, inv[2] = {1.f, 1.f};
    if (ln) {
        float s[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s[0] += acc[nt][0] + acc[nt][1];
            s[1] += acc[nt][2] + acc[nt][3];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            s[r] += __shfl_xor_sync(0xffffffffu, s[r], 1);
            s[r] += __shfl_xor_sync(0xffffffffu, s[r], 2);
            mean[r] = s[r] * (1.0f / PC);
        }
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float d0 = acc[nt][0] - mean[0], d1 = acc[nt][1] - mean[0], d2 = acc[nt][2] - mean[1], d3 = acc[nt][3] - mean[1];
            v[0] += d0 * d0 + d1 * d1;
            v[1] += d2 * d2 + d3 * d3;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            v[r] += __shfl_xor_sync(0xffffffffu, v[r], 1);
            v[r] += __shfl_xor_sync(0xffffffffu, v[r], 2);
            inv[r] = rsqrtf(v[r] * (1.0f / PC) + LN_EPS);
        }
    }
    const int F_out = L.F_conv * NPX;
    const long long orow = (long long)F_out * PC * 4, oplane = (long long)F_out * 16;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int ch = nt * 8 + 2 * q;          // output channel of acc[nt][0]
        float y[4] = {acc[nt][0], acc[nt][1], acc[nt][2], acc[nt][3]};
        if (ln) {
            const float2 gm = __ldg(reinterpret_cast<const float2*>(L.gamma + ch)), bt = __ldg(reinterpret_cast<const float2*>(L.beta + ch));
            y[0] = (y[0] - mean[0]) * inv[0] * gm.x + bt.x;
            y[1] = (y[1] - mean[0]) * inv[0] * gm.y + bt.y;
            y[2] = (y[2] - mean[1]) * inv[1] * gm.x + bt.x;
            y[3] = (y[3] - mean[1]) * inv[1] * gm.y + bt.y;
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = y[e] >= 0.0f ? y[e] : alpha * y[e];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!rv[r]) continue;
            const int obin = rf[r] * NPX + px;
            const int pos = L.out_eo ? (obin & 1) * (F_out >> 1) + (obin >> 1) : obin;
            const __half2 hh = __floats2half2_rn(y[2 * r], y[2 * r + 1]);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(y[2 * r] - hf.x, y[2 * r + 1] - hf.y);
            uint8_t* o = L.out + (long long)(s0 + rs[r]) * orow + ((long long)(ch >> 3) * F_out + pos) * 16 + (ch & 7) * 2;
            *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(&hh);
            *reinterpret_cast<uint32_t*>(o + (PC >> 3) * oplane) = *reinterpret_cast<const uint32_t*>(&ll);
        }
    }
}

// Bring the LSTM's input kernel into L1 ahead of its use (conv weights travel through shared memory, see the kernel).
__device__ __forceinline__ void fz_prefetch_weights(const FzLayer& L) {
    if (L.kind != FZ_LSTM) return;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(L.wk);
    const long long bytes = (long long)L.D * LSTM_GATES * 4;
    for (long long o = (long long)threadIdx.x * 128; o < bytes; o += (long long)FZ_THREADS * 128)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(base + o));
}

// ... and the layer's input rows of this CTA's streams (written by the previous layer a moment ago, or last step): one L2 round
// trip for the whole layer instead of one per k-step.
__device__ __forceinline__ void fz_prefetch_inputs(const FzLayer& L, int s0, int ns) {
    const uint8_t* src[4] = {L.a_cur, (L.kind == FZ_CONV && L.KT == 2) ? L.a_prev : nullptr, L.kind == FZ_CONV ? L.b_cur : nullptr,
                             (L.kind == FZ_CONV && L.KT == 2) ? L.b_prev : nullptr};
    const long long rb[2] = {L.kind == FZ_CONV ? (long long)L.F_in * L.Ca * 4 : (long long)L.D * 4, L.kind == FZ_CONV ? (long long)L.F_in * L.Cb * 4 : 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!src[i] || !rb[i >> 1]) continue;
        const uint8_t* base = src[i] + (long long)s0 * rb[i >> 1];
        const long long bytes = (long long)ns * rb[i >> 1];
        for (long long o = (long long)threadIdx.x * 128; o < bytes; o += (long long)FZ_THREADS * 128)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(base + o));
    }
}

// LSTM(21) + Dense of the CTA's streams (T = 1), eight streams at a time, all threads: the input projection is the only part with
// real work (84 x D MACs per stream) and is spread over (stream, gate) items; same arithmetic as lstm_block_kernel's stages.
__device__ __forceinline__ void fz_lstm_layer(const FzLayer& L, int s0, int ns, float (*xs)[256], float (*zs)[LSTM_GATES + 24]) {
    const int D = L.D, C = L.C, Fb = D / C, C8 = C >> 3, D8 = D >> 3;
    const int tid = threadIdx.x;
    for (int sb = 0; sb < ns; sb += 8) {
        const int nb = min(8, ns - sb);
        for (int it = tid; it < nb * D8; it += FZ_THREADS) {
            const int s = it / D8, k8 = it - s * D8;
            const int f = k8 / C8, c8 = k8 - f * C8;
            const uint8_t* row = L.a_cur + (long long)(s0 + sb + s) * D * 4;
            const uint4 hi = *reinterpret_cast<const uint4*>(row + ((long long)c8 * Fb + f) * 16);
            const uint4 lo = *reinterpret_cast<const uint4*>(row + ((long long)(C8 + c8) * Fb + f) * 16);
            float v[8];
            join8(hi, lo, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) xs[s][f * C + c8 * 8 + e] = v[e];
        }
        if (tid < nb * LSTM_UNITS) zs[tid / LSTM_UNITS][LSTM_GATES + tid % LSTM_UNITS] = L.h[(long long)(s0 + sb + tid / LSTM_UNITS) * LSTM_UNITS + tid % LSTM_UNITS];
        __syncthreads();
        // z[s][g] = bk[g] + sum_k x[s][k] Wk[k][g] + sum_j h[s][j] Wr[j][g]
        for (int it = tid; it < nb * LSTM_GATES; it += FZ_THREADS) {
            const int s = it / LSTM_GATES, g = it - s * LSTM_GATES;
            float z = __ldg(L.bk + g);
            const float* x = xs[s];
            const float* w = L.wk + g;
#pragma unroll 8
            for (int k = 0; k < D; ++k) z = fmaf(x[k], __ldg(w + (long long)k * LSTM_GATES), z);
            float zr = 0.0f;
            const float* hs = zs[s] + LSTM_GATES;
#pragma unroll
            for (int j = 0; j < LSTM_UNITS; ++j) zr = fmaf(hs[j], __ldg(L.wr + j * LSTM_GATES + g), zr);
            zs[s][g] = z + zr;
        }
        __syncthreads();
        if (tid < nb * LSTM_UNITS) {
            const int s = tid / LSTM_UNITS, u = tid % LSTM_UNITS;
            const long long so = (long long)(s0 + sb + s) * LSTM_UNITS + u;
            const float gi = fast_sigmoid(zs[s][u]), gf = fast_sigmoid(zs[s][LSTM_UNITS + u]), gc = fast_tanh(zs[s][2 * LSTM_UNITS + u]),
                        go = fast_sigmoid(zs[s][3 * LSTM_UNITS + u]);
            const float cn = fmaf(gf, L.c[so], gi * gc);
            const float hn = go * fast_tanh(cn);
            L.c[so] = cn;
            L.h[so] = hn;
            zs[s][LSTM_GATES + u] = hn;
        }
        __syncthreads();
        for (int it = tid; it < nb * D8; it += FZ_THREADS) {
            const int s = it / D8, k8 = it - s * D8;
            const float* hs = zs[s] + LSTM_GATES;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __ldg(L.bd + k8 * 8 + e);
#pragma unroll
            for (int j = 0; j < LSTM_UNITS; ++j) {
                const float hj = hs[j];
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(L.wd + (long long)j * D + k8 * 8));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(L.wd + (long long)j * D + k8 * 8) + 1);
                o[0] = fmaf(hj, w0.x, o[0]); o[1] = fmaf(hj, w0.y, o[1]); o[2] = fmaf(hj, w0.z, o[2]); o[3] = fmaf(hj, w0.w, o[3]);
                o[4] = fmaf(hj, w1.x, o[4]); o[5] = fmaf(hj, w1.y, o[5]); o[6] = fmaf(hj, w1.z, o[6]); o[7] = fmaf(hj, w1.w, o[7]);
            }
            const int f = k8 / C8, c8 = k8 - f * C8;
            sh16_store8(L.out + (long long)(s0 + sb + s) * D * 4, Fb, C, f, c8, o);
        }
        __syncthreads();
    }
}

// Weight chunks of the launch in execution order: (layer, first LN group, groups).  A chunk is at most FZ_WBUF bytes and is copied
// global -> shared by the copy engine (cp.async.bulk, completion on an mbarrier) while the previous chunk computes: two buffers.
struct FzChunk {
    int layer, px0, npx;
    int bytes;          // 0: no staged weights (LSTM)
    long long woff;     // byte offset of the chunk inside the layer's packed weights
};
__device__ __forceinline__ int fz_group_bytes(const FzLayer& L, int PC) { return L.KT * L.KF * ((L.Ca + L.Cb) >> 4) * (PC >> 3) * 512; }

__global__ void __launch_bounds__(FZ_THREADS, FZ_CTAS_PER_SM) fused_tail_kernel(const __grid_constant__ FzParams p) {
    extern __shared__ __align__(128) uint8_t fz_smem[];
    uint8_t* wbuf = fz_smem;                                                     // [2][FZ_WBUF]
    uint64_t* full = reinterpret_cast<uint64_t*>(fz_smem + 2 * FZ_WBUF);        // [2]
    __shared__ float xs_all[FZ_WARPS][256];
    __shared__ float zs_all[FZ_WARPS][LSTM_GATES + 24];
    __shared__ FzChunk chunks[2 * FZ_MAXL];
    __shared__ int nchunks_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s0 = blockIdx.x * p.G;
    const int ns = min(p.G, p.S - s0);
    if (ns <= 0) return;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
        int n = 0;
        for (int l = 0; l < p.nl; ++l) {
            const FzLayer& L = p.L[l];
            if (L.kind == FZ_LSTM) {
                chunks[n++] = FzChunk{l, 0, 0, 0, 0};
                continue;
            }
            const int PC = (L.epi == FZE_SHUF32) ? 32 : (L.epi == FZE_SHUF64) ? 64 : L.N;
            const int npx = L.N / PC, gb = fz_group_bytes(L, PC);
            const int per = max(1, min(npx, FZ_WBUF / gb));                      // LN groups per chunk
            for (int px = 0; px < npx; px += per) chunks[n++] = FzChunk{l, px, min(per, npx - px), gb * min(per, npx - px), (long long)gb * px};
        }
        nchunks_s = n;
    }
    __syncthreads();
    const int nchunks = nchunks_s;
    auto issue = [&](int c) {      // one thread: start the copy of chunk c into buffer c & 1
        const FzChunk& ch = chunks[c];
        if (!ch.bytes || (p.dbg & 4)) return;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.L[ch.layer].w) + ch.woff;
        uint8_t* dst = wbuf + (size_t)(c & 1) * FZ_WBUF;
        mbar_arrive_expect_tx(&full[c & 1], (uint32_t)ch.bytes);
        for (int o = 0; o < ch.bytes; o += 16384) bulk_g2s(dst + o, src + o, (uint32_t)min(16384, ch.bytes - o), &full[c & 1]);
    };
    if (threadIdx.x == 0) issue(0);
    uint32_t phase[2] = {0, 0};
    for (int c = 0; c < nchunks; ++c) {
        const FzChunk ch = chunks[c];
        const FzLayer& L = p.L[ch.layer];
        // buffer (c + 1) & 1 was last read by chunk c - 1, which every warp left at the barrier below
        if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1);
        if (ch.px0 == 0 && !(p.dbg & 8)) fz_prefetch_inputs(L, s0, ns);
        if (c + 1 < nchunks && !(p.dbg & 8)) fz_prefetch_weights(p.L[chunks[c + 1].layer]);
        if (L.kind == FZ_LSTM) {
            if (!(p.dbg & 2)) fz_lstm_layer(L, s0, ns, xs_all, zs_all);
        } else if (!(p.dbg & 1)) {
            if (!(p.dbg & 4)) mbar_wait(&full[c & 1], phase[c & 1]);
            phase[c & 1] ^= 1;
            const int mtiles = (ns * L.F_conv + 15) >> 4;
            const int PC = (L.epi == FZE_SHUF32) ? 32 : (L.epi == FZE_SHUF64) ? 64 : L.N;
            const int gb = fz_group_bytes(L, PC);
            const uint8_t* wsm = wbuf + (size_t)(c & 1) * FZ_WBUF;
            for (int w = warp; w < mtiles * ch.npx; w += FZ_WARPS) {
                const int mt = w / ch.npx, pl = w - mt * ch.npx;
                const uint4* wg = reinterpret_cast<const uint4*>(wsm + (size_t)pl * gb);
                if (PC == 32) fz_conv_tile<4>(L, wg, s0, ns, mt, ch.px0 + pl, lane);
                else fz_conv_tile<8>(L, wg, s0, ns, mt, ch.px0 + pl, lane);
            }
        }
        __syncthreads();          // the next chunk reads what this CTA wrote (block-scope ordering of global memory) and may refill this buffer
    }
}

}  // namespace nunet
