// "int8-hybrid" arithmetic of the DEPLOYED graph.  The reference ships its one-frame model as a
// dynamic-range-quantised .tflite (converter_proposed.py:901 Optimize.DEFAULT): 170 of 172 CONV_2D and 35 of 39
// FULLY_CONNECTED operators hold int8 weights, and the TFLite runtime executes them with its HYBRID kernels -- per call the
// float input tensor is quantised to int8 (asymmetric, one scale / zero point per batch item), products are accumulated in
// int32 and rescaled by input scale x per-channel weight scale (tensorflow/lite/kernels/conv.cc EvalHybridPerChannel,
// fully_connected.cc EvalHybrid, internal/reference/portable_tensor_utils.cc AsymmetricQuantizeFloats).  These kernels
// reproduce that arithmetic for S independent streams (each stream is its own "batch 1"), following the graph operator by
// operator -- no algebraic fusion across a quantisation point: the transpose convolution stays a float op on dequantised
// weights and is NOT composed with the following 1x1 convolution, whose input is quantised in between.
// Layout: fp32 NHWC per stream, [S][F][C]; history rows are the other step parity's buffers like in the float plan.
// The integer paths are exact (int32 accumulation of int8 x int8); float steps are evaluated in the operator order of the graph.
#pragma once
#include "common.cuh"

namespace nunet {

struct HqParams {      // per-stream quantisation of the tensor being convolved
    float scale;       // float32 input scale
    int zp;            // zero point (int8 range)
};

__device__ __forceinline__ float round_away(float x) { return copysignf(floorf(fabsf(x) + 0.5f), x); }
__device__ __forceinline__ double round_away_d(double x) { return copysign(floor(fabs(x) + 0.5), x); }

// AsymmetricQuantizeFloats: scale / zero point from (min, max) in double, values quantised in float (no FMA contraction)
__device__ __forceinline__ HqParams hq_asym_params(float mn, float mx, float* inv_out) {
    const double rmin = fmin(0.0, (double)mn), rmax = fmax(0.0, (double)mx);
    HqParams p;
    if (rmin == rmax) {
        p.scale = 1.0f;
        p.zp = 0;
        *inv_out = 0.0f;        // every value quantises to 0
        return p;
    }
    const double scale = (rmax - rmin) / 255.0;
    const double zp_min = -128.0 - rmin / scale, zp_max = 127.0 - rmax / scale;
    const double err_min = 128.0 + fabs(rmin / scale), err_max = 127.0 + fabs(rmax / scale);
    double zp = err_min < err_max ? zp_min : zp_max;
    zp = (zp <= -128.0) ? -128.0 : (zp >= 127.0 ? 127.0 : round_away_d(zp));
    p.scale = (float)scale;
    p.zp = (int)zp;
    *inv_out = __fdiv_rn(1.0f, p.scale);
    return p;
}
__device__ __forceinline__ int hq_asym_q(float x, float inv, int zp) {
    const float v = round_away(__fadd_rn((float)zp, __fmul_rn(x, inv)));
    return (int)fminf(fmaxf(v, -128.0f), 127.0f);
}

// ---- quantise the input tensor of a hybrid CONV_2D: rows x F x (CA + CB), rows = [prev | cur] for causal convs -----------
// One CTA per stream.  out_q: int8 [S][rows][F][Ctot]; qp [S].
__global__ void __launch_bounds__(256) hq_quantize_kernel(const float* __restrict__ a_prev, const float* __restrict__ a_cur,
                                                         const float* __restrict__ b_prev, const float* __restrict__ b_cur, int F, int CA,
                                                         int CB, int rows, int8_t* __restrict__ out_q, HqParams* __restrict__ qp) {
    __shared__ float red_mn[8], red_mx[8];
    __shared__ HqParams sp;
    __shared__ float s_inv;
    const long long s = blockIdx.x;
    const int Ct = CA + CB;
    const int n = rows * F * Ct;
    auto at = [&](int i) -> float {
        const int c = i % Ct, f = (i / Ct) % F, r = i / (Ct * F);
        const bool cur = (rows == 1) || r == 1;
        if (c < CA) return (cur ? a_cur : a_prev)[(s * F + f) * CA + c];
        return (cur ? b_cur : b_prev)[(s * F + f) * CB + (c - CA)];
    };
    float mn = 0.0f, mx = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = at(i);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, m));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    }
    if ((threadIdx.x & 31) == 0) {
        red_mn[threadIdx.x >> 5] = mn;
        red_mx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            mn = fminf(mn, red_mn[w]);
            mx = fmaxf(mx, red_mx[w]);
        }
        float inv;
        sp = hq_asym_params(mn, mx, &inv);
        s_inv = inv;
        qp[s] = sp;
    }
    __syncthreads();
    const float inv = s_inv;
    const int zp = sp.zp;
    int8_t* o = out_q + s * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = (int8_t)hq_asym_q(at(i), inv, zp);
}

// ---- hybrid CONV_2D (EvalHybridPerChannel) + the float operators that follow it in the graph ------------------------------
enum HqEpi { HQ_BIAS = 0, HQ_LN = 1, HQ_SHUF32 = 2, HQ_SHUF64 = 3 };
struct HqConv {
    const int8_t* q;          // [S][rows][F_in][Ct]
    const HqParams* qp;       // [S]
    const int* w;             // int8 x 4 words: [tap][Ct/4][Cout]
    const int* wtap;          // [tap][Cout] sum of the int8 weights of a tap (to subtract zero_point * sum over the valid taps)
    const float* wscale;      // [Cout]
    const float* bias;        // [Cout]
    const float* gamma;       // LayerNorm over the channels of an OUTPUT pixel
    const float* beta;
    const float* alpha;       // PReLU (one shared slope)
    float* out;               // [S][F_out][C_out]
    int F_in, Ct, Cout, KT, KF, padl, stride, F_conv, epi;
};
// CTA = (stream, tile of 128 / Cout conv positions); thread = (position, conv channel)
__global__ void __launch_bounds__(128) hq_conv_kernel(const HqConv p) {
    extern __shared__ float tile[];     // [PT][Cout]
    const long long s = blockIdx.x;
    const int PT = 128 / p.Cout;
    const int c = threadIdx.x % p.Cout, pl = threadIdx.x / p.Cout;
    const int f = blockIdx.y * PT + pl;
    const HqParams qp = p.qp[s];
    const int C4 = p.Ct >> 2;
    float y = 0.0f;
    if (f < p.F_conv) {
        int acc = 0, wsum = 0;
        for (int kt = 0; kt < p.KT; ++kt)
            for (int kf = 0; kf < p.KF; ++kf) {
                const int fi = f * p.stride - p.padl + kf;
                if (fi < 0 || fi >= p.F_in) continue;              // zero padding = the zero point: contributes nothing
                const int tap = kt * p.KF + kf;
                const int* a = reinterpret_cast<const int*>(p.q + ((s * p.KT + kt) * p.F_in + fi) * (long long)p.Ct);
                const int* w = p.w + (long long)tap * C4 * p.Cout + c;
#pragma unroll 4
                for (int k = 0; k < C4; ++k) acc = __dp4a(a[k], w[(long long)k * p.Cout], acc);
                wsum += p.wtap[tap * p.Cout + c];
            }
        const int r = acc - qp.zp * wsum;
        y = __fadd_rn(__fmul_rn((float)r, __fmul_rn(qp.scale, p.wscale[c])), p.bias[c]);
    }
    if (p.epi == HQ_BIAS) {
        if (f < p.F_conv) p.out[(s * p.F_conv + f) * p.Cout + c] = y;
        return;
    }
    tile[pl * p.Cout + c] = y;
    __syncthreads();
    if (f >= p.F_conv) return;
    // LayerNorm over the channels of one output pixel, then PReLU.  Groups: all conv channels (HQ_LN); the two interleaved
    // halves {2i + j} (HQ_SHUF32: pixel 2f + j, channel i); the two contiguous halves (HQ_SHUF64: pixel 2f + h, conv channel
    // 64h + 2i + j -> channel 32j + i)   [models/proposed.py:227-237, SURVEY 3A.2]
    const float* row = tile + pl * p.Cout;
    int G, base, step, opix, och;
    if (p.epi == HQ_LN) { G = p.Cout; base = 0; step = 1; opix = f; och = c; }
    else if (p.epi == HQ_SHUF32) { G = p.Cout / 2; base = c & 1; step = 2; opix = 2 * f + (c & 1); och = c >> 1; }
    else { const int h = c / 64, m = c % 64; G = 64; base = 64 * h; step = 1; opix = 2 * f + h; och = 32 * (m & 1) + (m >> 1); }
    float sum = 0.0f;
    for (int i = 0; i < G; ++i) sum += row[base + i * step];
    const float mean = sum / (float)G;
    float sq = 0.0f;
    for (int i = 0; i < G; ++i) {
        const float d = row[base + i * step] - mean;
        sq += d * d;
    }
    const float inv = rsqrtf(sq / (float)G + LN_EPS) * p.gamma[och];
    float o = y * inv + (p.beta[och] - mean * inv);
    o = o >= 0.0f ? o : p.alpha[0] * o;
    const int Fo = (p.epi == HQ_LN) ? p.F_conv : 2 * p.F_conv, Co = (p.epi == HQ_LN) ? p.Cout : p.Cout / 2;
    p.out[(s * Fo + opix) * Co + och] = o;
}

// ---- row-quantised hybrid FULLY_CONNECTED helpers (asymmetric_quantize_inputs = true in the shipped graph) ---------------
// quantise x[0..n) (shared memory) of one row; all threads of the CTA call it; q_out in shared memory
__device__ __forceinline__ HqParams hq_quantize_row(const float* x, int n, int* q_out, float* red /*[2]*/) {
    __syncthreads();
    if (threadIdx.x == 0) {
        float mn = 0.0f, mx = 0.0f;
        for (int i = 0; i < n; ++i) {
            mn = fminf(mn, x[i]);
            mx = fmaxf(mx, x[i]);
        }
        float inv;
        const HqParams p = hq_asym_params(mn, mx, &inv);
        red[0] = p.scale;
        red[1] = inv;
        reinterpret_cast<int*>(red)[2] = p.zp;
    }
    __syncthreads();
    HqParams p;
    p.scale = red[0];
    p.zp = reinterpret_cast<int*>(red)[2];
    for (int i = threadIdx.x; i < n; i += blockDim.x) q_out[i] = hq_asym_q(x[i], red[1], p.zp) - p.zp;   // q - zero_point
    __syncthreads();
    return p;
}

// LSTM(21) + Dense of one stream-frame with hybrid FULLY_CONNECTED operators: z = FC(x, W) + FC(h, U) + b (gate order
// i, f, c, o), c' = sig(f) c + sig(i) tanh(g), h' = sig(o) tanh(c'), y = FC(h', Wd) + bd.  Wd is int8 (dense_q != null) or,
// for the four 32-wide bottlenecks whose kernel has < 1024 elements, float.
struct HqLstm {
    const float* x;           // [S][D]
    float* y;                 // [S][D]
    float *h, *c;             // [S][21]
    const int8_t* wk;         // [84][D]
    const int8_t* wr;         // [84][21]
    const int8_t* wd;         // [D][21] or null
    const float* wd_f;        // [21][D] float (Keras layout) when wd == null
    const float* bk;          // [84]
    const float* bd;          // [D]
    float sk, sr, sd;         // per-tensor weight scales
    int D;
};
__global__ void __launch_bounds__(128) hq_lstm_kernel(const HqLstm p) {
    __shared__ float xs[256];
    __shared__ int xq[256];
    __shared__ float hs[LSTM_UNITS + 3];
    __shared__ int hq[LSTM_UNITS + 3];
    __shared__ float z[LSTM_GATES];
    __shared__ float red[4];
    const long long s = blockIdx.x;
    const int D = p.D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) xs[i] = p.x[s * D + i];
    if (threadIdx.x < LSTM_UNITS) hs[threadIdx.x] = p.h[s * LSTM_UNITS + threadIdx.x];
    const HqParams px = hq_quantize_row(xs, D, xq, red);
    float zx = 0.0f;
    if (threadIdx.x < LSTM_GATES) {
        int acc = 0;
        const int8_t* w = p.wk + (long long)threadIdx.x * D;
        for (int k = 0; k < D; ++k) acc += xq[k] * (int)w[k];
        zx = __fmul_rn((float)acc, __fmul_rn(px.scale, p.sk));
    }
    const HqParams ph = hq_quantize_row(hs, LSTM_UNITS, hq, red);
    if (threadIdx.x < LSTM_GATES) {
        int acc = 0;
        const int8_t* w = p.wr + threadIdx.x * LSTM_UNITS;
        for (int k = 0; k < LSTM_UNITS; ++k) acc += hq[k] * (int)w[k];
        const float zh = __fmul_rn((float)acc, __fmul_rn(ph.scale, p.sr));
        z[threadIdx.x] = __fadd_rn(__fadd_rn(zx, zh), p.bk[threadIdx.x]);
    }
    __syncthreads();
    if (threadIdx.x < LSTM_UNITS) {
        const int j = threadIdx.x;
        const float gi = sigmoidf_(z[j]), gf = sigmoidf_(z[LSTM_UNITS + j]), gc = tanhf(z[2 * LSTM_UNITS + j]), go = sigmoidf_(z[3 * LSTM_UNITS + j]);
        const float cn = __fadd_rn(__fmul_rn(gf, p.c[s * LSTM_UNITS + j]), __fmul_rn(gi, gc));
        const float hn = __fmul_rn(go, tanhf(cn));
        p.c[s * LSTM_UNITS + j] = cn;
        p.h[s * LSTM_UNITS + j] = hn;
        hs[j] = hn;
    }
    if (p.wd) {
        const HqParams pd = hq_quantize_row(hs, LSTM_UNITS, hq, red);
        for (int n = threadIdx.x; n < D; n += blockDim.x) {
            int acc = 0;
            const int8_t* w = p.wd + n * LSTM_UNITS;
            for (int k = 0; k < LSTM_UNITS; ++k) acc += hq[k] * (int)w[k];
            p.y[s * D + n] = __fadd_rn(__fmul_rn((float)acc, __fmul_rn(pd.scale, p.sd)), p.bd[n]);
        }
    } else {
        __syncthreads();
        for (int n = threadIdx.x; n < D; n += blockDim.x) {
            float acc = 0.0f;
            for (int k = 0; k < LSTM_UNITS; ++k) acc = fmaf(hs[k], p.wd_f[k * D + n], acc);
            p.y[s * D + n] = acc + p.bd[n];
        }
    }
}

// ---- CTFA of the one-frame graph (models/proposed.py:162 ctfa_rt) with hybrid 1x1 convolutions -----------------------------
// TA = sig(W2 relu(W1 mean_f x)), FA = sig(V2 relu(V1 (TA / 32))) -- the four convolutions are hybrid CONV_2D operators with
// per-channel weight scales, each quantising its whole input tensor (FA's input is TA broadcast over F, so one row gives the
// tensor's min / max) -- then out = x (FA TA) + residual.  One CTA (64 threads) per stream.
struct HqMlp {
    const int8_t *k0, *k1;    // [16][64], [64][16]  ([Cout][Cin])
    const float *s0, *s1;     // per-channel scales
    const float *b0, *b1;
};
__device__ __forceinline__ void hq_mlp(const HqMlp& m, float* v /*[64] in smem, overwritten with sigmoid output*/, float* hbuf /*[16]*/,
                                       int* qbuf /*[64]*/, float* red) {
    const int c = threadIdx.x;
    const HqParams p0 = hq_quantize_row(v, 64, qbuf, red);
    if (c < 16) {
        int acc = 0;
        for (int k = 0; k < 64; ++k) acc += qbuf[k] * (int)m.k0[c * 64 + k];
        const float y = __fadd_rn(__fmul_rn((float)acc, __fmul_rn(p0.scale, m.s0[c])), m.b0[c]);
        hbuf[c] = fmaxf(y, 0.0f);
    }
    const HqParams p1 = hq_quantize_row(hbuf, 16, qbuf, red);
    int acc = 0;
    for (int k = 0; k < 16; ++k) acc += qbuf[k] * (int)m.k1[c * 16 + k];
    const float y = __fadd_rn(__fmul_rn((float)acc, __fmul_rn(p1.scale, m.s1[c])), m.b1[c]);
    __syncthreads();
    v[c] = sigmoidf_(y);
    __syncthreads();
}
__global__ void __launch_bounds__(64) hq_ctfa_kernel(const float* __restrict__ x, const float* __restrict__ res, HqMlp ta, HqMlp fa,
                                                    float* __restrict__ out, int F) {
    __shared__ float v[64], tav[64], hbuf[16], red[4];
    __shared__ int qbuf[64];
    const long long s = blockIdx.x;
    const int c = threadIdx.x;
    const float* xs = x + s * F * 64;
    float sum = 0.0f;
    for (int f = 0; f < F; ++f) sum += xs[f * 64 + c];
    v[c] = sum / (float)F;
    hq_mlp(ta, v, hbuf, qbuf, red);
    tav[c] = v[c];
    v[c] = v[c] * (1.0f / CTFA_WINDOW);          // 31 zero rows + this frame, average-pooled over 32
    hq_mlp(fa, v, hbuf, qbuf, red);
    const float g = __fmul_rn(v[c], tav[c]);     // TFA = FA * TA
    const float* rs = res + s * F * 64;
    float* o = out + s * F * 64;
    for (int f = 0; f < F; ++f) o[f * 64 + c] = __fadd_rn(__fmul_rn(xs[f * 64 + c], g), rs[f * 64 + c]);
}

// ---- up_sampling: TRANSPOSE_CONV (1,3) stride (1,2) 'same' on DEQUANTISED weights (a float operator in the graph) ----------
// u[j][co] = b[co] + sum_ci sum_{(i,k): 2i + k = j + 0 (crop 0 left)} x[i][ci] W[k][co][ci]; even j = 2i: taps (i, 0), (i-1, 2);
// odd j = 2i + 1: tap (i, 1).  x = cat[a, b] on channels (64 + 64).  w: [k][ci][co] float.  grid (S, F_out), 128 threads = co.
__global__ void __launch_bounds__(128) hq_upsample_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int F_in) {
    __shared__ float xs[2][128];
    const long long s = blockIdx.x;
    const int j = blockIdx.y, co = threadIdx.x;
    const int i = j >> 1;
    const bool odd = j & 1;
    const int src_i[2] = {i, i - 1};
    for (int r = 0; r < 2; ++r) {
        const int ii = src_i[r];
        float v = 0.0f;
        if (ii >= 0 && ii < F_in && !(odd && r == 1)) v = (co < 64) ? a[(s * F_in + ii) * 64 + co] : b[(s * F_in + ii) * 64 + (co - 64)];
        xs[r][co] = v;
    }
    __syncthreads();
    float acc = 0.0f;
    const float* w0 = w + (size_t)(odd ? 1 : 0) * 128 * 128;
    for (int ci = 0; ci < 128; ++ci) acc = fmaf(xs[0][ci], w0[ci * 128 + co], acc);
    if (!odd && i - 1 >= 0) {
        const float* w2 = w + (size_t)2 * 128 * 128;
        for (int ci = 0; ci < 128; ++ci) acc = fmaf(xs[1][ci], w2[ci * 128 + co], acc);
    }
    out[(s * 2 * F_in + j) * 128 + co] = acc + bias[co];
}

}  // namespace nunet
