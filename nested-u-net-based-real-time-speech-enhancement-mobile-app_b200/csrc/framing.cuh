// Framing kernels: 512-point STFT analysis / synthesis around the network.
//   offline   tf.signal.stft / inverse_stft          models/proposed.py:285-291, :615-623
//   streaming shift-in, window, rfft / irfft, OLA    interpreter_proposed.py:200-213, :352-365
// One warp owns one frame.  The real 512-point transform is a 256-point COMPLEX transform of z[n] = x[2n] + i x[2n+1]
// plus an untangling step; the complex transform is a Stockham autosort FFT in three passes of radix 8, 8, 4 -- every lane
// holds eight points in registers per pass, the passes exchange them through a padded shared-memory buffer (two
// transposes, conflict-free), inputs come straight from HBM as coalesced 8-byte loads and the synthesis writes its last
// pass straight back.  (Round 1 ran a radix-2 512-point complex FFT in shared memory: 9 passes, 576 shared-memory
// accesses per lane against 56 / 32 here.)
// The phase is carried as the unit phasor X/|X| instead of angle(X): est*exp(j*angle(X)) == est*X/|X|.
#pragma once
#include "common.cuh"

namespace nunet {

constexpr int NFFT = 512;
constexpr int HOP = 256;
constexpr int NBINS = 257;
constexpr int FRAMES_PER_CTA = 4;
constexpr int FFT_N = NFFT / 2;               // points of the complex transform
constexpr int FFT_BUF = FFT_N + FFT_N / 8;    // one pad slot per eight points

struct FramingTables {
    const float2* tw;        // [256] exp(-2 pi i k / 512)
    const float* win;        // [512] analysis window
    const float* inv_win;    // [512] synthesis window  w / (w^2[n] + w^2[n +- 256])
};

struct FftSmem {
    float2 buf[FRAMES_PER_CTA][FFT_BUF];
    float2 tw[NFFT / 2];
};

__device__ __forceinline__ void load_tw(FftSmem& sm, const float2* tw) {
    for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x) sm.tw[i] = tw[i];
    __syncthreads();
}

__device__ __forceinline__ int fpad(int i) { return i + (i >> 3); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// exp(-2 pi i m / 512), 0 <= m < 512, from the half-circle table
__device__ __forceinline__ float2 tw512(const float2* tw_s, int m) {
    const float2 w = tw_s[m & 255];
    return (m & 256) ? make_float2(-w.x, -w.y) : w;
}
// two adjacent floats; the user's wav / hop pointers need not be 8-byte aligned
__device__ __forceinline__ float2 load2(const float* p, bool aligned) {
    return aligned ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(__ldg(p), __ldg(p + 1));
}
__device__ __forceinline__ void store2(float* p, float2 v, bool aligned) {
    if (aligned) {
        *reinterpret_cast<float2*>(p) = v;
    } else {
        p[0] = v.x;
        p[1] = v.y;
    }
}

// forward DFTs of 4 and 8 points, natural order in and out
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), s3 = csub(a1, a3);
    a0 = cadd(s0, s2);
    a2 = csub(s0, s2);
    a1 = make_float2(s1.x + s3.y, s1.y - s3.x);   // s1 - i s3
    a3 = make_float2(s1.x - s3.y, s1.y + s3.x);   // s1 + i s3
}
__device__ __forceinline__ void dft8(float2* a) {
    float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752f;
    const float2 t1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));      // o1 * exp(-i pi / 4)
    const float2 t2 = make_float2(o2.y, -o2.x);                               // o2 * (-i)
    const float2 t3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));     // o3 * exp(-3 i pi / 4)
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, t1); a[5] = csub(e1, t1);
    a[2] = cadd(e2, t2); a[6] = csub(e2, t2);
    a[3] = cadd(e3, t3); a[7] = csub(e3, t3);
}

// Passes 1 and 2 (radix 8, 8) of the 256-point forward transform.  v[r] = z[lane + 32 r] on entry; on return the buffer holds
// the input of pass 3 (Stockham order).
__device__ __forceinline__ void fft256_pass12(float2* v, float2* buf, const float2* tw_s, int lane) {
    dft8(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[fpad(8 * lane + r)] = v[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = buf[fpad(lane + 32 * r)];
    __syncwarp();
    const int k = lane & 7;
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tw512(tw_s, 8 * k * r));    // exp(-2 pi i k r / 64)
    dft8(v);
    const int jo = (lane - k) * 8 + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[fpad(jo + 8 * r)] = v[r];
    __syncwarp();
}
// Pass 3 (radix 4): lane works on j = lane and lane + 32; w[h][r] = Z[j + 64 r], natural order.
__device__ __forceinline__ void fft256_pass3(float2 (*w)[4], const float2* buf, const float2* tw_s, int lane) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
#pragma unroll
        for (int r = 0; r < 4; ++r) w[h][r] = buf[fpad(j + 64 * r)];
#pragma unroll
        for (int r = 1; r < 4; ++r) w[h][r] = cmul(w[h][r], tw512(tw_s, 2 * j * r));   // exp(-2 pi i j r / 256)
        dft4(w[h][0], w[h][1], w[h][2], w[h][3]);
    }
}

// Analysis: v[r] = windowed (x[2n], x[2n+1]) for n = lane + 32 r.  Writes |X[1..256]| and the unit phasor of X[0..256].
__device__ __forceinline__ void analyse_store(float2* v, float2* buf, const float2* tw_s, int lane, float* mag256, float2* ph257) {
    fft256_pass12(v, buf, tw_s, lane);
    float2 w[2][4];
    fft256_pass3(w, buf, tw_s, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 4; ++r) buf[fpad(lane + 32 * h + 64 * r)] = w[h][r];   // the slots this lane has just read
    __syncwarp();
    // X[k] = (Z[k] + conj Z[256-k]) / 2 + exp(-2 pi i k / 512) (Z[k] - conj Z[256-k]) / (2i)
    auto emit = [&](int k, float xr, float xi) {
        const float m = sqrtf(xr * xr + xi * xi);
        if (k >= 1) mag256[k - 1] = m;
        ph257[k] = (m > 0.0f) ? make_float2(xr / m, xi / m) : make_float2(1.0f, 0.0f);
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = lane + 32 * q;
        const float2 zk = buf[fpad(k)], zm = buf[fpad((FFT_N - k) & (FFT_N - 1))];
        const float2 xe = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 xo = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        const float2 t = cmul(tw_s[k], xo);
        emit(k, xe.x + t.x, xe.y + t.y);
        if (k == 0) emit(FFT_N, xe.x - t.x, xe.y - t.y);     // the Nyquist bin: exp(-i pi) = -1
    }
}

// Synthesis: Y[k] = est[k] * phasor[k] (k = 0..256, Hermitian), x = irfft(Y); returns w[h][r] = (x[2n], x[2n+1]) * 512 for
// n = lane + 32 h + 64 r.  est256 holds bins 1..256; the DC magnitude is 0 (tf.pad, models/proposed.py:617) or est[1]
// (np.pad 'edge', interpreter_proposed.py:352).  Imaginary parts of the DC and Nyquist bins are ignored like a C2R transform.
__device__ __forceinline__ void synthesise(float2 (*w)[4], float2* buf, const float2* tw_s, int lane, const float* est256,
                                           const float2* ph257, int dc_edge) {
    auto spec = [&](int k) {
        const float e = (k >= 1) ? est256[k - 1] : (dc_edge ? est256[0] : 0.0f);
        const float2 p = ph257[k];
        return make_float2(e * p.x, (k == 0 || k == FFT_N) ? 0.0f : e * p.y);
    };
    // Z[k] = Xe[k] + i Xo[k],  Xe = (Y[k] + conj Y[256-k]) / 2,  Xo = (Y[k] - conj Y[256-k]) / 2 * exp(+2 pi i k / 512); the
    // halves are left to the caller's 1/512.  The inverse transform is the conjugate of the forward transform of conj Z.
    float2 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = lane + 32 * q;
        const float2 yk = spec(k), ym = spec(FFT_N - k);
        const float2 xe = make_float2(yk.x + ym.x, yk.y - ym.y);
        const float2 d = make_float2(yk.x - ym.x, yk.y + ym.y);
        const float2 wk = tw_s[k];
        const float2 xo = cmul(d, make_float2(wk.x, -wk.y));
        v[q] = make_float2(xe.x - xo.y, -(xe.y + xo.x));
    }
    fft256_pass12(v, buf, tw_s, lane);
    fft256_pass3(w, buf, tw_s, lane);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 4; ++r) w[h][r].y = -w[h][r].y;
}

// ---- offline ---------------------------------------------------------------------------------------
// wav [B][n_samples] -> mag [B*T][256], phasor [B*T][257]
__global__ void __launch_bounds__(32 * FRAMES_PER_CTA) stft_kernel(const float* __restrict__ wav, FramingTables tb,
                                                                  float* __restrict__ mag, float2* __restrict__ ph,
                                                                  int B, int T, int n_samples) {
    __shared__ FftSmem sm;
    load_tw(sm, tb.tw);
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long frame = (long long)blockIdx.x * FRAMES_PER_CTA + wi;
    if (frame >= (long long)B * T) return;
    const int b = (int)(frame / T), t = (int)(frame - (long long)b * T);
    const float* src = wav + (size_t)b * n_samples + (size_t)t * HOP;
    const bool al = (reinterpret_cast<size_t>(src) & 7) == 0;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = lane + 32 * r;
        const float2 x = load2(src + 2 * n, al), wn = __ldg(reinterpret_cast<const float2*>(tb.win) + n);
        v[r] = make_float2(x.x * wn.x, x.y * wn.y);
    }
    analyse_store(v, sm.buf[wi], sm.tw, lane, mag + frame * 256, ph + frame * NBINS);
}

// est [B*T][257 (DC slot unused)] + phasor -> windowed time frames [B*T][512]
__global__ void __launch_bounds__(32 * FRAMES_PER_CTA) istft_frames_kernel(const float* __restrict__ est257,
                                                                          const float2* __restrict__ ph, FramingTables tb,
                                                                          float* __restrict__ frames, long long nframes) {
    __shared__ FftSmem sm;
    load_tw(sm, tb.tw);
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long frame = (long long)blockIdx.x * FRAMES_PER_CTA + wi;
    if (frame >= nframes) return;
    float2 w[2][4];
    synthesise(w, sm.buf[wi], sm.tw, lane, est257 + frame * NBINS + 1, ph + frame * NBINS, 0);
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = lane + 32 * h + 64 * r;
            const float2 iw = __ldg(reinterpret_cast<const float2*>(tb.inv_win) + n);
            reinterpret_cast<float2*>(frames + frame * NFFT)[n] =
                make_float2(w[h][r].x * (1.0f / NFFT) * iw.x, w[h][r].y * (1.0f / NFFT) * iw.y);
        }
}

// overlap-add of one (time chunk of a) batch: out[b][j] (j < span, relative to the chunk's first sample, rows n_out apart) =
// sum over the (at most two) frames of the chunk covering sample j.  add_head: the chunk continues its clips, so its first
// hop adds to what the previous chunk's last frame already left there (0 + a + b either way: bit-identical to one pass).
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ frames, float* __restrict__ out,
                                                         int B, int T, long long span, long long n_out /* per clip */, int add_head) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long total = (long long)B * span;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int b = (int)(idx / span);
        const int j = (int)(idx - (long long)b * span);
        const int t1 = j >> 8;   // frame starting at or before j
        float* o = out + (long long)b * n_out + j;
        float s = (add_head && j < HOP) ? *o : 0.0f;
        if (t1 - 1 >= 0 && t1 - 1 < T) s += frames[((size_t)b * T + (t1 - 1)) * NFFT + (j - ((t1 - 1) << 8))];
        if (t1 < T) s += frames[((size_t)b * T + t1) * NFFT + (j - (t1 << 8))];
        *o = s;
    }
}

// ---- streaming (one hop of S streams) --------------------------------------------------------------
// in_buf [S][512] is shifted left by 256 and the hop appended (interpreter_proposed.py:203-204).
__global__ void __launch_bounds__(32 * FRAMES_PER_CTA) stream_analysis_kernel(const float* __restrict__ hop,
                                                                             float* __restrict__ in_buf, FramingTables tb,
                                                                             float* __restrict__ mag, float2* __restrict__ ph,
                                                                             int S) {
    __shared__ FftSmem sm;
    load_tw(sm, tb.tw);
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * FRAMES_PER_CTA + wi;
    if (s >= S) return;
    float2* ib = reinterpret_cast<float2*>(in_buf + (size_t)s * NFFT);
    const float* hp = hop + (size_t)s * HOP;
    const bool al = (reinterpret_cast<size_t>(hp) & 7) == 0;
    // sample pair n = lane + 32 r: pairs 0..127 take the old second half (pairs n + 128, held by the same lane as r + 4), pairs
    // 128..255 the new hop; every read of the old buffer happens before this lane overwrites it
    float2 x[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) x[r] = ib[lane + 32 * r + 128];
#pragma unroll
    for (int r = 4; r < 8; ++r) x[r] = load2(hp + 2 * (lane + 32 * r - 128), al);
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = lane + 32 * r;
        ib[n] = x[r];
        const float2 wn = __ldg(reinterpret_cast<const float2*>(tb.win) + n);
        v[r] = make_float2(x[r].x * wn.x, x[r].y * wn.y);
    }
    analyse_store(v, sm.buf[wi], sm.tw, lane, mag + (size_t)s * 256, ph + (size_t)s * NBINS);
}

// est [S][256] -> irfft * inverse window, shift/add into out_buf [S][512], emit the first 256 samples
// (interpreter_proposed.py:352-365).
__global__ void __launch_bounds__(32 * FRAMES_PER_CTA) stream_synthesis_kernel(const float* __restrict__ est256,
                                                                              const float2* __restrict__ ph, FramingTables tb,
                                                                              float* __restrict__ out_buf,
                                                                              float* __restrict__ out_hop, int S, int dc_edge) {
    __shared__ FftSmem sm;
    load_tw(sm, tb.tw);
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * FRAMES_PER_CTA + wi;
    if (s >= S) return;
    float2 w[2][4];
    synthesise(w, sm.buf[wi], sm.tw, lane, est256 + (size_t)s * 256, ph + (size_t)s * NBINS, dc_edge);
    float2* ob = reinterpret_cast<float2*>(out_buf + (size_t)s * NFFT);
    float* oh = out_hop + (size_t)s * HOP;
    const bool al = (reinterpret_cast<size_t>(oh) & 7) == 0;
    // pair n = lane + 32 h + 64 r: pairs below 128 (r < 2) add the old second half (pair n + 128 = the same lane's r + 2)
    float2 old[2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 2; ++r) old[h][r] = ob[lane + 32 * h + 64 * r + 128];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = lane + 32 * h + 64 * r;
            const float2 iw = __ldg(reinterpret_cast<const float2*>(tb.inv_win) + n);
            const float2 blk = make_float2(w[h][r].x * (1.0f / NFFT) * iw.x, w[h][r].y * (1.0f / NFFT) * iw.y);
            const float2 o = (r < 2) ? old[h][r] : make_float2(0.0f, 0.0f);
            const float2 val = make_float2(o.x + blk.x, o.y + blk.y);
            ob[n] = val;
            if (r < 2) store2(oh + 2 * n, val, al);
        }
}

}  // namespace nunet
