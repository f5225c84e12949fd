// Framing kernels: 512-point STFT analysis / synthesis around the network.
//   offline   tf.signal.stft / inverse_stft          models/proposed.py:285-291, :615-623
//   streaming shift-in, window, rfft / irfft, OLA    interpreter_proposed.py:200-213, :352-365
// One warp owns one frame: a 512-point radix-2 complex FFT in shared memory (9 stages, 8 butterflies per
// lane per stage), inputs read coalesced from HBM, magnitudes and the unit phasor written coalesced.
// The phase is carried as the unit phasor X/|X| instead of angle(X): est*exp(j*angle(X)) == est*X/|X|.
#pragma once
#include "common.cuh"

namespace nunet {

constexpr int NFFT = 512;
constexpr int HOP = 256;
constexpr int NBINS = 257;
constexpr int FRAMES_PER_CTA = 4;

struct FramingTables {
    const float2* tw;        // [256] exp(-2 pi i k / 512)
    const float* win;        // [512] analysis window
    const float* inv_win;    // [512] synthesis window  w / (w^2[n] + w^2[n +- 256])
};

__device__ __forceinline__ int brev9(int i) { return (int)(__brev((unsigned)i) >> 23); }

// In-place forward DFT of (re, im) given in bit-reversed order; natural order out.
__device__ __forceinline__ void fft512_warp(float* re, float* im, const float2* tw_s, int lane) {
#pragma unroll 1
    for (int s = 1; s <= 9; ++s) {
        const int half = 1 << (s - 1);
        const int tstep = NFFT >> s;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = lane + 32 * q;
            const int pos = j & (half - 1);
            const int i0 = ((j >> (s - 1)) << s) + pos;
            const int i1 = i0 + half;
            const float2 w = tw_s[pos * tstep];
            const float xr = re[i1], xi = im[i1];
            const float tr = w.x * xr - w.y * xi;
            const float ti = w.x * xi + w.y * xr;
            const float ur = re[i0], ui = im[i0];
            re[i0] = ur + tr;
            im[i0] = ui + ti;
            re[i1] = ur - tr;
            im[i1] = ui - ti;
        }
        __syncwarp();
    }
}

struct FftSmem {
    float re[FRAMES_PER_CTA][NFFT];
    float im[FRAMES_PER_CTA][NFFT];
    float2 tw[NFFT / 2];
};

__device__ __forceinline__ void load_tw(FftSmem& sm, const float2* tw) {
    for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x) sm.tw[i] = tw[i];
    __syncthreads();
}

// Analysis of one frame already placed (windowed, bit-reversed) in sm.re/sm.im[wi]: writes |X[1..256]| and
// the unit phasor of X[0..256].
__device__ __forceinline__ void analyse_store(FftSmem& sm, int wi, int lane, float* mag256, float2* ph257) {
    fft512_warp(sm.re[wi], sm.im[wi], sm.tw, lane);
    for (int k = lane; k < NBINS; k += 32) {
        const float xr = sm.re[wi][k], xi = sm.im[wi][k];
        const float m = sqrtf(xr * xr + xi * xi);
        if (k >= 1) mag256[k - 1] = m;
        ph257[k] = (m > 0.0f) ? make_float2(xr / m, xi / m) : make_float2(1.0f, 0.0f);
    }
}

// Synthesis: Y[k] = est[k] * phasor[k] (k = 0..256, Hermitian-extended), x = irfft(Y) -> sm.re[wi][0..511] / 512.
// est256 holds bins 1..256; the DC magnitude is 0 (tf.pad, models/proposed.py:617) or est[1] (np.pad 'edge',
// interpreter_proposed.py:352).  Imaginary parts of the DC and Nyquist bins are ignored like a C2R transform.
__device__ __forceinline__ void synthesise(FftSmem& sm, int wi, int lane, const float* est256, const float2* ph257,
                                           int dc_edge) {
    for (int k = lane; k < NBINS; k += 32) {
        const float e = (k >= 1) ? est256[k - 1] : (dc_edge ? est256[0] : 0.0f);
        const float2 p = ph257[k];
        const float yr = e * p.x;
        const float yi = (k == 0 || k == NFFT / 2) ? 0.0f : e * p.y;
        // forward FFT of conj(Y) gives conj(N * ifft(Y)); we only need the real part.
        const int r0 = brev9(k);
        sm.re[wi][r0] = yr;
        sm.im[wi][r0] = -yi;
        if (k >= 1 && k < NFFT / 2) {
            const int r1 = brev9(NFFT - k);
            sm.re[wi][r1] = yr;
            sm.im[wi][r1] = yi;
        }
    }
    __syncwarp();
    fft512_warp(sm.re[wi], sm.im[wi], sm.tw, lane);
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
#pragma unroll 4
    for (int q = 0; q < NFFT / 32; ++q) {
        const int i = lane + 32 * q;
        const int r = brev9(i);
        sm.re[wi][r] = __ldg(src + i) * __ldg(tb.win + i);
        sm.im[wi][r] = 0.0f;
    }
    __syncwarp();
    analyse_store(sm, wi, lane, mag + frame * 256, ph + frame * NBINS);
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
    synthesise(sm, wi, lane, est257 + frame * NBINS + 1, ph + frame * NBINS, 0);
#pragma unroll 4
    for (int q = 0; q < NFFT / 32; ++q) {
        const int i = lane + 32 * q;
        frames[frame * NFFT + i] = sm.re[wi][i] * (1.0f / NFFT) * __ldg(tb.inv_win + i);
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
    float* ib = in_buf + (size_t)s * NFFT;
#pragma unroll 4
    for (int q = 0; q < NFFT / 32; ++q) {
        const int i = lane + 32 * q;
        const float v = (i < HOP) ? ib[i + HOP] : __ldg(hop + (size_t)s * HOP + (i - HOP));
        ib[i] = v;   // same lane read ib[i] (as ib[(i-256)+256]) eight iterations earlier
        const int r = brev9(i);
        sm.re[wi][r] = v * __ldg(tb.win + i);
        sm.im[wi][r] = 0.0f;
    }
    __syncwarp();
    analyse_store(sm, wi, lane, mag + (size_t)s * 256, ph + (size_t)s * NBINS);
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
    synthesise(sm, wi, lane, est256 + (size_t)s * 256, ph + (size_t)s * NBINS, dc_edge);
    float* ob = out_buf + (size_t)s * NFFT;
#pragma unroll 4
    for (int q = 0; q < NFFT / 32; ++q) {
        const int i = lane + 32 * q;
        const float blk = sm.re[wi][i] * (1.0f / NFFT) * __ldg(tb.inv_win + i);
        const float v = ((i < HOP) ? ob[i + HOP] : 0.0f) + blk;
        ob[i] = v;
        if (i < HOP) out_hop[(size_t)s * HOP + i] = v;
    }
}

}  // namespace nunet
