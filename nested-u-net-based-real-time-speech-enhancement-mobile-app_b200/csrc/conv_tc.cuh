// Fused causal convolution unit on the 5th-generation tensor cores (tcgen05, sm_100a), fp32-grade accuracy by
// 3xTF32 error compensation:  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  with  x_hi = rna_tf32(x),
// x_lo = rna_tf32(x - x_hi), fp32 accumulation in tensor memory (TMEM).  Same layer coverage and the same fused
// epilogues (bias, two-pass LayerNorm over channels, PReLU, both sub-pixel shuffles) as conv_simt.cuh.
//
// "Flat padded implicit GEMM".  For one layer the output positions of the whole batch are laid out on ONE flat
// axis  q = rho*P + x,  rho = clip*(T+padrow) + (t+padrow)  (one all-zero row in front of every clip when the unit
// has a time tap),  x in [0,P)  a frequency position including the zero-pad columns.  A tap (kt,kf) of the
// convolution is then a constant shift of q, so the A operand of every tap is the SAME shared-memory image read
// through a UMMA descriptor whose start address is shifted by `tap_off` rows.  That only works for a layout in
// which consecutive rows are a constant 16 bytes apart, hence the K-major *no-swizzle* canonical layout with
// SBO = 128 B (8-row core matrices back to back) and the 4-channel K chunks in separate planes (LBO = plane):
//      A_image[hi|lo][k_chunk][slot][4 floats]
// Stride-2 units read two images (even / odd input bins) so that their taps are unit-stride shifts as well.
// Pad rows / pad columns produce garbage accumulator rows that the epilogue simply does not store.
//
// CTA = 10 warps, persistent over 256-position tile pairs (two M=128 accumulators share every weight stage):
//   warps 0-3  epilogue   TMEM -> registers (one thread owns one output position, all channels: LayerNorm is
//                         thread-local), bias/LN/PReLU/shuffle, 128-bit global stores
//   warps 4-7  A loaders  global NHWC -> hi/lo split -> shared image (double buffered per 16-channel phase)
//   warp  8    MMA issuer one elected thread, tcgen05.mma kind::tf32, M=128, N=COUT, K=8
//   warp  9    W producer cp.async.bulk (TMA 1-D) of pre-split, pre-laid-out weight stages, 4-stage mbarrier ring
// TMEM: 2 (double buffer) x 2 (tile pair) x N columns <= 512.
#pragma once
#include "common.cuh"
#include "conv_simt.cuh"   // Epi enum

namespace nunet {

constexpr int TC_KCH = 16;        // input channels per phase (2 MMA K-steps of 8)
constexpr int TC_MT = 2;          // M=128 tiles per CTA iteration
constexpr int TC_WSTAGES = 4;
constexpr int TC_MAXTAPS = 6;
constexpr int TC_THREADS = 320;
constexpr int TC_MAXABUF = 3;
constexpr int TC_TBL_INTS = 2048;   // slot table: 2 tiles x (nimg*slots <= 1024)
constexpr int TC_STAGE_PITCH = 68;  // floats per staged output row (64 + 4: conflict-free 128-bit rows)
constexpr int TC_STAGE_BYTES = 4 * 32 * TC_STAGE_PITCH * 4 + 4 * 32 * 8;   // 4 epilogue warps: rows + row offsets

struct TcParams {
    const float* src0;   // [frames][F_in][C0]
    const float* src1;   // [frames][F_in][C1] or null
    const float* wpk;    // packed weights: [phase][tap][hi|lo][kchunk 4][N][4]
    const float* bias;
    const float* gamma;
    const float* beta;
    const float* alpha;
    float* out;
    int C0, C1;
    int B, T, F_in, F_conv;
    int P;               // flat positions per row
    int padrow;          // 1: a zero row precedes every clip (units with a time tap)
    int lead;            // image slot j <-> flat position q0 - lead + j
    int xlo;             // output valid iff xlo <= x < xlo + F_conv; bin f = x - xlo
    int nimg;
    int img_mul[2], img_add[2];   // input bin of image position x: fi = mul*x + add (zero outside [0,F_in))
    int ntaps;
    int tap_img[TC_MAXTAPS], tap_off[TC_MAXTAPS];   // slot = m + tap_off
    int nphase;          // (C0 + C1) / 16
    int slots;           // image length (positions) = 256 + max tap_off
    int plane_bytes;     // byte stride between K-chunk planes of the image, (plane_bytes/16) % 8 == 2
    int total_flat;      // B * (T + padrow) * P  (< 2^31, checked on the host)
    int ntiles;          // tile pairs
    int nabuf;           // A image buffers in the ring (2 or 3)
    int mt;              // M=128 tiles per CTA iteration (1 or 2); tile = mt*128 flat positions
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// K-major, no swizzle: rows 16 B apart (SBO = 128 B per 8 rows), K chunks `lbo_bytes` apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// LayerNorm (two-pass, like the reference's non-fused Keras path) + PReLU over v[0..CG) in place.
template <int CG, int STRIDE>
__device__ __forceinline__ void ln_prelu(float* v, const float* gamma, const float* beta, int goff, float alpha) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CG; ++i) s += v[i * STRIDE];
    const float mean = s * (1.0f / CG);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CG; ++i) {
        const float d = v[i * STRIDE] - mean;
        q = fmaf(d, d, q);
    }
    const float inv = rsqrtf(q * (1.0f / CG) + LN_EPS);
#pragma unroll
    for (int i = 0; i < CG; ++i) {
        const float sc = inv * __ldg(gamma + goff + i);
        const float y = fmaf(v[i * STRIDE], sc, __ldg(beta + goff + i) - mean * sc);
        v[i * STRIDE] = y >= 0.f ? y : alpha * y;
    }
}

template <int N, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // carve: barriers | tmem ptr | W ring | A buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* a_full = bars;                       // [3]
    uint64_t* a_empty = bars + 3;                  // [3]
    uint64_t* w_full = bars + 6;                   // [TC_WSTAGES]
    uint64_t* w_empty = bars + 6 + TC_WSTAGES;     // [TC_WSTAGES]
    uint64_t* acc_full = bars + 6 + 2 * TC_WSTAGES;   // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 24);
    constexpr uint32_t WSTAGE_BYTES = 2 * (TC_KCH / 4) * N * 16;   // hi + lo
    int* slot_tbl = reinterpret_cast<int*>(smem_raw + 256);        // [2][nimg*slots] input pixel index or -1
    float* stage_all = reinterpret_cast<float*>(smem_raw + 256 + TC_TBL_INTS * 4);
    long long* goff_all = reinterpret_cast<long long*>(smem_raw + 256 + TC_TBL_INTS * 4 + 4 * 32 * TC_STAGE_PITCH * 4);
    uint8_t* wring = smem_raw + 256 + TC_TBL_INTS * 4 + TC_STAGE_BYTES;
    const uint32_t abuf_bytes = (uint32_t)p.nimg * 2 * (TC_KCH / 4) * p.plane_bytes;
    uint8_t* abuf0 = wring + TC_WSTAGES * WSTAGE_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t ACC_COLS = 2 * TC_MT * N;          // 2 buffers x tile pair
    constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAXABUF; ++i) {
            mbar_init(&a_full[i], 128);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 128);
        }
        for (int i = 0; i < TC_WSTAGES; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    const int Tp = p.T + p.padrow;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x, +grid, ...

    if (warp < 4) {
        // ================================================================= epilogue
        // One thread owns one output position (TMEM lane) with all its channels, so LayerNorm is thread-local.
        // The finished row is staged in shared memory and written out by the whole warp, 16 lanes per 256-byte
        // row, so that every store instruction covers whole 128-byte lines (8x fewer LSU wavefronts than
        // per-thread row stores).
        const int row = threadIdx.x;   // TMEM lane == tile row
        const float alpha = (EPI == EPI_BIAS) ? 0.f : __ldg(p.alpha);
        float* stg = stage_all + warp * 32 * TC_STAGE_PITCH;
        long long* goff = goff_all + warp * 32;
        constexpr int ROWF = (N == 32) ? 32 : 64;          // floats per staged row
        constexpr int PASSES = (EPI == EPI_SHUF64) ? 2 : 1;
        constexpr int LPR = ROWF / 4;                      // lanes per row in the copy-out
        constexpr int RPI = 32 / LPR;                      // rows per store instruction
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int ab = it & 1;
            mbar_wait(&acc_full[ab], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < p.mt; ++mt) {
                const int q = tile * (p.mt * 128) + mt * 128 + row;
                const int rho = q / p.P;
                const int x = q - rho * p.P;
                const int b = rho / Tp;
                const int t = (rho - b * Tp) - p.padrow;
                const bool valid = (q < p.total_flat) && (t >= 0) && (x >= p.xlo) && (x < p.xlo + p.F_conv);
                const long long pix = ((long long)b * p.T + t) * p.F_conv + (x - p.xlo);   // conv-output pixel index
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((ab * p.mt + mt) * N);
#pragma unroll 1
                for (int pass = 0; pass < PASSES; ++pass) {
                    float v[ROWF];
#pragma unroll
                    for (int c = 0; c < ROWF; c += 32) tmem_ld32(taddr + pass * ROWF + c, v + c);
#pragma unroll
                    for (int c = 0; c < ROWF; ++c) v[c] += __ldg(p.bias + pass * ROWF + c);
                    if (EPI == EPI_LN) {
                        ln_prelu<ROWF, 1>(v, p.gamma, p.beta, 0, alpha);
                    } else if (EPI == EPI_SHUF32) {
                        // out[frame, 2f+j, i] = y[frame, f, 2i+j]: LN over the 32 channels of each parity j
                        ln_prelu<ROWF / 2, 2>(v, p.gamma, p.beta, 0, alpha);
                        ln_prelu<ROWF / 2, 2>(v + 1, p.gamma, p.beta, 0, alpha);
                    } else if (EPI == EPI_SHUF64) {
                        // half h = pass: out[frame, 2f+h, 32j+i] = y[frame, f, 64h+2i+j]; LN over the 64 channels of the half,
                        // gamma/beta indexed by the OUTPUT channel 32j+i
                        float sum = 0.f;
#pragma unroll
                        for (int c = 0; c < ROWF; ++c) sum += v[c];
                        const float mean = sum * (1.0f / ROWF);
                        float qq = 0.f;
#pragma unroll
                        for (int c = 0; c < ROWF; ++c) {
                            const float d = v[c] - mean;
                            qq = fmaf(d, d, qq);
                        }
                        const float inv = rsqrtf(qq * (1.0f / ROWF) + LN_EPS);
#pragma unroll
                        for (int c = 0; c < ROWF; ++c) {
                            const int oc = (ROWF / 2) * (c & 1) + (c >> 1);
                            const float sc = inv * __ldg(p.gamma + oc);
                            const float y = fmaf(v[c], sc, __ldg(p.beta + oc) - mean * sc);
                            v[c] = y >= 0.f ? y : alpha * y;
                        }
                    }
                    // stage the row in its final channel order
                    float4* srow = reinterpret_cast<float4*>(stg + lane * TC_STAGE_PITCH);
                    if (EPI == EPI_LN || EPI == EPI_BIAS) {
#pragma unroll
                        for (int k = 0; k < ROWF / 4; ++k) srow[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int i4 = 0; i4 < ROWF / 8; ++i4)
                                srow[j * (ROWF / 8) + i4] = make_float4(v[8 * i4 + j], v[8 * i4 + 2 + j], v[8 * i4 + 4 + j], v[8 * i4 + 6 + j]);
                    }
                    long long off = -1;
                    if (valid) off = (EPI == EPI_SHUF64) ? (pix * 2 + pass) * ROWF : pix * (long long)ROWF;
                    goff[lane] = off;
                    __syncwarp();
#pragma unroll 4
                    for (int i = 0; i < 32 / RPI; ++i) {
                        const int r = i * RPI + lane / LPR, ch4 = lane % LPR;
                        const long long o = goff[r];
                        if (o >= 0)
                            *reinterpret_cast<float4*>(p.out + o + ch4 * 4) =
                                *reinterpret_cast<const float4*>(stg + r * TC_STAGE_PITCH + ch4 * 4);
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);
        }
    } else if (warp < 8) {
        // ================================================================= A loaders
        // Raw fp32 rows are copied global -> shared with cp.async (zero-filled outside the image) straight into the
        // hi planes of a ring buffer, NB-1 phases ahead; each thread later converts exactly the items it copied
        // (hi = rna_tf32(x) in place, lo = rna_tf32(x - hi) into the lo planes).  The position -> input-pixel map is
        // the same for every phase of a tile, so it is computed once per tile into a small shared table.
        const int lt = threadIdx.x - 128;   // 0..127
        const int nslot = p.nimg * p.slots;
        const int items = nslot * (TC_KCH / 4);
        const int NB = p.nabuf, D = NB - 1;
        const int gtotal = my_tiles * p.nphase;
        const uint32_t lo_off = (TC_KCH / 4) * p.plane_bytes;

        auto build_table = [&](int it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int q0 = tile * (p.mt * 128) - p.lead;
            int* tb = slot_tbl + (it & 1) * (TC_TBL_INTS / 2);
            for (int e = lt; e < nslot; e += 128) {
                const int img = (e >= p.slots) ? 1 : 0;
                const int slot = e - img * p.slots;
                const int q = q0 + slot;
                int off = -1;
                if (q >= 0 && q < p.total_flat) {
                    const int rho = q / p.P;
                    const int x = q - rho * p.P;
                    const int b = rho / Tp;
                    const int t = (rho - b * Tp) - p.padrow;
                    const int fi = p.img_mul[img] * x + p.img_add[img];
                    if (t >= 0 && fi >= 0 && fi < p.F_in) off = (b * p.T + t) * p.F_in + fi;
                }
                tb[e] = off;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        };
        auto issue = [&](int g) {
            const int it = g / p.nphase, ph = g - it * p.nphase;
            if (ph == 0) build_table(it);
            const int buf = g % NB;
            if (g >= NB) mbar_wait(&a_empty[buf], ((g / NB) - 1) & 1);
            uint8_t* ab = abuf0 + (size_t)buf * abuf_bytes;
            const int* tb = slot_tbl + (it & 1) * (TC_TBL_INTS / 2);
            const int c0 = ph * TC_KCH;
            const float* src = (c0 < p.C0) ? p.src0 : p.src1;
            const int C = (c0 < p.C0) ? p.C0 : p.C1;
            const int cc = (c0 < p.C0) ? c0 : c0 - p.C0;
            for (int idx = lt; idx < items; idx += 128) {
                const int c4 = idx & 3, e = idx >> 2;
                const int img = (e >= p.slots) ? 1 : 0;
                const int slot = e - img * p.slots;
                const int off = tb[e];
                const float* gp = (off >= 0) ? (src + (size_t)off * C + cc + c4 * 4) : src;
                cp_async16(ab + ((img * 2) * (TC_KCH / 4) + c4) * p.plane_bytes + slot * 16, gp, off >= 0 ? 16 : 0);
            }
        };
        for (int g = 0; g < D && g < gtotal; ++g) {
            issue(g);
            cp_async_commit();
        }
        for (int g = 0; g < gtotal; ++g) {
            if (D == 2) cp_async_wait<1>(); else cp_async_wait<0>();   // phase g has landed (this thread's items)
            const int buf = g % NB;
            uint8_t* ab = abuf0 + (size_t)buf * abuf_bytes;
            for (int idx = lt; idx < items; idx += 128) {
                const int c4 = idx & 3, e = idx >> 2;
                const int img = (e >= p.slots) ? 1 : 0;
                const int slot = e - img * p.slots;
                uint8_t* d = ab + ((img * 2) * (TC_KCH / 4) + c4) * p.plane_bytes + slot * 16;
                const float4 v = *reinterpret_cast<const float4*>(d);
                uint4 hi, lo;
                hi.x = rna_tf32(v.x); hi.y = rna_tf32(v.y); hi.z = rna_tf32(v.z); hi.w = rna_tf32(v.w);
                lo.x = rna_tf32(v.x - __uint_as_float(hi.x));
                lo.y = rna_tf32(v.y - __uint_as_float(hi.y));
                lo.z = rna_tf32(v.z - __uint_as_float(hi.z));
                lo.w = rna_tf32(v.w - __uint_as_float(hi.w));
                *reinterpret_cast<uint4*>(d) = hi;
                *reinterpret_cast<uint4*>(d + lo_off) = lo;
            }
            fence_proxy_async();
            mbar_arrive(&a_full[buf]);
            if (g + D < gtotal) issue(g + D);   // its buffer was last read by the MMAs of phase g - 1
            cp_async_commit();
        }
    } else if (warp == 8) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            int gph = 0, gws = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int accb = it & 1;
                if (it >= 2) mbar_wait(&acc_empty[accb], ((it >> 1) - 1) & 1);
                tc_fence_after();
                for (int ph = 0; ph < p.nphase; ++ph, ++gph) {
                    const int buf = gph % p.nabuf;
                    mbar_wait(&a_full[buf], (gph / p.nabuf) & 1);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(abuf0 + (size_t)buf * abuf_bytes);
                    for (int tap = 0; tap < p.ntaps; ++tap, ++gws) {
                        const int ws = gws % TC_WSTAGES;
                        mbar_wait(&w_full[ws], (gws / TC_WSTAGES) & 1);
                        tc_fence_after();
                        const uint32_t w_base = smem_u32(wring + (size_t)ws * WSTAGE_BYTES);
                        const uint32_t a_img = a_base + (uint32_t)(p.tap_img[tap] * 2 * (TC_KCH / 4)) * p.plane_bytes +
                                               (uint32_t)p.tap_off[tap] * 16;
#pragma unroll 1
                        for (int mt = 0; mt < p.mt; ++mt) {
                            const uint32_t d = tmem_base + (uint32_t)((accb * p.mt + mt) * N);
#pragma unroll
                            for (int ks = 0; ks < TC_KCH / 8; ++ks) {
                                const uint32_t a_hi = a_img + (uint32_t)(2 * ks) * p.plane_bytes + mt * 128 * 16;
                                const uint32_t a_lo = a_hi + (TC_KCH / 4) * p.plane_bytes;
                                const uint32_t b_hi = w_base + (uint32_t)(2 * ks) * N * 16;
                                const uint32_t b_lo = b_hi + (TC_KCH / 4) * N * 16;
                                const uint64_t da_hi = make_desc(a_hi, p.plane_bytes), da_lo = make_desc(a_lo, p.plane_bytes);
                                const uint64_t db_hi = make_desc(b_hi, N * 16), db_lo = make_desc(b_lo, N * 16);
                                const uint32_t first = (ph == 0 && tap == 0 && ks == 0) ? 0u : 1u;
                                tc_mma_tf32(d, da_lo, db_hi, IDESC, first);   // small terms first
                                tc_mma_tf32(d, da_hi, db_lo, IDESC, 1u);
                                tc_mma_tf32(d, da_hi, db_hi, IDESC, 1u);
                            }
                        }
                        tc_commit(&w_empty[ws]);
                    }
                    tc_commit(&a_empty[buf]);
                }
                tc_commit(&acc_full[accb]);
            }
        }
    } else {
        // ================================================================= weight producer (1-D TMA)
        if (lane == 0) {
            int gws = 0;
            for (int it = 0; it < my_tiles; ++it) {
                for (int ph = 0; ph < p.nphase; ++ph) {
                    for (int tap = 0; tap < p.ntaps; ++tap, ++gws) {
                        const int ws = gws % TC_WSTAGES;
                        if (gws >= TC_WSTAGES) mbar_wait(&w_empty[ws], ((gws / TC_WSTAGES) - 1) & 1);
                        mbar_arrive_expect_tx(&w_full[ws], WSTAGE_BYTES);
                        bulk_g2s(wring + (size_t)ws * WSTAGE_BYTES,
                                 reinterpret_cast<const uint8_t*>(p.wpk) + ((size_t)ph * p.ntaps + tap) * WSTAGE_BYTES, WSTAGE_BYTES,
                                 &w_full[ws]);
                    }
                }
            }
        }
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// host: round-to-nearest (ties away) fp32 -> tf32, the same rounding as cvt.rna.tf32.f32
__host__ inline float host_rna_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

}  // namespace nunet
