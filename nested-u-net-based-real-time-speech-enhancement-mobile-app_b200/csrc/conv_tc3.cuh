// Fused causal convolution unit on tcgen05 (sm_100a) over "split-half" activations -- the offline hot path.
//
// Numerics: every operand x is carried as two IEEE halves  x ~= hi + lo,  hi = rn_f16(x), lo = rn_f16(x - hi)
// (22 significant bits while |x| < 65504; the products hi*hi, hi*lo, lo*hi are exact in fp32), and a
// contraction is the three products  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  accumulated in fp32 in tensor memory,
// issued as TWO kind::f16 MMAs per K step:  a_hi x [b_hi | b_lo]  (2N accumulator columns) and  a_lo x b_hi
// (into the first N columns); the epilogue adds the two column blocks.  Sharing the a_hi read between two products
// matters because at M = 128 the shared-memory read of A, not the tensor pipe, bounds an N <= 64 MMA.
// That is the accuracy of a 3xTF32 scheme (the first kernel of this repo, since removed) at twice the tensor rate and half the shared-memory
// operand traffic, and -- the point of the format -- the split is done ONCE by the producer's epilogue, so the
// consumer's loaders are pure 16-byte cp.async copies (no conversion pass: the old kernel was bound by it).
//
// HBM format "sh16" of an activation tensor [frame][F][C] (same footprint as fp32, so the arena layout of the plan
// is unchanged): every frame row is stored PLANAR, exactly like the shared-memory operand image,
//      row[hi|lo][chunk C/8][f][8 halves]          (a "plane" = F x 16 bytes, one 8-channel chunk of one part)
// so that (a) consecutive bins of a plane are consecutive 16-byte units both in HBM and in the image: a loader warp
// copies 512 contiguous bytes per cp.async instruction, and (b) the epilogue, where lane = output position, writes
// 512 contiguous bytes per store instruction straight from registers -- no staging through shared memory.
// A tensor may keep its bins in "even/odd" order inside every plane, [even bins | odd bins] (flag eo): that is the
// order in which the two CTAs of a 128-channel unit produce them (each writes one contiguous half) and the order in
// which the stride-2 units consume them (their even / odd images become contiguous copies).
// Weights are pre-split on the host and scaled by a power of two per layer so that their lo parts stay normal
// halves (undone in the epilogue).
//
// Geometry: a "flat padded implicit GEMM" -- output positions of the whole batch on one flat
// axis q = rho*P + x with zero pad rows / columns, every tap a constant shift of q, so all taps read the SAME
// shared-memory image through K-major no-swizzle UMMA descriptors (rows 16 B apart, 8-channel K chunks in planes):
//      buffer[hi|lo][chunk 2][img][slot][8 halves],   one image buffer = 16 input channels = one K=16 MMA step.
// Stride-2 units keep two images (even / odd input bins) back to back inside every plane.
//
// Weights stay RESIDENT in shared memory for the whole (persistent) CTA: units with 128 conv channels (last spconv
// of a block, up_sampling o inconv) are two 64-column problems (their LayerNorm runs over each half), so every CTA
// holds <= 96 KB of weights.  The host permutes the weight columns into output-channel order, so the sub-pixel
// shuffles cost nothing here.  Those units run as CTA PAIRS (template PAIR, 2-CTA clusters): CTA r owns one
// 128-position tile and the weights of half r, and the leader issues tcgen05.mma.cta_group::2 (M = 256) over both
// CTAs' images and weight halves -- see Tc3Params::pair.  (Fallback NUNET_TC3_PAIR=0: CTA 2c / 2c+1 take half 0 / 1
// of the same tiles independently.)
//
// CTA = 16 warps (13 working + 3 idle ones that only balance the register file, see T3_WARPS), persistent over tiles of
// mt*128 positions:
//   warps 0-7   epilogue  two groups of four warps (two warps share every SM sub-partition and hide each other's
//                         latencies): TMEM -> registers (one thread = one position, all its channels: LayerNorm is
//                         thread-local) -> bias / two-pass LN / PReLU -> hi/lo split -> 16/32-byte stores, coalesced
//                         across the warp by the planar layout
//   warps 8-11  loaders   one warp per plane (hi|lo x chunk) into a ring of image buffers.  Default (tma = 2): ONE
//                         cp.async.bulk.tensor box per tile image -- whole frame rows, pads and the causal time pad
//                         zero-filled by the copy engine (Tc3Params::tm_*); tiles that straddle two clips, streaming
//                         (history row from the other parity's buffer) and NUNET_TC3_TMA=0 use 16-byte cp.async from a
//                         per-tile slot table; NUNET_TC3_TMA=1 uses 1-D bulk copies per frame-row segment.  Units with many
//                         short phases let the warps take the ring buffers in turn, four lanes issuing the four planes (ld_rr)
//   warp  12    MMA       one elected thread: tcgen05.mma kind::f16, M=128 (256 for pairs), K=16; accumulators
//                         double-buffered in TMEM, buffers freed by tcgen05.commit
// Tiles are consecutive runs of 128 flat positions, or, for units with a multiple of 128 output bins, 128-bin pieces
// of one frame row (Tc3Params::row_tpr): no pad position is computed and box images need the fewest rows.
#pragma once
#include <cuda.h>        // CUtensorMap
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_simt.cuh"    // Epi enum
#include "tcgen05_ptx.cuh"  // mbarrier / tcgen05 PTX wrappers, make_desc, tmem_ld32

namespace nunet {

constexpr int T3_KCH = 16;          // input channels per image buffer (one K=16 MMA step)
constexpr int T3_MT = 2;            // M=128 tiles per CTA iteration
constexpr int T3_MAXTAPS = 6;
constexpr int T3_EPI_WARPS = 8;
constexpr int T3_LD_WARPS = 4;
constexpr int T3_LD_THREADS = 32 * T3_LD_WARPS;
// 16 warps = 4 per SM sub-partition: {epilogue, epilogue, loader, MMA issuer | idle}.  The block is launched with 128 registers
// per thread (the whole register file); right after the prologue every role resizes its allocation with setmaxnreg so that
// the epilogue threads -- one output row each, N accumulator values + N + 32 raw tcgen05.ld words live at the peak -- get
// 192 registers and never spill:  2 x 192 + 80 (loader) + 48 (MMA issuer; the three idle warps keep 24) = 512 = 16 K / 32.
constexpr int T3_WARPS = 16;
constexpr int T3_THREADS = 32 * T3_WARPS;
constexpr int T3_REGS_EPI = 192, T3_REGS_LD = 80, T3_REGS_MMA = 48, T3_REGS_IDLE = 24;
constexpr int T3_MAXNB = 8;         // image ring depth
constexpr int T3_TBL = 1024;        // slot table entries (nimg*slots <= 1024)
constexpr int T3_PREV_FLAG = 1 << 30;   // slot-table entry: read the history tensor (streaming) instead of the source
constexpr int T3_PEER_OFF = 256;    // pair mode: barriers the partner CTA arrives on (peer_full[T3_MAXNB], peer_acc_empty[2])
constexpr int T3_PAR_OFF = 512;     // bias[128] | gamma[64] | beta[64] staged as floats
constexpr int T3_TBL_OFF = 2048;
constexpr int T3_FIXED_BYTES = T3_TBL_OFF + T3_TBL * 4;

struct Tc3Params {
    const uint8_t* src0;   // sh16 [frames][F_in][C0]
    const uint8_t* src1;   // sh16 [frames][F_in][C1] or null (C1 == C0 when present)
    const uint8_t* prev0;  // history row of src0 per clip / stream [B][F_in][C0]: the other parity's buffer (streaming) or the row
                           // carried over from the previous time chunk (offline); null = zero padding
    const uint8_t* prev1;
    const uint8_t* wpk;    // [nhalf][phase][tap][chunk 2][hi N | lo N][8 halves]
    const float* bias;     // [nhalf * N], packed-column order
    const float* gamma;    // [PC] LayerNorm scale / offset by output channel
    const float* beta;
    const float* alpha;
    uint8_t* out;          // sh16 [frames][F_out][PC], F_out = F_conv * nhalf * N / PC
    uint8_t* out2;         // optional second copy of the output with its bins stored [even | odd] (read by a stride-2 unit), or null
    int F_out;
    float wscale_inv;      // undoes the power-of-two weight scale
    int C0, C1;
    int B, T, F_in, F_conv;
    int P, padrow, lead, xlo;
    int nimg;
    int img_mul[2], img_add[2];
    int ntaps;
    int tap_img[T3_MAXTAPS], tap_off[T3_MAXTAPS];
    int nphase;            // (C0 + C1) / 16
    int slots;             // image length in positions = mt*128 + max tap_off
    int plane_bytes;       // byte stride between the 4 planes of a buffer: >= 16 * roundup(nimg*slots, 32), (/16) % 8 == 2
    int total_flat;        // B * (T + padrow) * P
    int ntiles;
    int nabuf;             // image ring depth (2..T3_MAXNB)
    int mt;                // M=128 tiles per iteration (1 or 2)
    int nhalf;             // 1, or 2: CTA parity selects the 64-column half
    int w_half_bytes;      // nphase * ntaps * N * 64
    int fence_mode;        // experiments only: 2 = skip the consumer-side fence.proxy.async
    int src_eo, out_eo;    // bins of the sources / of the output are stored [even | odd] inside each plane
    int pair;              // 1: 2-CTA clusters issue cta_group::2 MMAs (M = 256): CTA r of a cluster owns one 128-position tile and
                           // holds the weights of channel half r; the leader's MMA reads both CTAs' images and weight halves.  The
                           // two tiles of a pair lie pair_m tiles (a whole number of frame rows) apart, so both images start at the
                           // same offset inside their first frame row and one A descriptor serves both CTAs.
    int pair_m;            // P / gcd(P, 128); row tiles: tiles per frame row (the partner works one frame row further)
    int row_tpr;           // 0: tiles are consecutive runs of 128 flat positions.  > 0 ("row tiles", units with F_conv a multiple of
                           // 128): tile u covers bins [128 j, 128 j + 128) of frame row u / row_tpr, j = u % row_tpr -- no pad position is
                           // ever computed and a tile image starts at a fixed offset of its first frame row (fewest box rows)
    int tile2_off;         // flat distance between the two tiles of an iteration (mt == 2): 128, or P for one row tile per row
    int prev_rows;         // 1: prev0 / prev1 hold the row in front of every clip (streaming step, or a later time chunk)
    unsigned long long mg_P, mg_Tp, mg_pair_m, mg_row_tpr;   // floor(2^64 / d) + 1 for d = P, T + padrow, pair_m, row_tpr (0: d = 1 or unused)
    int ld_rr;             // tensor-box loaders: 1 = the warps take the ring buffers in turn, four lanes issue the four planes (see the loader)
    int tm_dmin;           // min(tm_delta): the image may start tm_dmin positions late without losing its first position
    int cluster;           // 1: launched as 2-CTA clusters (the two halves of a 128-channel unit): bulk copies are multicast
    int tma;               // 1: row segments move with 1-D bulk copies; 2: whole tile images move as tensor-map boxes (see tm*)
    // tma == 2: a tile whose frame rows lie inside one clip is fetched with ONE cp.async.bulk.tensor per (plane, image):
    // the box is tm_rows whole frame rows x P positions starting at storage position tm_c[img]; positions outside the
    // plane (the frequency pads) and rows outside the clip (the causal time pad, the tail) are zero-filled by the copy
    // engine.  The image then starts on a row boundary, so the MMA adds (q0 mod P) + tm_delta[img] to its tap offsets.
    // Tiles that straddle two clips fall back to the slot-table cp.async loader.
    int tm_rank;           // 4: {8-byte words of a plane row, plane, t, b}; 5: {words of a 128-byte line | of a position, line | position, plane, t, b}
    int tm_rows;           // frame rows per box
    int tm_c[2];           // box start per image: coordinate 0 (rank 4) / coordinate 1 (rank 5) of the first position of a box row
    int tm_delta[2];       // image index of flat position (rho_a, x) is x + tm_delta[img]
    int tm_par[2];         // [even | odd] sources: which parity half image img reads
    int tm_img_bytes;      // byte offset of image 1 inside a plane (128-byte multiple)
    int tm_box_bytes;      // tm_rows * P * 16
    alignas(64) CUtensorMap tm_map[2];   // src0 / src1
    unsigned long long* timing;   // experiments: CTA 0 writes per-role cycle counters here (null = off)
    int dbg;               // experiments only: 1 = no loads, 2 = no stores, 8 = no LN/split math (the MMA-side switches of round 1
                           // are gone: their tests sat in the issue loop, whose instruction count is on the critical path)
};

// mbarrier wait for warps that are NOT on the critical path (epilogue, loaders): the try_wait carries a suspend-time
// hint so that a waiting warp sleeps in hardware instead of competing for issue slots with the MMA-issuing thread.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITR_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAITR_DONE;\n\t"
        "bra WAITR_LOOP;\n\t"
        "WAITR_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}

// n / d for 0 <= n < 2^22 with a float reciprocal and a one-step fix-up
__device__ __forceinline__ int small_div(int n, int d, float rcp) {
    int q = (int)((float)n * rcp);
    const int r = n - q * d;
    q += (r >= d) ? 1 : 0;
    q -= (r < 0) ? 1 : 0;
    return q;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void tensor_g2s_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tensor_g2s_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int floor_div(int n, int d) { return (n >= 0) ? n / d : -((-n + d - 1) / d); }
// Division by a launch constant through its reciprocal M = floor(2^64 / d) + 1 (Tc3Params::mg_*, 0 for d = 1): exact for every
// 0 <= n < 2^32 and four multiply-adds instead of the ~40 instructions of an integer division -- the tile geometry below runs
// once per tile in the loaders and the MMA warp (whose per-tile bookkeeping leaves the tensor pipe idle) and once per row in
// the epilogue.
__device__ __forceinline__ int fast_div(int n, uint64_t M) {
    if (M == 0) return n;
    const uint32_t mlo = (uint32_t)M, mhi = (uint32_t)(M >> 32);
    const uint64_t t = (uint64_t)(uint32_t)n * mlo;
    const uint64_t u = (uint64_t)(uint32_t)n * mhi + (t >> 32);
    return (int)(u >> 32);
}
__device__ __forceinline__ int floor_div_m(int n, int d, uint64_t M) { return (n >= 0) ? fast_div(n, M) : -fast_div(-n + d - 1, M); }
// tma == 2: does the tile whose image starts at flat position qa fit one clip (-> tensor-map boxes), and where does the
// image start inside its first frame row?  Evaluated identically by the loaders and the MMA warp.
struct T3TileGeo {
    int box;       // 1: tensor boxes, 0: slot-table fallback
    int xoff;      // qa - rho_a * P
    int b, t;      // clip and frame (may be -1: the causal pad row) of frame row rho_a
};
__device__ __forceinline__ T3TileGeo t3_tile_geo(int qa, int P, int Tp, int padrow, int slots, int dmax, int dmin, int prev_rows, uint64_t mgP,
                                                  uint64_t mgTp) {
    T3TileGeo g;
    const int rho_a = floor_div_m(qa + dmin, P, mgP);   // image index of flat position f is f - rho_a * P + delta >= 0
    g.xoff = qa - rho_a * P;                      // >= -dmin
    const int need = fast_div(g.xoff + dmax + slots - 1, mgP) + 1;       // frame rows the MMAs can touch
    g.b = floor_div_m(rho_a, Tp, mgTp);
    const int tp = rho_a - g.b * Tp;
    g.t = tp - padrow;
    // with carried history rows (time-chunked offline calls) the pad row in front of a clip is real data held in another
    // tensor: tiles that touch it (tp == 0) take the slot-table loader, which reads it through prev0 / prev1
    g.box = (rho_a >= 0 && tp + need <= Tp && !(prev_rows && tp < padrow)) ? 1 : 0;
    return g;
}
// ---- CTA-pair (cta_group::2) helpers
// Arrive on the same barrier of CTA `cta` of the cluster.  Relaxed: what the leader consumes afterwards was written by the copy
// engine / read by tcgen05.ld and is ordered by the local barrier the forwarding thread waited on; a release at cluster scope
// costs ~1000 cycles per arrival (measured) and would serialise the forwarding loop.
__device__ __forceinline__ void mbar_arrive_peer(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // completion of the pair's MMAs arrives on both CTAs' barrier
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_w2(uint32_t d_tmem, uint32_t a_low, uint32_t a_high, uint32_t b_low, uint32_t b_high,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_low), "r"(a_high), "r"(b_low), "r"(b_high), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred;
}
// one MMA from 32-bit descriptor words (the high words are loop invariants)
__device__ __forceinline__ void tc_mma_f16_w(uint32_t d_tmem, uint32_t a_low, uint32_t a_high, uint32_t b_low, uint32_t b_high,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_low), "r"(a_high), "r"(b_low), "r"(b_high), "r"(idesc), "r"(accumulate) : "memory");
}

// All MMAs of ONE tap in one asm block.  The issuing thread's instruction stream is on the critical path of the small-N units
// (a tap's MMAs keep the tensor pipe busy for ~100 cycles and the pipe queues only a few of them), so the descriptors are
// assembled with the fewest instructions: one predicate, one B descriptor, the A descriptors by 32-bit adds on the low word.
//   one tile:   a_hi x [b_hi | b_lo] -> d0 (2N columns),  a_lo x b_hi -> d0 (N columns)
//   two tiles:  the same for the second tile (a + t2 -> d1), interleaved so that back-to-back MMAs never chain on one accumulator
__device__ __forceinline__ void tc_mma_tap1(uint32_t d0, uint32_t a_low, uint32_t a_lo_delta, uint32_t a_high, uint32_t b_low,
                                            uint32_t b_high, uint32_t idesc_2n, uint32_t idesc_n, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 al;\n\t"
        ".reg .b64 da, dl, db;\n\t"
        "setp.ne.b32 p, %8, 0;\n\t"
        "mov.b64 db, {%4, %5};\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "add.u32 al, %1, %2;\n\t"
        "mov.b64 dl, {al, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], dl, db, %7, 1;\n\t"
        "}" ::"r"(d0), "r"(a_low), "r"(a_lo_delta), "r"(a_high), "r"(b_low), "r"(b_high), "r"(idesc_2n), "r"(idesc_n), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_tap2(uint32_t d0, uint32_t d1, uint32_t a_low, uint32_t t2, uint32_t a_lo_delta, uint32_t a_high,
                                            uint32_t b_low, uint32_t b_high, uint32_t idesc_2n, uint32_t idesc_n, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 a1, al0, al1;\n\t"
        ".reg .b64 da0, da1, dl0, dl1, db;\n\t"
        "setp.ne.b32 p, %10, 0;\n\t"
        "mov.b64 db, {%6, %7};\n\t"
        "add.u32 a1, %2, %3;\n\t"
        "add.u32 al0, %2, %4;\n\t"
        "add.u32 al1, a1, %4;\n\t"
        "mov.b64 da0, {%2, %5};\n\t"
        "mov.b64 da1, {a1, %5};\n\t"
        "mov.b64 dl0, {al0, %5};\n\t"
        "mov.b64 dl1, {al1, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db, %8, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], da1, db, %8, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], dl0, db, %9, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], dl1, db, %9, 1;\n\t"
        "}" ::"r"(d0), "r"(d1), "r"(a_low), "r"(t2), "r"(a_lo_delta), "r"(a_high), "r"(b_low), "r"(b_high), "r"(idesc_2n), "r"(idesc_n),
        "r"(accumulate)
        : "memory");
}
// pair mode: a_hi x [b_hi | b_lo] of both halves -> d0 (4N columns), a_lo x b_hi of both halves -> d0 + N (2N columns)
__device__ __forceinline__ void tc_mma_tap_pair(uint32_t d0, uint32_t d0n, uint32_t a_low, uint32_t a_lo_delta, uint32_t a_high, uint32_t b_low,
                                                uint32_t b_high, uint32_t idesc_4n, uint32_t idesc_2n, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 al;\n\t"
        ".reg .b64 da, dl, db;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "mov.b64 db, {%5, %6};\n\t"
        "mov.b64 da, {%2, %4};\n\t"
        "add.u32 al, %2, %3;\n\t"
        "mov.b64 dl, {al, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %7, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], dl, db, %8, 1;\n\t"
        "}" ::"r"(d0), "r"(d0n), "r"(a_low), "r"(a_lo_delta), "r"(a_high), "r"(b_low), "r"(b_high), "r"(idesc_4n), "r"(idesc_2n),
        "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void cp_async16_s(uint32_t smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        default: cp_async_wait<5>(); break;
    }
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void st_global_32B(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// tcgen05.ld of 32 columns WITHOUT the wait, so that several loads are in flight before one tcgen05.wait::ld.  The values
// may be used only after tmem_ld_wait() + tmem_ld_fence32() on the same registers (the empty asm ties them to the wait).
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_fence32(uint32_t* r) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// hi/lo split of 8 consecutive values into two 16-byte vectors of halves
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// 8 consecutive channels of an sh16 record back to floats
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float* v) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
        v[2 * i] = a.x + b.x;
        v[2 * i + 1] = a.y + b.y;
    }
}

// LayerNorm (two-pass: mean, centred variance -- the reference's non-fused Keras path) + PReLU over the CG values
// v[0..CG/2) (float2 pairs) in place, gamma / beta from shared memory.  Packed fp32x2 arithmetic (FADD2 / FFMA2 on
// sm_100) halves the instruction count of the element-wise passes; four partial sums keep the reductions off one
// dependent chain.
template <int CG>
__device__ __forceinline__ void ln_prelu_s(float2* v, const float* gamma_s, const float* beta_s, float alpha) {
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CG / 2; i += 2) {
        s0 = __fadd2_rn(s0, v[i]);
        s1 = __fadd2_rn(s1, v[i + 1]);
    }
    const float mean = ((s0.x + s0.y) + (s1.x + s1.y)) * (1.0f / CG);
    const float2 nmean = make_float2(-mean, -mean);
    float2 q0 = make_float2(0.f, 0.f), q1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CG / 2; i += 2) {
        const float2 d0 = __fadd2_rn(v[i], nmean), d1 = __fadd2_rn(v[i + 1], nmean);
        q0 = __ffma2_rn(d0, d0, q0);
        q1 = __ffma2_rn(d1, d1, q1);
    }
    const float inv = rsqrtf(((q0.x + q0.y) + (q1.x + q1.y)) * (1.0f / CG) + LN_EPS);
    const float2 inv2 = make_float2(inv, inv), minv2 = make_float2(-mean * inv, -mean * inv);
#pragma unroll
    for (int i = 0; i < CG / 2; i += 2) {
        const float4 g = *reinterpret_cast<const float4*>(gamma_s + 2 * i);
        const float4 b = *reinterpret_cast<const float4*>(beta_s + 2 * i);
        float2 y0 = __ffma2_rn(__ffma2_rn(v[i], inv2, minv2), make_float2(g.x, g.y), make_float2(b.x, b.y));
        float2 y1 = __ffma2_rn(__ffma2_rn(v[i + 1], inv2, minv2), make_float2(g.z, g.w), make_float2(b.z, b.w));
        if (y0.x < 0.f) y0.x *= alpha;
        if (y0.y < 0.f) y0.y *= alpha;
        if (y1.x < 0.f) y1.x *= alpha;
        if (y1.y < 0.f) y1.y *= alpha;
        v[i] = y0;
        v[i + 1] = y1;
    }
}

// hi/lo split of 8 consecutive values (4 float2 pairs) into two 16-byte vectors of halves
__device__ __forceinline__ void split8p(const float2* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __float22half2_rn(v[i]);
        const float2 r = __fadd2_rn(v[i], make_float2(-__low2float(hh), -__high2float(hh)));
        const __half2 ll = __float22half2_rn(r);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// cycle counter for the per-role profile (Tc3Params::timing); not read at all in normal runs -- the reads sat in the loops whose
// instruction count matters
__device__ __forceinline__ long long t3_clock(bool timed) { return timed ? clock64() : 0ll; }

// N: conv channels of this CTA (32 / 64); PC: channels per OUTPUT pixel (N, or 32 for the 64-column sub-pixel
// shuffle that makes two pixels); LN: LayerNorm + PReLU over each PC-channel group (false: bias only).
// PAIR: the cta_group::2 variant (must be launched as 2-CTA clusters; Tc3Params::pair set)
template <int N, int PC, bool LN, bool PAIR = false>
__global__ void __launch_bounds__(T3_THREADS, 1) conv_tc3_kernel(const __grid_constant__ Tc3Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* a_full = bars;                     // [T3_MAXNB]
    uint64_t* a_empty = bars + T3_MAXNB;         // [T3_MAXNB]
    uint64_t* acc_full = bars + 2 * T3_MAXNB;    // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint64_t* w_full = acc_empty + 2;            // [1]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 24);
    uint64_t* peer_full = reinterpret_cast<uint64_t*>(smem_raw + T3_PEER_OFF);   // [T3_MAXNB] pair mode, leader CTA: the partner's image landed
    uint64_t* peer_acc_empty = peer_full + T3_MAXNB;                             // [2] pair mode, leader CTA: the partner drained its accumulator
    float* par_s = reinterpret_cast<float*>(smem_raw + T3_PAR_OFF);   // bias[64] | gamma[64] | beta[64]
    int* slot_tbl = reinterpret_cast<int*>(smem_raw + T3_TBL_OFF);
    uint8_t* wsm = smem_raw + T3_FIXED_BYTES;
    uint8_t* abuf0 = wsm + p.w_half_bytes;
    const uint32_t abuf_bytes = 4u * (uint32_t)p.plane_bytes;

    // Programmatic dependent launch: let the next kernel of the stream start its prologue (barrier / TMEM set-up,
    // weight load -- none of which depends on this kernel's output) on SMs that this grid has already left.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool timed = p.timing != nullptr;
    constexpr int MMA_WARP = T3_EPI_WARPS + T3_LD_WARPS;
    constexpr uint32_t ACC_COLS = 2 * T3_MT * 2 * N;   // 2 buffers x 2 tiles x (N + N) columns
    constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));

    const int half = (int)blockIdx.x % p.nhalf;
    if (threadIdx.x == 0) {
        for (int i = 0; i < T3_MAXNB; ++i) {
            mbar_init(&a_full[i], p.tma == 1 ? T3_LD_WARPS : T3_LD_THREADS);   // bulk rows: one arrival per loader warp (+ tx bytes)
            mbar_init(&a_empty[i], p.cluster ? 2 : 1);     // cluster: both CTAs' MMAs must be done before either refills
            mbar_init(&peer_full[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], PAIR ? 256 : 128 * p.mt);   // pair: both warp groups drain every tile (one channel half each)
            mbar_init(&peer_acc_empty[i], 1);
        }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 128) {   // pair mode stages the bias of both channel halves
        const int nb = PAIR ? 2 * N : N;
        par_s[threadIdx.x] = ((int)threadIdx.x < nb) ? __ldg(p.bias + (PAIR ? 0 : half * N) + threadIdx.x) : 0.f;
        if (threadIdx.x < 64) {
            par_s[128 + threadIdx.x] = (LN && threadIdx.x < PC) ? __ldg(p.gamma + threadIdx.x) : 0.f;
            par_s[192 + threadIdx.x] = (LN && threadIdx.x < PC) ? __ldg(p.beta + threadIdx.x) : 0.f;
        }
    }
    if (warp == MMA_WARP) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.cluster || PAIR) cluster_sync_all();     // the partner's barriers exist before anything is multicast to them
    const uint32_t tmem_base = *tmem_ptr_s;
    const int Tp = p.T + p.padrow;
    const int cta = (int)blockIdx.x / p.nhalf, ncta = (int)gridDim.x / p.nhalf;
    const int my_tiles = (cta < p.ntiles) ? (p.ntiles - cta + ncta - 1) / ncta : 0;
    // first flat position of the tile CTA-rank `r` works on in iteration `it`.  Pair mode: cluster unit u = (block, j) maps to
    // the 128-position tiles block * 2m + j (rank 0) and block * 2m + j + m (rank 1), m * 128 = a whole number of frame rows.
    auto tile_q = [&](int it, int r) {
        const int u = cta + it * ncta;
        int unit;                                  // index of the 128-position tile
        if (PAIR) {
            const int blk = fast_div(u, p.mg_pair_m), j = u - blk * p.pair_m;
            unit = blk * 2 * p.pair_m + j + r * p.pair_m;
        } else {
            unit = u * p.mt;
        }
        if (p.row_tpr == 0) return unit * 128;
        const int rw = fast_div(unit, p.mg_row_tpr);
        return rw * p.P + p.xlo + (unit - rw * p.row_tpr) * 128;
    };

    // Register re-partitioning: every role resizes its allocation as the first thing in its branch (warpgroups 0-1 epilogue,
    // 2 loaders, 3 = MMA issuer + three idle warps); the shrinking roles release what the epilogue's increase waits for.
    if (warp < T3_EPI_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T3_REGS_EPI));
        // ================================================================= epilogue
        // Accumulator (it, mt) is drained by warp group ((it * p.mt + mt) & 1): with two tiles per iteration each
        // group owns one of them, with one tile per iteration the groups alternate iterations.
        const int eg = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;   // TMEM lane == tile row
        const float alpha = LN ? __ldg(p.alpha) : 0.f;
        asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel may still read what we overwrite
        constexpr int NPX = N / PC;            // output pixels per conv pixel made by this CTA
        constexpr int CPP = PC / 8;            // 16-byte chunks per part of an output pixel
        const int npx = NPX * p.nhalf;         // output bins per conv bin
        const long long out_rs = (long long)p.F_out * (PC * 4);   // bytes per output frame row
        const long long plane = (long long)p.F_out * 16;          // bytes per output plane
        long long t_wait = 0, t_ld = 0;                           // per-role cycle counters (Tc3Params::timing)
        const long long t_begin = t3_clock(timed);
        for (int it = 0; it < my_tiles; ++it) {
            const int mt = (p.mt == 2) ? eg : 0;
            if (p.mt == 1 && !PAIR && (it & 1) != eg) continue;
            const int ab = it & 1;
            const long long tq0 = t3_clock(timed);
            mbar_wait_relaxed(&acc_full[ab], (it >> 1) & 1);
            tc_fence_after();
            const long long tq1 = t3_clock(timed);
            t_wait += tq1 - tq0;
            // pair accumulator: [hh | hl + lh] of half 0, then [hh + lh | hl] of half 1 (the a_lo x b_hi product lands on columns N .. 3N)
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(PAIR ? (ab * 2 + eg) * 2 * N : (ab * p.mt + mt) * 2 * N);
            const float* bias_s = par_s + (PAIR ? eg * N : 0);
            float2 v[N / 2];
            {
                // both column blocks of the accumulator, the loads in flight together: one wait per 32 + 32..64 columns
                uint32_t rv[N], ru[32];
#pragma unroll
                for (int c = 0; c < N; c += 32) tmem_ld32_issue(taddr + c, rv + c);          // a_hi*b_hi (+ a_lo*b_hi)
                tmem_ld32_issue(taddr + N, ru);                                                // a_hi*b_lo (+ a_lo*b_hi)
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < N; c += 32) tmem_ld_fence32(rv + c);
                tmem_ld_fence32(ru);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] = __fadd2_rn(make_float2(__uint_as_float(rv[2 * i]), __uint_as_float(rv[2 * i + 1])),
                                      make_float2(__uint_as_float(ru[2 * i]), __uint_as_float(ru[2 * i + 1])));
                if (N == 64) {
                    tmem_ld32_issue(taddr + N + 32, ru);
                    tmem_ld_wait();
                    tmem_ld_fence32(ru);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        v[16 + i] = __fadd2_rn(make_float2(__uint_as_float(rv[32 + 2 * i]), __uint_as_float(rv[32 + 2 * i + 1])),
                                               make_float2(__uint_as_float(ru[2 * i]), __uint_as_float(ru[2 * i + 1])));
                }
            }
            // the accumulator is in registers: hand the TMEM buffer back to the MMA warp right away
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);
            t_ld += t3_clock(timed) - tq1;
            {
                const float2 sc = make_float2(p.wscale_inv, p.wscale_inv);
#pragma unroll
                for (int c = 0; c < N; c += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(bias_s + c);
                    v[c / 2] = __ffma2_rn(v[c / 2], sc, make_float2(bv.x, bv.y));
                    v[c / 2 + 1] = __ffma2_rn(v[c / 2 + 1], sc, make_float2(bv.z, bv.w));
                }
            }
            if (LN && !(p.dbg & 8)) {
#pragma unroll
                for (int g = 0; g < NPX; ++g) ln_prelu_s<PC>(v + g * PC / 2, par_s + 128, par_s + 192, alpha);
            }
            // where this row goes (computed only now: nothing of it has to stay live across the arithmetic above)
            const int q = tile_q(it, half) + mt * p.tile2_off + row;
            const int rho = fast_div(q, p.mg_P);
            const int x = q - rho * p.P;
            const int b = fast_div(rho, p.mg_Tp);
            const int t = (rho - b * Tp) - p.padrow;
            const bool valid = (q < p.total_flat) && (t >= 0) && (x >= p.xlo) && (x < p.xlo + p.F_conv);
            // output bins of this conv pixel: obin0 + g (g < NPX); storage position inside a plane of the frame row
            const int ohalf = PAIR ? eg : half;       // channel half (= output pixel group) this thread finishes
            const int obin0 = (x - p.xlo) * npx + ohalf * NPX;
            // one copy in the order the flags ask for; units whose output also feeds a stride-2 unit write a second,
            // [even | odd] copy (Tc3Params::out2) so that consumer moves whole plane rows instead of 16-byte pieces
            auto emit = [&](uint8_t* base, const int eo) {
                const int opos0 = eo ? (obin0 & 1) * (p.F_out >> 1) + (obin0 >> 1) : obin0;
                uint8_t* orow = base + ((long long)b * p.T + t) * out_rs + (long long)opos0 * 16;
                if (NPX == 2 && !eo) {
                    // two neighbouring output bins per thread: one 32-byte store per chunk (whole sectors)
#pragma unroll
                    for (int c = 0; c < CPP; ++c) {
                        uint4 h0, l0, h1, l1;
                        split8p(v + 4 * c, h0, l0);
                        split8p(v + PC / 2 + 4 * c, h1, l1);
                        st_global_32B(orow + c * plane, h0, h1);
                        st_global_32B(orow + (CPP + c) * plane, l0, l1);
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < NPX; ++g) {
                        // with [even | odd] storage the second bin of a pair lives half a plane further
                        uint8_t* o = orow + (eo ? (long long)g * (p.F_out >> 1) * 16 : (long long)g * 16);
#pragma unroll
                        for (int c = 0; c < CPP; ++c) {
                            uint4 hi, lo;
                            split8p(v + g * PC / 2 + 4 * c, hi, lo);
                            *reinterpret_cast<uint4*>(o + c * plane) = hi;
                            *reinterpret_cast<uint4*>(o + (CPP + c) * plane) = lo;
                        }
                    }
                }
            };
            if (valid && !(p.dbg & 2)) {
                emit(p.out, p.out_eo);
                if (!PAIR && p.out2) emit(p.out2, 1);
            }
        }
        if (p.timing && blockIdx.x == 0 && threadIdx.x == 0) {
            p.timing[4] = (unsigned long long)t_wait;
            p.timing[5] = (unsigned long long)t_ld;
            p.timing[6] = (unsigned long long)(t3_clock(timed) - t_begin);
        }
    } else if (warp < MMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T3_REGS_LD));
        // ================================================================= loaders
        // Loader warp w owns plane w = (part hi|lo, chunk) of every buffer; its lanes walk the table entries
        // e = lane + 32 k, so one cp.async instruction moves 32 consecutive 16-byte slots.  A table entry is the
        // 16-byte-unit offset of (frame row, bin) inside a source plane 0; the plane of the current phase adds
        // a constant.  Entry e = img * slots + slot; the planes of a buffer hold both images back to back.
        const int lt = threadIdx.x - 32 * T3_EPI_WARPS;
        const int lw = lt >> 5;
        const int part = lw >> 1, chunk = lw & 1;
        constexpr int ESTEP = 32;
        const int e0 = lane;
        const int nslot = p.nimg * p.slots;
        const int nit = (nslot + ESTEP - 1) / ESTEP;           // entries >= nslot are -1 (zero fill into padding)
        const uint32_t dst0 = smem_u32(abuf0) + (uint32_t)(lw * p.plane_bytes + e0 * 16);
        const int cpp0 = p.C0 >> 3;                            // chunks per part of a source (C1 == C0)
        const int rs16 = p.F_in * (p.C0 >> 2);                 // 16-byte units per source frame row
        const int NB = p.nabuf;
        const float rcpP = 1.0f / (float)p.P, rcpTp = 1.0f / (float)Tp;
        long long tl_table = 0, tl_wait = 0, tl_fence = 0, tl_issue = 0;   // per-role cycle counters (Tc3Params::timing)
        const long long tl_begin = t3_clock(timed);
        // Our sources are the previous kernels' outputs: wait for them (programmatic dependent launch) -- but only right before
        // the first copy, so that the slot table of the first tile is built while the previous kernel drains.
        bool dep_waited = false;
        auto dep_wait = [&]() {
            if (!dep_waited) {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                dep_waited = true;
            }
        };
        int buf = 0, round = 0;        // ring position of phase g and the parity of its use count
        int g = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int q0 = tile_q(it, half) - p.lead;
            const long long tl0 = t3_clock(timed);
            if (p.tma == 2) {
                const T3TileGeo tg = t3_tile_geo(q0, p.P, Tp, p.padrow, p.slots, max(p.tm_delta[0], p.tm_delta[1]), p.tm_dmin, p.prev_rows, p.mg_P, p.mg_Tp);
                // pair mode: one A descriptor serves both CTAs, so both must use the same image layout
                const int peer_box = PAIR ? t3_tile_geo(tile_q(it, half ^ 1) - p.lead, p.P, Tp, p.padrow, p.slots,
                                                          max(p.tm_delta[0], p.tm_delta[1]), p.tm_dmin, p.prev_rows, p.mg_P, p.mg_Tp).box : 1;
                if (tg.box && peer_box) {
                    tl_table += t3_clock(timed) - tl0;      // (box tiles: the tile geometry)
                    dep_wait();
                    // One box per (plane, image).  Default: loader warp w issues plane w of every phase (lane 0); 32 arrivals per
                    // plane keep the barrier count of the fallback.  Issuing a box costs the lane ~500 cycles (expect_tx, the tensor
                    // copy, the arrival), which bounds units with many short phases (up_sampling o inconv: 8 phases of 2 taps): for
                    // those (Tc3Params::ld_rr) the warps take the BUFFERS in turn and lanes 0-3 issue the four planes at once.  (A
                    // buffer always belongs to the same warp, so its waits on that buffer's barrier stay in order -- a warp that
                    // skipped a use could mistake an older completed phase of the same parity for the one it needs.)
                    const bool rr = p.ld_rr != 0;
                    for (int ph = 0; ph < p.nphase; ++ph, ++g) {
                        if (!rr || (buf & (T3_LD_WARPS - 1)) == lw) {
                            const long long tl1 = t3_clock(timed);
                            if (g >= NB) mbar_wait_relaxed(&a_empty[buf], round ^ 1);
                            const long long tl2 = t3_clock(timed);
                            tl_wait += tl2 - tl1;
                            if (rr ? (lane < 4) : (lane == 0)) {
                                const int pw = rr ? lane : lw;              // plane of the buffer this lane fills
                                const int c0 = ph * T3_KCH;
                                const bool first = c0 < p.C0;
                                const int cc = first ? c0 : c0 - p.C0;
                                const int plane = (pw >> 1) * cpp0 + (cc >> 3) + (pw & 1);
                                const CUtensorMap* map = &p.tm_map[first ? 0 : 1];
                                uint8_t* dstp = abuf0 + (size_t)buf * abuf_bytes + (size_t)pw * p.plane_bytes;
                                if (!(p.dbg & 1)) {
                                    mbar_expect_tx(&a_full[buf], (uint32_t)(p.nimg * p.tm_box_bytes));
                                    for (int img = 0; img < p.nimg; ++img) {
                                        const int pl = p.src_eo ? 2 * plane + p.tm_par[img] : plane;
                                        if (p.tm_rank == 4)
                                            tensor_g2s_4d(dstp + (size_t)img * p.tm_img_bytes, map, p.tm_c[img], pl, tg.t, tg.b, &a_full[buf]);
                                        else
                                            tensor_g2s_5d(dstp + (size_t)img * p.tm_img_bytes, map, 0, p.tm_c[img], pl, tg.t, tg.b, &a_full[buf]);
                                    }
                                }
                                mbar_arrive_n(&a_full[buf], 32);
                            }
                            __syncwarp();
                            tl_issue += t3_clock(timed) - tl2;
                        }
                        if (++buf == NB) {
                            buf = 0;
                            round ^= 1;
                        }
                    }
                    continue;
                }
            }
            if (it > 0) asm volatile("bar.sync 1, %0;" ::"n"(T3_LD_THREADS) : "memory");   // everyone is done with the old table
            {
                // (row, x) of slot 0 by two real divisions per tile; every entry then needs only small-number
                // divisions (float reciprocal + fix-up): slot < 2^11, rows per tile < 2^11
                const int rho0 = floor_div_m(q0, p.P, p.mg_P);
                const int x0 = q0 - rho0 * p.P;
                const int b0 = (rho0 >= 0) ? fast_div(rho0, p.mg_Tp) : 0;
                const int t0 = rho0 - b0 * Tp;                     // row inside clip b0 (negative before the first clip)
                for (int e = lt; e < nit * ESTEP; e += T3_LD_THREADS) {
                    const int img = (e >= p.slots) ? 1 : 0;
                    const int slot = e - img * p.slots;
                    const int xs = x0 + slot;
                    const int drho = small_div(xs, p.P, rcpP);
                    const int x = xs - drho * p.P;
                    const int ts = t0 + drho;
                    int o = -1;
                    if (e < nslot && ts >= 0 && q0 + slot < p.total_flat) {
                        const int db = small_div(ts, Tp, rcpTp);
                        const int t = ts - db * Tp - p.padrow;
                        const int fi = p.img_mul[img] * x + p.img_add[img];
                        if (fi >= 0 && fi < p.F_in) {
                            const int pos = p.src_eo ? (fi & 1) * (p.F_in >> 1) + (fi >> 1) : fi;
                            if (t >= 0) o = ((b0 + db) * p.T + t) * rs16 + pos;
                            else if (p.prev0 != nullptr) o = ((b0 + db) * rs16 + pos) | T3_PREV_FLAG;   // carried history row
                        }
                    }
                    slot_tbl[e] = o;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(T3_LD_THREADS) : "memory");
            // TMA mode: lane r describes the part of frame row (first row of the tile + r) that lies inside the image,
            // and remembers which of its table entries are pads (bit k <-> entry lane + 32 k) so that the per-buffer
            // zero fill touches only those
            int seg_dst = 0, seg_src = 0, seg_n = 0;
            uint32_t padmask = 0;
            if (p.tma == 1) {
                for (int k = 0; k < nit; ++k) padmask |= (slot_tbl[e0 + k * ESTEP] < 0) ? (1u << k) : 0u;
                // two-image (stride-2) units: lanes 0-15 describe image 0, lanes 16-31 image 1 (<= 16 rows per tile);
                // their sources are stored [even | odd], so each image's bins are consecutive storage positions
                const int img = (p.nimg == 2) ? (lane >> 4) : 0;
                const int r = (p.nimg == 2) ? (lane & 15) : lane;
                const int mul = p.img_mul[img], add = p.img_add[img];
                const int rho0 = floor_div_m(q0, p.P, p.mg_P);
                const int rho = rho0 + r;
                const int qs = rho * p.P;                           // flat position of x = 0 of this row
                const int x_first = (mul == 1) ? -add : ((add < 0) ? (1 - add) / 2 : 0);       // first x with a real source bin
                const int x_end = (mul == 1) ? p.F_in - add : (p.F_in - add + 1) / 2;          // first x past the last bin
                int xa = max(q0 - qs, x_first);
                int xb = min(min(q0 + p.slots - qs, x_end), p.P);
                const int b = (rho >= 0) ? fast_div(rho, p.mg_Tp) : 0;
                const int t = rho - b * Tp - p.padrow;
                if (rho >= 0 && (long long)qs < (long long)p.total_flat && t >= 0 && xb > xa) {
                    const int fi = mul * xa + add;
                    const int pos = p.src_eo ? (fi & 1) * (p.F_in >> 1) + (fi >> 1) : fi;
                    seg_dst = img * p.slots + qs + xa - q0;
                    seg_src = (b * p.T + t) * rs16 + pos;
                    seg_n = xb - xa;
                }
            }
            tl_table += t3_clock(timed) - tl0;
            dep_wait();
            for (int ph = 0; ph < p.nphase; ++ph, ++g) {
                const long long tl1 = t3_clock(timed);
                if (g >= NB) mbar_wait_relaxed(&a_empty[buf], round ^ 1);
                tl_wait += t3_clock(timed) - tl1;
                const int c0 = ph * T3_KCH;
                const bool first = c0 < p.C0;
                const uint8_t* src = first ? p.src0 : p.src1;
                const int cc = first ? c0 : c0 - p.C0;
                const long long plane_off = (long long)((part * cpp0 + (cc >> 3) + chunk) * p.F_in) * 16;
                const uint8_t* pb = src + plane_off;                                                          // source plane
                const uint8_t* pbp = (first ? p.prev0 : p.prev1) + plane_off;                                 // history plane
                uint32_t d = dst0 + (uint32_t)buf * abuf_bytes;
                const int* tb = slot_tbl + e0;
                if (p.tma == 1) {
                    // zero the pad slots of this plane (generic proxy), then hand the row segments to the copy engine
                    const long long tf0 = t3_clock(timed);
                    if (padmask) {
                        for (uint32_t m = padmask; m; m &= m - 1) {
                            const int k = __ffs(m) - 1;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d + (uint32_t)k * (ESTEP * 16)), "r"(0) : "memory");
                        }
                        fence_proxy_async();
                    }
                    const long long tf1 = t3_clock(timed);
                    tl_fence += tf1 - tf0;
                    if (seg_n > 0 && !(p.dbg & 1)) {
                        mbar_expect_tx(&a_full[buf], (uint32_t)seg_n * 16);
                        uint8_t* dstp = abuf0 + (size_t)buf * abuf_bytes + (size_t)lw * p.plane_bytes + (size_t)seg_dst * 16;
                        if (!p.cluster)
                            bulk_g2s(dstp, pb + (long long)seg_src * 16, (uint32_t)seg_n * 16, &a_full[buf]);
                        else if (part == half)      // CTA 0 fetches the hi planes, CTA 1 the lo planes, for both CTAs
                            bulk_g2s_mc(dstp, pb + (long long)seg_src * 16, (uint32_t)seg_n * 16, &a_full[buf], (uint16_t)3);
                    }
                    tl_issue += t3_clock(timed) - tf1;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[buf]);
                } else {
                    if (!(p.dbg & 1))
#pragma unroll 4
                    for (int k = 0; k < nit; ++k) {
                        const int o = tb[k * ESTEP];
                        const uint8_t* base = (o & T3_PREV_FLAG) && o >= 0 ? pbp : pb;
                        cp_async16_s(d, base + (long long)(o < 0 ? 0 : (o & (T3_PREV_FLAG - 1))) * 16, (o >= 0) ? 16 : 0);
                        d += ESTEP * 16;
                    }
                    // completion is signalled by the copy engine itself: a_full[buf] collects one arrival per loader
                    // thread, each triggered when that thread's copies above have landed -- the loaders never wait for
                    // data, only for a free buffer, so signalling is decoupled from how far ahead they can issue
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&a_full[buf])) : "memory");
                }
                if (++buf == NB) {
                    buf = 0;
                    round ^= 1;
                }
            }
        }
        cp_async_wait<0>();
        if (p.timing && blockIdx.x == 0 && lt == 0) {
            p.timing[7] = (unsigned long long)tl_wait;
            p.timing[8] = (unsigned long long)tl_table;
            p.timing[9] = (unsigned long long)(t3_clock(timed) - tl_begin);
            p.timing[10] = (unsigned long long)tl_fence;
            p.timing[11] = (unsigned long long)tl_issue;
        }
    } else if (warp == MMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T3_REGS_MMA));
        // ================================================================= MMA issuer (+ one-off weight load)
        // The whole warp stays converged (waits are warp-wide); one elected lane issues, so the tcgen05
        // instructions see warp-uniform operands and need no per-lane serialisation loop.
        const uint32_t leader = elect_one();
        if (my_tiles > 0) {
            if (leader) {
                mbar_arrive_expect_tx(w_full, (uint32_t)p.w_half_bytes);
                const uint8_t* wg = p.wpk + (size_t)half * p.w_half_bytes;
                for (int o = 0; o < p.w_half_bytes; o += 16384) {
                    const int n = (p.w_half_bytes - o < 16384) ? p.w_half_bytes - o : 16384;
                    bulk_g2s(wsm + o, wg + o, (uint32_t)n, w_full);
                }
            }
            mbar_wait(w_full, 0);
        }
        constexpr uint32_t IDESC_2N = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, N' = 2N
        constexpr uint32_t IDESC_N = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        // pair mode: M = 256 over both CTAs, N' = both CTAs' [b_hi | b_lo] blocks = 4N, then a_lo x both b_hi blocks = 2N
        constexpr uint32_t IDESC_P4N = (1u << 4) | ((uint32_t)((4 * N) >> 3) << 17) | ((256u >> 4) << 24);
        constexpr uint32_t IDESC_P2N = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((256u >> 4) << 24);
        // Descriptors differ only in their 14-bit start-address field (bytes >> 4): keep the invariant high words and
        // per-tap address deltas in registers so that one MMA costs a couple of integer adds.
        const uint64_t da0 = make_desc(smem_u32(abuf0), p.plane_bytes);
        const uint64_t db0 = make_desc(smem_u32(wsm), 2 * N * 16);
        const uint32_t da_hiw = (uint32_t)(da0 >> 32), db_hiw = (uint32_t)(db0 >> 32);
        const uint32_t da_low0 = (uint32_t)da0, db_low0 = (uint32_t)db0;
        const uint32_t a_lo_delta = (uint32_t)(2 * p.plane_bytes) >> 4;
        const uint32_t abuf16 = abuf_bytes >> 4;
        uint32_t tapd[T3_MAXTAPS];
#pragma unroll
        for (int tap = 0; tap < T3_MAXTAPS; ++tap)
            tapd[tap] = (tap < p.ntaps) ? (uint32_t)(p.tap_img[tap] * p.slots + p.tap_off[tap]) : 0u;
        int buf = 0, round = 0;
        long long tm_full = 0, tm_acc = 0;                         // per-role cycle counters (Tc3Params::timing)
        const long long tm_begin = t3_clock(timed);
        const uint32_t t2 = (uint32_t)p.tile2_off;                 // second tile of an iteration, in image positions
        uint32_t tap_img1 = 0;                                     // bit tap: the tap reads image 1
        for (int tap = 0; tap < p.ntaps; ++tap) tap_img1 |= (uint32_t)p.tap_img[tap] << tap;
        const int ntaps = p.ntaps;
        const bool mt2 = p.mt == 2;
        if (PAIR && half == 1) {
            // The partner of the leader issues no MMAs: it forwards "my image landed" and "my accumulator is drained" to the
            // leader's barriers, in the order the leader waits for them.
            for (int it = 0; it < my_tiles; ++it) {
                if (it >= 2) {
                    mbar_wait(&acc_empty[it & 1], ((it >> 1) - 1) & 1);
                    if (lane == 0) mbar_arrive_peer(&peer_acc_empty[it & 1], 0);
                }
                const int dm = max(p.tm_delta[0], p.tm_delta[1]);
                const bool boxed = t3_tile_geo(tile_q(it, 0) - p.lead, p.P, Tp, p.padrow, p.slots, dm, p.tm_dmin, p.prev_rows, p.mg_P, p.mg_Tp).box &&
                                   t3_tile_geo(tile_q(it, 1) - p.lead, p.P, Tp, p.padrow, p.slots, dm, p.tm_dmin, p.prev_rows, p.mg_P, p.mg_Tp).box;
                for (int ph = 0; ph < p.nphase; ++ph) {
                    mbar_wait(&a_full[buf], round);
                    if (!boxed) fence_proxy_async();   // only the cp.async fallback writes through the generic proxy
                    if (lane == 0) mbar_arrive_peer(&peer_full[buf], 0);
                    __syncwarp();
                    if (++buf == p.nabuf) {
                        buf = 0;
                        round ^= 1;
                    }
                }
            }
        } else
        for (int it = 0; it < my_tiles; ++it) {
            const int accb = it & 1;
            // tensor-box tiles: the image starts on a frame-row boundary and image 1 sits at tm_img_bytes
            uint32_t adj0 = 0, adj1 = 0;
            bool boxed = false;
            if (p.tma == 2) {
                const T3TileGeo tg = t3_tile_geo(tile_q(it, 0) - p.lead, p.P, Tp, p.padrow, p.slots, max(p.tm_delta[0], p.tm_delta[1]), p.tm_dmin, p.prev_rows, p.mg_P, p.mg_Tp);
                const int peer_box = PAIR ? t3_tile_geo(tile_q(it, 1) - p.lead, p.P, Tp, p.padrow, p.slots,
                                                          max(p.tm_delta[0], p.tm_delta[1]), p.tm_dmin, p.prev_rows, p.mg_P, p.mg_Tp).box : 1;
                if (tg.box && peer_box) {
                    boxed = true;
                    adj0 = (uint32_t)(tg.xoff + p.tm_delta[0]);
                    adj1 = (uint32_t)(tg.xoff + p.tm_delta[1] + (p.tm_img_bytes >> 4) - p.slots);
                }
            }
            uint32_t toff[T3_MAXTAPS];   // image offset of every tap for this tile
#pragma unroll
            for (int tap = 0; tap < T3_MAXTAPS; ++tap) toff[tap] = tapd[tap] + (((tap_img1 >> tap) & 1u) ? adj1 : adj0);
            const long long tm0 = t3_clock(timed);
            if (it >= 2) {
                mbar_wait(&acc_empty[accb], ((it >> 1) - 1) & 1);
                if (PAIR) mbar_wait(&peer_acc_empty[accb], ((it >> 1) - 1) & 1);
            }
            tc_fence_after();
            tm_acc += t3_clock(timed) - tm0;
            uint32_t wlow = db_low0;
            const uint32_t d0 = tmem_base + (uint32_t)(PAIR ? accb * 4 * N : accb * p.mt * 2 * N);
            for (int ph = 0; ph < p.nphase; ++ph) {
                const long long tm1 = t3_clock(timed);
                mbar_wait(&a_full[buf], round);
                if (PAIR) mbar_wait(&peer_full[buf], round);   // (a cluster-scope acquire here costs ~1000 cycles per wait: measured)
                // cp.async / st.shared wrote through the generic proxy, the MMA reads through the async proxy; tensor-box tiles
                // were written by the copy engine itself and need no proxy fence
                if (!boxed && p.fence_mode != 2) fence_proxy_async();
                tc_fence_after();
                tm_full += t3_clock(timed) - tm1;
                if (leader) {
                    const uint32_t alow = da_low0 + (uint32_t)buf * abuf16;
                    if (PAIR) {
#pragma unroll
                        for (int tap = 0; tap < T3_MAXTAPS; ++tap)
                            if (tap < ntaps) {
                                tc_mma_tap_pair(d0, d0 + N, alow + toff[tap], a_lo_delta, da_hiw, wlow, db_hiw, IDESC_P4N, IDESC_P2N,
                                                (tap == 0) ? (uint32_t)ph : 1u);
                                wlow += N * 4;   // next (phase, tap) stage: N * 64 bytes
                            }
                    } else if (mt2) {
#pragma unroll
                        for (int tap = 0; tap < T3_MAXTAPS; ++tap)
                            if (tap < ntaps) {
                                tc_mma_tap2(d0, d0 + 2 * N, alow + toff[tap], t2, a_lo_delta, da_hiw, wlow, db_hiw, IDESC_2N, IDESC_N,
                                            (tap == 0) ? (uint32_t)ph : 1u);
                                wlow += N * 4;
                            }
                    } else {
#pragma unroll
                        for (int tap = 0; tap < T3_MAXTAPS; ++tap)
                            if (tap < ntaps) {
                                tc_mma_tap1(d0, alow + toff[tap], a_lo_delta, da_hiw, wlow, db_hiw, IDESC_2N, IDESC_N, (tap == 0) ? (uint32_t)ph : 1u);
                                wlow += N * 4;
                            }
                    }
                    if (PAIR) tc_commit_pair(&a_empty[buf]);
                    else if (p.cluster) tc_commit_mc(&a_empty[buf], (uint16_t)3);
                    else tc_commit(&a_empty[buf]);
                }
                __syncwarp();
                if (++buf == p.nabuf) {
                    buf = 0;
                    round ^= 1;
                }
            }
            if (leader) {
                if (PAIR) tc_commit_pair(&acc_full[accb]);
                else tc_commit(&acc_full[accb]);
            }
            __syncwarp();
        }
        if (p.timing && blockIdx.x == 0 && lane == 0) {
            p.timing[0] = (unsigned long long)(t3_clock(timed) - tm_begin);
            p.timing[1] = (unsigned long long)tm_full;
            p.timing[2] = (unsigned long long)tm_acc;
            p.timing[3] = (unsigned long long)my_tiles;
        }
    }

    else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T3_REGS_IDLE));
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (p.cluster || PAIR) cluster_sync_all();     // nobody leaves while the partner can still write to / arrive on this CTA
    if (warp == MMA_WARP) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace nunet
