// Host engine + C ABI (include/nunet_b200.h) of the B200-native NUNet-TLS-LSTM inference path.
//
// The engine turns the reference topology (models/proposed.py:284-625 offline; the one-frame stateful
// form converter_proposed.py:188-867) into a static list of kernel launches ("plan") over a pre-sized
// HBM arena.  Two plans are built from the same topology code:
//   offline  : tensors are [max_frames][F][C]; the previous-frame tap of a conv is the same tensor one
//              frame earlier (zero at t = 0); scratch tensors of one nested sub-U-Net are recycled by the next.
//   streaming: tensors are [max_streams][F][C] x 2 (ping-pong by step parity); the previous-frame tap is the
//              other parity's buffer, so the reference's 104 conv-history tensors are simply last step's
//              activations and never copied.  LSTM h/c are [max_streams][21], updated in place.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/nunet_b200.h"
#include "conv_simt.cuh"
#include "conv_tc3.cuh"
#include "framing.cuh"
#include "misc_kernels.cuh"
#include "sh16_kernels.cuh"
#include "lstm_kernels.cuh"
#include "ddb_kernels.cuh"
#include "hybrid_kernels.cuh"
#include "fused_tail.cuh"
#include "ddb_fused.cuh"

namespace nunet {

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] static void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define CUDA_OK(expr)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) fail(NUNET_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------ weight blob
struct Arr {
    std::vector<int> dims;
    const float* data = nullptr;
    size_t n = 0;
};

struct Blob {
    std::map<std::string, Arr> m;
    int variant = 0;
    void parse(const void* blob, size_t bytes) {
        const uint8_t* p = static_cast<const uint8_t*>(blob);
        if (bytes < 16 || memcmp(p, "NUNETW01", 8) != 0) fail(NUNET_EINVAL, "weight blob: bad magic");
        uint32_t cnt, var;
        memcpy(&cnt, p + 8, 4);
        memcpy(&var, p + 12, 4);
        variant = (int)var;
        const size_t ENTRY = 96;
        const size_t base = 16 + (size_t)cnt * ENTRY;
        if (bytes < base) fail(NUNET_EINVAL, "weight blob: truncated table");
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint8_t* e = p + 16 + i * ENTRY;
            char name[65];
            memcpy(name, e, 64);
            name[64] = 0;
            uint32_t ndim, d[4];
            uint64_t off;
            memcpy(&ndim, e + 64, 4);
            memcpy(d, e + 68, 16);
            memcpy(&off, e + 84, 8);
            if (ndim > 4) fail(NUNET_EINVAL, "weight blob: %s has rank %u", name, ndim);
            Arr a;
            a.n = 1;
            const size_t avail = (bytes - base) / 4;          // floats behind the table
            for (uint32_t k = 0; k < ndim; ++k) {
                if (d[k] > 0x7fffffffu) fail(NUNET_EINVAL, "weight blob: %s has a dimension of %u", name, d[k]);
                a.dims.push_back((int)d[k]);
                if (d[k] != 0 && a.n > avail / d[k]) fail(NUNET_EINVAL, "weight blob: %s out of range", name);
                a.n *= d[k];
            }
            if (off > avail || a.n > avail - off) fail(NUNET_EINVAL, "weight blob: %s out of range", name);
            a.data = reinterpret_cast<const float*>(p + base + off * 4);
            m[name] = a;
        }
    }
    const Arr& get(const std::string& name, std::initializer_list<int> dims = {}) const {
        auto it = m.find(name);
        if (it == m.end()) fail(NUNET_EINVAL, "weight blob: missing tensor %s", name.c_str());
        if (dims.size()) {
            std::vector<int> want(dims);
            if (want != it->second.dims) fail(NUNET_EINVAL, "weight blob: %s has an unexpected shape", name.c_str());
        }
        return it->second;
    }
};

// All packed parameters live in one device allocation; layers hold float offsets into it.
struct ParamPool {
    std::vector<float> host;
    float* dev = nullptr;
    size_t add(const float* p, size_t n) {
        size_t off = (host.size() + 63) & ~size_t(63);
        host.resize(off + n);
        memcpy(host.data() + off, p, n * sizeof(float));
        return off;
    }
    size_t add(const std::vector<float>& v) { return add(v.data(), v.size()); }
    void upload() {
        CUDA_OK(cudaMalloc(&dev, host.size() * sizeof(float) + 256));
        CUDA_OK(cudaMemcpy(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    const float* at(size_t off) const { return dev + off; }
    ~ParamPool() {
        if (dev) cudaFree(dev);
    }
};

struct ConvLayer {
    size_t w = 0, bias = 0, gamma = 0, beta = 0, alpha = 0;
    int CA = 0, CB = 0, COUT = 0, KT = 1, KF = 1, padl = 0, stride = 1, epi = EPI_LN;
    // split-half tensor-core path (conv_tc3.cuh): fp16 hi/lo weights, columns in output-channel order
    size_t w3 = 0, b3 = 0;   // float offsets into the pool (w3 holds raw halves)
    size_t wfz = 0;          // the same scaled hi / lo weights in mma.sync fragment order (fused_tail.cuh)
    int N3 = 0, PC3 = 0, nhalf3 = 1;
    float wscale_inv = 1.0f;
};
struct MlpLayer {
    size_t k0, b0, k1, b1;
};
struct LstmLayer {
    size_t wk, wk4, wr, wb, dk, db;
    int D;
};
struct DdbLayer {   // one dilated dense block: float offsets into the pool
    int C = 0;
    size_t w_in, b_in, a_in, w_out, b_out, a_out;
    size_t w0[6], b0[6], w1[6], b1[6], gamma[6], beta[6], alpha[6];
};
// int8-hybrid variant (hybrid_kernels.cuh): float offsets into the pool; int8 / int32 data is stored as raw bytes
struct HqConvLayer {
    size_t w = 0, wtap = 0, wscale = 0, bias = 0, gamma = 0, beta = 0, alpha = 0;
    int Ct = 0, Cout = 0, KT = 1, KF = 1, padl = 0, stride = 1, epi = HQ_BIAS;
};
struct HqMlpLayer {
    size_t k0, k1, s0, s1, b0, b1;
};
struct HqLstmLayer {
    size_t wk, wr, wd, wd_f, bk, bd;
    float sk, sr, sd;
    bool dense_q;
    int D;
};
struct VecLayer {   // input_layer / out_conv
    size_t w, b, gamma, beta, alpha;
};

static int conv_cn(int COUT) { return COUT == 32 ? 4 : 8; }

// kernel: logical [taps][Cin][COUT] -> column-permuted copy for conv_unit_kernel
static std::vector<float> permute_cols(const std::vector<float>& k, int rows, int COUT) {
    std::vector<float> out(k.size());
    const int CN = conv_cn(COUT);
    for (int r = 0; r < rows; ++r)
        for (int q = 0; q < COUT; ++q) out[(size_t)r * COUT + q] = k[(size_t)r * COUT + conv_col_to_channel(q, COUT, CN)];
    return out;
}

// Packed column n of half h  <->  logical conv channel (the sub-pixel shuffles of models/proposed.py:227-237 are
// folded into this order so that every output pixel's channels are contiguous columns).
static int tc3_col_to_channel(int epi, int h, int n) {
    if (epi == EPI_SHUF32) return 2 * (n % 32) + n / 32;             // column j*32+i  <-> channel 2i+j (pixel 2f+j)
    if (epi == EPI_SHUF64) return 64 * h + 2 * (n % 32) + n / 32;    // column 32j+i of half h <-> channel 64h+2i+j
    return n;
}

// logical [taps][Cin][COUT] -> resident fp16 hi/lo operand image of conv_tc3_kernel:
// [half][phase][tap][chunk 2][hi N | lo N][8 halves], scaled by 2^s (s chosen so that max |w| lands in [2^14, 2^15)).
static std::vector<float> pack_tc3(const std::vector<float>& k, int taps, int Cin, int COUT, int epi, int nhalf, int N,
                                   float* scale_inv) {
    float mx = 0.f;
    for (float w : k) mx = std::max(mx, std::fabs(w));
    int e = 0;
    if (mx > 0.f) e = 14 - (int)std::floor(std::log2((double)mx));
    e = std::max(-20, std::min(40, e));
    const float scale = std::ldexp(1.0f, e);
    *scale_inv = std::ldexp(1.0f, -e);
    const int nph = Cin / T3_KCH;
    std::vector<__half> out((size_t)nhalf * nph * taps * N * 32);
    for (int h = 0; h < nhalf; ++h)
        for (int ph = 0; ph < nph; ++ph)
            for (int tap = 0; tap < taps; ++tap) {
                const size_t stage = (((size_t)h * nph + ph) * taps + tap) * ((size_t)N * 32);
                for (int ch = 0; ch < 2; ++ch)
                    for (int n = 0; n < N; ++n)
                        for (int e8 = 0; e8 < 8; ++e8) {
                            const int ci = ph * T3_KCH + ch * 8 + e8;
                            const int co = tc3_col_to_channel(epi, h, n);
                            const float w = k[((size_t)tap * Cin + ci) * COUT + co] * scale;
                            const __half hi = __float2half_rn(w);
                            const __half lo = __float2half_rn(w - __half2float(hi));
                            out[stage + ((size_t)ch * 2 * N + n) * 8 + e8] = hi;
                            out[stage + ((size_t)ch * 2 * N + N + n) * 8 + e8] = lo;
                        }
            }
    std::vector<float> raw(out.size() / 2);
    memcpy(raw.data(), out.data(), out.size() * sizeof(__half));
    return raw;
}

// logical [taps][Cin][COUT] -> B fragments of mma.sync.m16n8k16 for fused_tail.cuh: [k-step = (tap, 16-channel group)][8-column tile]
// [lane] x {b_hi k 2q..2q+1, b_hi k 2q+8..9, b_lo same} with n = 8 tile + lane / 4, q = lane % 4; columns in the packed order of
// tc3_col_to_channel (all COUT columns: the two 64-column halves of a 128-channel unit follow each other); same power-of-two scale
// as pack_tc3.
static std::vector<float> pack_fz(const std::vector<float>& k, int taps, int Cin, int COUT, int epi, float scale_inv) {
    const float scale = 1.0f / scale_inv;
    const int cgn = Cin / 16, NT = COUT / 8, nhalf = (COUT == 128) ? 2 : 1, Nh = COUT / nhalf;
    std::vector<uint32_t> out((size_t)taps * cgn * NT * 32 * 4);
    auto halves = [&](int tap, int ci, int n, __half& hi, __half& lo) {
        const int co = tc3_col_to_channel(epi, n / Nh, n % Nh);
        const float w = k[((size_t)tap * Cin + ci) * COUT + co] * scale;
        hi = __float2half_rn(w);
        lo = __float2half_rn(w - __half2float(hi));
    };
    auto pack2 = [](__half a, __half b) {
        uint16_t x, y;
        memcpy(&x, &a, 2);
        memcpy(&y, &b, 2);
        return (uint32_t)x | ((uint32_t)y << 16);
    };
    // LayerNorm groups (the PC channels of one output pixel) are stored one after the other so that a group -- or a run of
    // groups -- is one contiguous weight chunk for the kernel's shared-memory staging
    const int PC = (epi == EPI_SHUF32) ? 32 : (epi == EPI_SHUF64) ? 64 : COUT, NTG = PC / 8;
    for (int tap = 0; tap < taps; ++tap)
        for (int cg = 0; cg < cgn; ++cg)
            for (int nt = 0; nt < NT; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int n = nt * 8 + (lane >> 2), q = lane & 3;
                    __half h[4], l[4];
                    const int kk[4] = {2 * q, 2 * q + 1, 2 * q + 8, 2 * q + 9};
                    for (int i = 0; i < 4; ++i) halves(tap, cg * 16 + kk[i], n, h[i], l[i]);
                    const int px = nt / NTG, ntl = nt % NTG;
                    uint32_t* o = &out[(((((size_t)px * taps + tap) * cgn + cg) * NTG + ntl) * 32 + lane) * 4];
                    o[0] = pack2(h[0], h[1]);
                    o[1] = pack2(h[2], h[3]);
                    o[2] = pack2(l[0], l[1]);
                    o[3] = pack2(l[2], l[3]);
                }
    std::vector<float> raw(out.size());
    memcpy(raw.data(), out.data(), out.size() * 4);
    return raw;
}

// ------------------------------------------------------------------------------------------ tensors / plans
struct Ten {
    int F = 0, C = 0;
    size_t off[2] = {0, 0};   // float offset per unit (frame / stream) inside the arena; [1] only when ping-ponged
    bool scratch = false;     // offline: lives on the recyclable stack (offset fixed up by Plan::finalize)
    bool pingpong = false;    // streaming: two copies selected by step parity
    bool sh = false;          // holds split-half records (sh16 plans) rather than floats
    bool eo = false;          // sh16: bins stored [even | odd] inside every plane
    Ten* twin = nullptr;      // offline sh16: a second copy in [even | odd] order written by the same producer, for stride-2 readers
    std::string name;
    size_t numel() const { return (size_t)F * C; }
};

struct Run {   // arguments of one forward / step call
    int B = 0, T = 0;
    int parity = 0;
    cudaStream_t st = nullptr;
    const float* mag_in = nullptr;   // [B*T][256]
    float* est_out = nullptr;        // [B*T][est_stride], written at +est_off
    int est_stride = 256, est_off = 0;
    int ring_pos = 0;
    int step = 0;                    // streaming: absolute step counter (DDB rings)
    // offline calls cut into time chunks (T frames of every clip per chunk, history carried between chunks):
    bool use_carry = false;          // this chunk continues its clips: row t = -1, LSTM h / c and the TA window come from the carry arena
    bool save_carry = false;         // another chunk follows: leave them there
    int t0 = 0;                      // clip-relative index of the chunk's first frame
};

struct Engine;
using Op = std::function<void(Engine&, const Run&)>;

// A plan = tensors laid out in one arena + the launch list.  Every tensor is a dense [cap][F*C] block at
// arena + off*cap.  Offline plans keep two bump stacks: `persistent` (second-level skips, block outputs) and
// `scratch` (recycled by each nested sub-U-Net); streaming plans keep everything and ping-pong activations.
struct Plan {
    bool streaming = false;
    bool sh16 = false;           // activations are split-half records (conv_tc3.cuh) instead of fp32
    int cap = 0;                 // frames or streams
    float* arena = nullptr;
    size_t unit_floats = 0;      // floats per frame / stream
    size_t ptop = 0, stop = 0, shigh = 0;
    std::vector<std::unique_ptr<Ten>> tens;
    std::map<std::string, Ten*> named;
    std::vector<Op> ops;
    // streaming state: reference name (without _prev/_cur) -> up to two tensors concatenated on channels
    struct StateRef {
        std::string name;
        Ten* a = nullptr;
        Ten* b = nullptr;
        bool is_lstm = false;    // a = the [21] h or c vector
        // DDB layer k (converter_nunet_tls.py:373-411): the last d rows of cat[out_{k-1}, .., out_0], gathered from
        // the per-stream rings of out_0 .. out_{k-1}
        int ddb_k = 0, ddb_d = 0;
        Ten* rings[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    };
    std::vector<StateRef> states;
    std::vector<Ten*> rings;
    // offline: carried state of time-chunked calls, [carry_cap clips] per region
    float* carry = nullptr;
    size_t ctop = 0;             // floats per clip
    int carry_cap = 0;
    // (tensor, carry slot) pairs the sub-U-Net being built must save.  A slot belongs to ONE sub-U-Net: the spconv outputs of
    // an encoder block feed causal convs in the encoder (next spconv) and in the paired decoder (second-level skip), and the
    // decoder must still see the previous chunk's row after the encoder has saved this chunk's.
    std::vector<std::pair<Ten*, size_t>> pending_save;
    size_t carry_alloc(size_t n) {
        const size_t off = ctop;
        ctop += (n + 63) & ~size_t(63);
        return off;
    }
    size_t want_carry(Ten* t) {
        for (auto& pr : pending_save)
            if (pr.first == t) return pr.second;
        pending_save.emplace_back(t, carry_alloc(t->numel()));
        return pending_save.back().second;
    }
    float* carry_at(size_t coff) const { return carry + coff * (size_t)carry_cap; }

    Ten* make(const std::string& name, int F, int C, bool persistent, bool pingpong = true) {
        auto t = std::make_unique<Ten>();
        t->F = F;
        t->C = C;
        t->name = name;
        const size_t n = ((size_t)F * C + 63) & ~size_t(63);
        if (streaming) {
            t->pingpong = pingpong;
            t->off[0] = ptop;
            ptop += n;
            if (pingpong) {
                t->off[1] = ptop;
                ptop += n;
            } else {
                t->off[1] = t->off[0];
            }
        } else if (persistent) {
            t->off[0] = t->off[1] = ptop;
            ptop += n;
        } else {
            t->scratch = true;
            t->off[0] = t->off[1] = stop;
            stop += n;
            if (stop > shigh) shigh = stop;
        }
        Ten* raw = t.get();
        tens.push_back(std::move(t));
        if (!name.empty()) named[name] = raw;
        return raw;
    }
    size_t mark() const { return stop; }
    void release(size_t m) { stop = m; }
    void finalize() {
        for (auto& t : tens)
            if (t->scratch) {
                t->off[0] += ptop;
                t->off[1] += ptop;
            }
        unit_floats = ptop + shigh;
    }
    float* ptr(size_t off) const { return arena + off * (size_t)cap; }
    // unit0: first unit (stream) the ops enqueued next work on -- lets one step be enqueued as several independent chains
    // over disjoint stream ranges (CUDA-graph branches); 0 everywhere else
    int unit0 = 0;
    float* cur(const Ten* t, int parity) const { return ptr(t->off[parity & 1]) + (size_t)unit0 * t->numel(); }
    float* prev(const Ten* t, int parity) const {
        return streaming ? ptr(t->off[(parity ^ 1) & 1]) + (size_t)unit0 * t->numel() : nullptr;
    }
    ~Plan() {
        if (arena) cudaFree(arena);
        if (carry) cudaFree(carry);
    }
};

static const char* ENC_NAMES[6] = {"msfe6_en", "msfe5_en", "msfe4_en", "msfe4_en2", "msfe4_en3", "msfe3_en"};
static const int ENC_F0[6] = {256, 128, 64, 32, 16, 8};
static const int ENC_N[6] = {6, 5, 4, 4, 4, 3};
static const char* DEC_NAMES[6] = {"msfe3_de", "msfe4_de", "msfe4_de2", "msfe4_de3", "msfe5_de", "msfe6_de"};
static const int DEC_F0[6] = {8, 16, 32, 64, 128, 256};
static const int DEC_N[6] = {3, 4, 4, 4, 5, 6};
static const char* DOWN_NAMES[6] = {"msfe6_down_sampling", "msfe5_down_sampling", "msfe4_down_sampling",
                                    "msfe4_down_sampling2", "msfe4_down_sampling3", "msfe3_down_sampling"};
static const char* UP_NAMES[6] = {"msfe3_upsampling", "msfe4_upsampling", "msfe4_upsampling2",
                                  "msfe4_upsampling3", "msfe5_upsampling", "msfe6_upsampling"};

// converter_proposed.py:26-187 naming: block 'msfe4_en2' -> conv history 'msfe4_ee2', spconv history 'msfe4_ed2'
static void state_prefixes(const std::string& block, std::string& pc, std::string& ps) {
    const size_t us = block.find('_');
    const std::string head = block.substr(0, us), tail = block.substr(us + 1);
    const std::string side = tail.substr(0, 2), idx = tail.substr(2);
    const char a = side == "en" ? 'e' : 'd';
    pc = head + "_" + a + "e" + idx;
    ps = head + "_" + a + "d" + idx;
}

struct Engine {
    nunet_config cfg{};   // (cfg.chunk_frames: forced time-chunk length of offline calls, 0 = only when a clip exceeds max_frames)
    Blob blob;
    ParamPool pool;
    std::map<std::string, ConvLayer> convs;
    std::map<std::string, MlpLayer> mlps;
    std::map<std::string, LstmLayer> lstms;
    std::map<std::string, DdbLayer> ddbs;
    VecLayer in_layer{}, out_layer{};
    std::map<std::string, HqConvLayer> hconvs;
    std::map<std::string, HqMlpLayer> hmlps;
    std::map<std::string, HqLstmLayer> hlstms;
    std::map<std::string, std::pair<size_t, size_t>> hups;   // up_sampling: dequantised kernel [k][ci][co], bias
    Ten *hq_qbuf = nullptr, *hq_qp = nullptr;
    // streaming plans: the layers of a nested sub-U-Net that work on <= 32 bins are collected into one fused_tail_kernel launch
    struct FzGroup {
        std::vector<std::function<void(Engine&, const Run&, FzLayer&)>> fill;
        double alg_bytes = 0.0;
    };
    std::unique_ptr<FzGroup> fz_open;     // group being collected while the plan is built
    int fz_dbg = 0;                       // NUNET_FZ_DBG (experiments)
    int stream_fuse = 0;                  // NUNET_STREAM_FUSE=1 (debug knob): fused_tail_kernel for the <= 32-bin layers.  Off by default:
                                          // measured at 1 .. 2048 streams it is 0-15 % SLOWER than one conv_tc3 launch per layer (94 vs 174
                                          // launches per step) -- legacy mma.sync sustains ~290 FMA/clk/SM on B200, 28x below tcgen05, so the
                                          // fused chain is bound by the tensor path it uses (profiles/r2_fused_tail_*.txt)
    size_t tw_off = 0, win_off = 0, win_stream_off = 0, inv_win_off = 0;

    Plan offline, stream;
    // offline extras (per frame offsets)
    Ten *o_mag = nullptr, *o_ph = nullptr, *o_est = nullptr, *o_frames = nullptr;
    // streaming extras (per stream offsets)
    Ten *s_mag = nullptr, *s_ph = nullptr, *s_est = nullptr, *s_inbuf = nullptr, *s_outbuf = nullptr;
    int stream_parity = 0;   // parity of the most recent step (its buffers hold the history)
    int stream_steps = 0;
    long long state_gen = 1;   // bumped whenever the resident history changes (step, reset, import): nunet_state_generation

    cudaStream_t own_stream = nullptr;
    float *h_in = nullptr, *h_out = nullptr;   // device staging for the *_host calls
    size_t h_in_cap = 0, h_out_cap = 0;
    int launches = 0;
    int last_B = 0, last_T = 0;
    int num_sms = 148;
    bool lstm_stream = true; // NUNET_LSTM_STREAM=0 (experiments): streaming steps on lstm_block_kernel (one CTA per stream)
    bool ctfa_gate4 = true;  // NUNET_CTFA_GATE4=0 (experiments): the one-frame-per-warp gate kernel
    bool tc3_twin = true;    // NUNET_TC3_TWIN=0: no [even | odd] twins (stride-2 units then use strided boxes over bin-ordered sources)
    bool no_recycle = false; // NUNET_NO_RECYCLE (tests/layer_report.py): every offline tensor keeps its own storage
    bool use_tc = true;     // NUNET_CONV=simt forces the FP32 SIMT units everywhere
    int tc3_fence_mode = 0;  // NUNET_TC3_FENCE
    int tc3_dbg = 0;         // NUNET_TC3_DBG (experiments)
    unsigned long long* tc3_timing_buf = nullptr;   // NUNET_TC3_TIMING=1 (experiments): per-role cycle counters of CTA 0
    bool tc3_pdl = true;     // NUNET_TC3_PDL=0: plain stream order between consecutive conv kernels
    bool stream_tc3 = true;  // NUNET_STREAM_CONV=simt keeps the streaming plan on the FP32 SIMT units
    int tc3_force_mt = 0;    // NUNET_TC3_MT (experiments)
    int tc3_tma = 2;         // NUNET_TC3_TMA: 0 = cp.async loaders everywhere, 1 = 1-D bulk copies per frame-row segment (F >= 32),
                             // 2 (default) = one tensor-map box per tile image where the source order allows, else as 1
    int tc3_box_strided = 1; // NUNET_TC3_BOX_STRIDED=0 (experiments): stride-2 units over bin-ordered sources stay on the slot-table loader
    int tc3_row_tiles = 1;   // NUNET_TC3_ROW_TILES=0: flat 128-position tiles for every unit
    int tc3_pair = 1;        // NUNET_TC3_PAIR=0: 128-channel units run as two independent CTAs per tile instead of cta_group::2 pairs
    int tc3_pair_minf = 4;   // NUNET_TC3_PAIR_MINF (experiments)
    bool tc3_ld_rr = true;   // NUNET_TC3_LD_RR=0 (experiments): one loader warp per plane in every unit
    int tc3_box_minf = 4;    // NUNET_TC3_BOX_MINF (experiments): smallest F_conv of a unit that uses tensor-map boxes
    int tc3_tma_minf = 32;   // NUNET_TC3_TMA_MINF (experiments): smallest F_in of a stride-1 unit that uses bulk copies
    int tc3_cluster = 0;     // NUNET_TC3_CLUSTER=1: the two CTAs of a 128-channel unit form a cluster and multicast their bulk copies
                             // (one L2 read feeds both); measured neutral on B200, kept as an option
    // per-launch profiling (bench.py roofline leg): one CUDA event after every launch on the launching stream
    bool prof_on = false;
    cudaStream_t prof_stream = nullptr;
    std::string cur_op;
    struct ProfEntry {
        std::string name;
        double alg_bytes;
        cudaEvent_t ev;
    };
    std::vector<ProfEntry> prof;
    cudaEvent_t prof_start = nullptr;
    // A handle owns ONE arena: calls issued on different streams are ordered through this event so that they
    // never overlap on the device.
    cudaEvent_t last_done = nullptr;
    void order_begin(cudaStream_t st) {
        if (!last_done) CUDA_OK(cudaEventCreateWithFlags(&last_done, cudaEventDisableTiming));
        else CUDA_OK(cudaStreamWaitEvent(st, last_done, 0));
    }
    void order_end(cudaStream_t st) { CUDA_OK(cudaEventRecord(last_done, st)); }

    ~Engine() {
        if (h_in) cudaFree(h_in);
        if (h_out) cudaFree(h_out);
        if (last_done) cudaEventDestroy(last_done);
        for (auto& pe : prof) cudaEventDestroy(pe.ev);
        if (prof_start) cudaEventDestroy(prof_start);
        if (own_stream) cudaStreamDestroy(own_stream);
        for (auto& kv : step_graphs) cudaGraphExecDestroy(kv.second.exec);
        for (int i = 0; i < 4; ++i) {
            if (cap_stream[i]) cudaStreamDestroy(cap_stream[i]);
            if (cap_event[i]) cudaEventDestroy(cap_event[i]);
        }
    }

    // -------------------------------------------------------------------------------- parameter packing
    size_t add_arr(const std::string& name, std::initializer_list<int> dims = {}) {
        const Arr& a = blob.get(name, dims);
        return pool.add(a.data, a.n);
    }

    void add_conv(const std::string& role, int CA, int CB, int COUT, int KT, int KF, int padl, int stride, int epi) {
        const Arr& k = blob.get(role + "/kernel", {KT, KF, CA + CB, COUT});
        ConvLayer L;
        L.CA = CA; L.CB = CB; L.COUT = COUT; L.KT = KT; L.KF = KF; L.padl = padl; L.stride = stride; L.epi = epi;
        std::vector<float> kv(k.data, k.data + k.n);
        L.w = pool.add(permute_cols(kv, KT * KF * (CA + CB), COUT));
        L.bias = add_arr(role + "/bias", {COUT});
        {
            const Arr& bb = blob.get(role + "/bias", {COUT});
            add_tc3(L, kv, std::vector<float>(bb.data, bb.data + bb.n));
        }
        if (epi != EPI_BIAS) {
            const int cln = (epi == EPI_LN) ? COUT : COUT / 2;
            L.gamma = add_arr(role + "/gamma", {cln});
            L.beta = add_arr(role + "/beta", {cln});
            L.alpha = add_arr(role + "/alpha", {1});
        }
        convs[role] = L;
    }

    void add_tc3(ConvLayer& L, const std::vector<float>& kv, const std::vector<float>& bias) {
        L.nhalf3 = (L.COUT == 128) ? 2 : 1;
        L.N3 = L.COUT / L.nhalf3;
        L.PC3 = (L.epi == EPI_SHUF32) ? 32 : L.N3;
        L.w3 = pool.add(pack_tc3(kv, L.KT * L.KF, L.CA + L.CB, L.COUT, L.epi, L.nhalf3, L.N3, &L.wscale_inv));
        std::vector<float> b3((size_t)L.COUT);
        for (int h = 0; h < L.nhalf3; ++h)
            for (int n = 0; n < L.N3; ++n) b3[(size_t)h * L.N3 + n] = bias[tc3_col_to_channel(L.epi, h, n)];
        L.b3 = pool.add(b3);
        if (stream_fuse && cfg.max_streams > 0 && (L.CA + L.CB) % 16 == 0) L.wfz = pool.add(pack_fz(kv, L.KT * L.KF, L.CA + L.CB, L.COUT, L.epi, L.wscale_inv));
    }

    // up_sampling (Conv2DTranspose (1,3) stride (1,2) 'same', models/proposed.py:260) followed by the decoder
    // block's inconv (1x1 + LN + PReLU, :218) with nothing in between: composed into ONE conv unit with two
    // frequency taps (x[i-1], x[i]) and 2 x 64 output channels that the SHUF64 epilogue scatters to bins 2i, 2i+1.
    //   u[2i]   = x[i] W0 + x[i-1] W2 + b_up,  u[2i+1] = x[i] W1 + b_up,  z = u Win + b_in
    void add_up_in(const std::string& up, const std::string& in_role) {
        const Arr& ku = blob.get(up + "/kernel", {1, 3, 128, 128});      // (kh, kw, Cout, Cin)
        const Arr& bu = blob.get(up + "/bias", {128});
        const Arr& ki = blob.get(in_role + "/kernel", {1, 1, 128, 64});  // (1,1,Cin,Cout)
        const Arr& bi = blob.get(in_role + "/bias", {64});
        // M[k][ci][m] = sum_co Wup[0,k,co,ci] * Win[co,m]
        std::vector<double> M(3 * 128 * 64, 0.0);
        for (int k = 0; k < 3; ++k)
            for (int co = 0; co < 128; ++co)
                for (int ci = 0; ci < 128; ++ci) {
                    const double wu = ku.data[((size_t)k * 128 + co) * 128 + ci];
                    const float* wi = ki.data + (size_t)co * 64;
                    double* dst = &M[((size_t)k * 128 + ci) * 64];
                    for (int m = 0; m < 64; ++m) dst[m] += wu * (double)wi[m];
                }
        std::vector<double> bz(64, 0.0);
        for (int m = 0; m < 64; ++m) {
            double s = bi.data[m];
            for (int co = 0; co < 128; ++co) s += (double)bu.data[co] * (double)ki.data[(size_t)co * 64 + m];
            bz[m] = s;
        }
        // conv channel c = 64h + 2i' + j  <->  output bin 2i+h, output channel m = 32j + i'
        std::vector<float> W(2 * 128 * 128, 0.0f), B(128, 0.0f);
        for (int c = 0; c < 128; ++c) {
            const int h = c / 64, ip = (c % 64) / 2, j = c & 1, m = 32 * j + ip;
            B[c] = (float)bz[m];
            for (int ci = 0; ci < 128; ++ci) {
                // tap kf = 0 reads x[i-1]: contributes W2 to the even bin only; tap kf = 1 reads x[i]: W0 (even), W1 (odd)
                W[((size_t)0 * 128 + ci) * 128 + c] = (h == 0) ? (float)M[((size_t)2 * 128 + ci) * 64 + m] : 0.0f;
                W[((size_t)1 * 128 + ci) * 128 + c] = (float)M[((size_t)(h == 0 ? 0 : 1) * 128 + ci) * 64 + m];
            }
        }
        ConvLayer L;
        L.CA = 64; L.CB = 64; L.COUT = 128; L.KT = 1; L.KF = 2; L.padl = 1; L.stride = 1; L.epi = EPI_SHUF64;
        L.w = pool.add(permute_cols(W, 2 * 128, 128));
        L.bias = pool.add(B);
        add_tc3(L, W, B);
        L.gamma = add_arr(in_role + "/gamma", {64});
        L.beta = add_arr(in_role + "/beta", {64});
        L.alpha = add_arr(in_role + "/alpha", {1});
        convs[in_role] = L;
    }

    void add_mlp(const std::string& role) {
        MlpLayer m;
        m.k0 = add_arr(role + "/kernel0", {64, 16});
        m.b0 = add_arr(role + "/bias0", {16});
        m.k1 = add_arr(role + "/kernel1", {16, 64});
        m.b1 = add_arr(role + "/bias1", {64});
        mlps[role] = m;
    }
    void add_lstm(const std::string& lstm, const std::string& dense, int D) {
        LstmLayer l;
        l.D = D;
        l.wk = add_arr(lstm + "/kernel", {D, LSTM_GATES});
        {   // [D/4][84][4]: four consecutive k of one gate in one 16-byte word (lstm_kernels.cuh)
            const Arr& k = blob.get(lstm + "/kernel", {D, LSTM_GATES});
            std::vector<float> k4((size_t)D * LSTM_GATES);
            for (int kk = 0; kk < D; ++kk)
                for (int g = 0; g < LSTM_GATES; ++g) k4[((size_t)(kk >> 2) * LSTM_GATES + g) * 4 + (kk & 3)] = k.data[(size_t)kk * LSTM_GATES + g];
            l.wk4 = pool.add(k4);
        }
        l.wr = add_arr(lstm + "/recurrent_kernel", {LSTM_UNITS, LSTM_GATES});
        l.wb = add_arr(lstm + "/bias", {LSTM_GATES});
        l.dk = add_arr(dense + "/kernel", {LSTM_UNITS, D});
        l.db = add_arr(dense + "/bias", {D});
        lstms[lstm] = l;
    }

    // role = "<block>_ddb" or "ddb": in / 1..6 / out (weights.py: ddb_weights_from_tflite)
    void add_ddb(const std::string& role, int C) {
        const int h = C / 2;
        DdbLayer d;
        d.C = C;
        d.w_in = add_arr(role + "_in/kernel", {2, 3, C, h});
        d.b_in = add_arr(role + "_in/bias", {h});
        d.a_in = add_arr(role + "_in/alpha", {1});
        for (int k = 1; k <= 6; ++k) {
            const std::string r = role + "_" + std::to_string(k);
            d.w0[k - 1] = add_arr(r + "/kernel0", {2, 3, k, h});
            d.b0[k - 1] = add_arr(r + "/bias0", {h});
            d.w1[k - 1] = add_arr(r + "/kernel1", {h, h});
            d.b1[k - 1] = add_arr(r + "/bias1", {h});
            d.gamma[k - 1] = add_arr(r + "/gamma", {h});
            d.beta[k - 1] = add_arr(r + "/beta", {h});
            d.alpha[k - 1] = add_arr(r + "/alpha", {1});
        }
        d.w_out = add_arr(role + "_out/kernel", {2, 3, h, C});
        d.b_out = add_arr(role + "_out/bias", {C});
        d.a_out = add_arr(role + "_out/alpha", {1});
        ddbs[role] = d;
    }
    bool is_ddb() const { return cfg.variant == NUNET_VARIANT_DDB; }
    bool is_hybrid() const { return cfg.variant == NUNET_VARIANT_LSTM_HYBRID; }

    // ---- int8-hybrid variant: int8 weights (stored in the blob as exact float values, TFLite layouts) + scales
    size_t add_bytes(const std::vector<int8_t>& v) {
        std::vector<float> raw((v.size() + 3) / 4, 0.0f);
        memcpy(raw.data(), v.data(), v.size());
        return pool.add(raw);
    }
    size_t add_ints(const std::vector<int>& v) {
        std::vector<float> raw(v.size());
        memcpy(raw.data(), v.data(), v.size() * 4);
        return pool.add(raw);
    }
    static int8_t q8(float f) {
        if (f != std::floor(f) || f < -128.f || f > 127.f) fail(NUNET_EINVAL, "weight blob: a quantised tensor holds %g", f);
        return (int8_t)f;
    }
    // role/kernel_q [Cout][KT][KF][Ct] + role/kernel_scale [Cout]
    void add_hq_conv(const std::string& role, int Ct, int Cout, int KT, int KF, int padl, int stride, int epi) {
        const Arr& q = blob.get(role + "/kernel_q", {Cout, KT, KF, Ct});
        HqConvLayer L;
        L.Ct = Ct; L.Cout = Cout; L.KT = KT; L.KF = KF; L.padl = padl; L.stride = stride; L.epi = epi;
        const int taps = KT * KF, C4 = Ct / 4;
        std::vector<int8_t> w((size_t)taps * Ct * Cout);
        std::vector<int> wtap((size_t)taps * Cout, 0);
        for (int c = 0; c < Cout; ++c)
            for (int t = 0; t < taps; ++t)
                for (int ci = 0; ci < Ct; ++ci) {
                    const int8_t v = q8(q.data[((size_t)c * taps + t) * Ct + ci]);
                    w[(((size_t)t * C4 + ci / 4) * Cout + c) * 4 + (ci & 3)] = v;       // word (tap, ci/4, c), byte ci % 4
                    wtap[(size_t)t * Cout + c] += v;
                }
        L.w = add_bytes(w);
        L.wtap = add_ints(wtap);
        L.wscale = add_arr(role + "/kernel_scale", {Cout});
        L.bias = add_arr(role + "/bias", {Cout});
        if (epi != HQ_BIAS) {
            const int cln = (epi == HQ_LN) ? Cout : Cout / 2;
            L.gamma = add_arr(role + "/gamma", {cln});
            L.beta = add_arr(role + "/beta", {cln});
            L.alpha = add_arr(role + "/alpha", {1});
        }
        hconvs[role] = L;
    }
    void add_hq_mlp(const std::string& role) {
        HqMlpLayer m;
        auto bytes = [&](const std::string& n, int co, int ci) {
            const Arr& q = blob.get(n, {co, 1, 1, ci});
            std::vector<int8_t> v(q.n);
            for (size_t i = 0; i < q.n; ++i) v[i] = q8(q.data[i]);
            return add_bytes(v);
        };
        m.k0 = bytes(role + "/kernel0_q", 16, 64);
        m.k1 = bytes(role + "/kernel1_q", 64, 16);
        m.s0 = add_arr(role + "/kernel0_scale", {16});
        m.s1 = add_arr(role + "/kernel1_scale", {64});
        m.b0 = add_arr(role + "/bias0", {16});
        m.b1 = add_arr(role + "/bias1", {64});
        hmlps[role] = m;
    }
    void add_hq_lstm(const std::string& lstm, const std::string& dense, int D) {
        HqLstmLayer l{};
        l.D = D;
        auto bytes = [&](const std::string& n, int a, int b) {
            const Arr& q = blob.get(n, {a, b});
            std::vector<int8_t> v(q.n);
            for (size_t i = 0; i < q.n; ++i) v[i] = q8(q.data[i]);
            return add_bytes(v);
        };
        l.wk = bytes(lstm + "/kernel_q", LSTM_GATES, D);
        l.wr = bytes(lstm + "/recurrent_kernel_q", LSTM_GATES, LSTM_UNITS);
        l.sk = blob.get(lstm + "/kernel_scale", {1}).data[0];
        l.sr = blob.get(lstm + "/recurrent_kernel_scale", {1}).data[0];
        l.bk = add_arr(lstm + "/bias", {LSTM_GATES});
        l.bd = add_arr(dense + "/bias", {D});
        l.dense_q = blob.m.count(dense + "/kernel_q") != 0;
        if (l.dense_q) {
            l.wd = bytes(dense + "/kernel_q", D, LSTM_UNITS);
            l.sd = blob.get(dense + "/kernel_scale", {1}).data[0];
        } else {
            l.wd_f = add_arr(dense + "/kernel", {LSTM_UNITS, D});
        }
        hlstms[lstm] = l;
    }
    void add_hq_up(const std::string& up) {
        const Arr& ku = blob.get(up + "/kernel", {1, 3, 128, 128});      // (kh, kw, Cout, Cin), dequantised
        std::vector<float> w((size_t)3 * 128 * 128);
        for (int k = 0; k < 3; ++k)
            for (int co = 0; co < 128; ++co)
                for (int ci = 0; ci < 128; ++ci) w[((size_t)k * 128 + ci) * 128 + co] = ku.data[((size_t)k * 128 + co) * 128 + ci];
        const size_t wo = pool.add(w);
        hups[up] = {wo, add_arr(up + "/bias", {128})};
    }
    void pack_params_hybrid() {
        in_layer.w = add_arr("input_layer/kernel", {1, 1, 1, 64});
        in_layer.b = add_arr("input_layer/bias", {64});
        in_layer.gamma = add_arr("input_layer/gamma", {64});
        in_layer.beta = add_arr("input_layer/beta", {64});
        in_layer.alpha = add_arr("input_layer/alpha", {1});
        out_layer.w = add_arr("out_conv/kernel", {1, 1, 64, 1});
        out_layer.b = add_arr("out_conv/bias", {1});
        for (int side = 0; side < 2; ++side)
            for (int i = 0; i < 6; ++i) {
                const std::string blk = side ? DEC_NAMES[i] : ENC_NAMES[i];
                const int n = side ? DEC_N[i] : ENC_N[i];
                const int F0 = side ? DEC_F0[i] : ENC_F0[i];
                add_hq_conv(blk + "_in", side ? 128 : 64, 64, 1, 1, 0, 1, HQ_LN);
                if (side) add_hq_up(UP_NAMES[i]);
                for (int k = 1; k <= n; ++k) add_hq_conv(blk + "_conv" + std::to_string(k), ((k == 1) ? 64 : 32) * (side ? 2 : 1), 32, 2, 3, 1, 2, HQ_LN);
                add_hq_lstm(blk + "_lstm", blk + "_dense", (F0 >> n) * 32);
                for (int k = 1; k <= n; ++k)
                    add_hq_conv(blk + "_spconv" + std::to_string(k), 64, (k == n) ? 128 : 64, 2, 3, 1, 1, (k == n) ? HQ_SHUF64 : HQ_SHUF32);
                add_hq_mlp(blk + "_ta");
                add_hq_mlp(blk + "_fa");
                if (side == 0) add_hq_conv(DOWN_NAMES[i], 64, 64, 1, 3, 0, 2, HQ_BIAS);
            }
        add_hq_lstm("lstm", "dense", 256);
        pack_framing_tables();
    }

    void pack_params() {
        if (blob.variant != cfg.variant ||
            (cfg.variant != NUNET_VARIANT_LSTM && cfg.variant != NUNET_VARIANT_DDB && cfg.variant != NUNET_VARIANT_LSTM_HYBRID))
            fail(NUNET_EINVAL, "weight blob variant %d does not match the requested variant %d", blob.variant, cfg.variant);
        if (is_hybrid()) {
            pack_params_hybrid();
            return;
        }
        in_layer.w = add_arr("input_layer/kernel", {1, 1, 1, 64});
        in_layer.b = add_arr("input_layer/bias", {64});
        in_layer.gamma = add_arr("input_layer/gamma", {64});
        in_layer.beta = add_arr("input_layer/beta", {64});
        in_layer.alpha = add_arr("input_layer/alpha", {1});
        out_layer.w = add_arr("out_conv/kernel", {1, 1, 64, 1});
        out_layer.b = add_arr("out_conv/bias", {1});
        for (int side = 0; side < 2; ++side)
            for (int i = 0; i < 6; ++i) {
                const std::string blk = side ? DEC_NAMES[i] : ENC_NAMES[i];
                const int n = side ? DEC_N[i] : ENC_N[i];
                const int F0 = side ? DEC_F0[i] : ENC_F0[i];
                if (side == 0) add_conv(blk + "_in", 64, 0, 64, 1, 1, 0, 1, EPI_LN);
                else add_up_in(UP_NAMES[i], blk + "_in");
                for (int k = 1; k <= n; ++k) {
                    int CA, CB;
                    if (side == 0) { CA = (k == 1) ? 64 : 32; CB = 0; }
                    else { CA = (k == 1) ? 64 : 32; CB = CA; }
                    add_conv(blk + "_conv" + std::to_string(k), CA, CB, 32, 2, 3, 1, 2, EPI_LN);
                }
                if (is_ddb()) add_ddb(blk + "_ddb", 32);
                else add_lstm(blk + "_lstm", blk + "_dense", (F0 >> n) * 32);
                for (int k = 1; k <= n; ++k) {
                    const bool last = (k == n);
                    add_conv(blk + "_spconv" + std::to_string(k), 32, 32, last ? 128 : 64, 2, 3, 1, 1,
                             last ? EPI_SHUF64 : EPI_SHUF32);
                }
                add_mlp(blk + "_ta");
                add_mlp(blk + "_fa");
                if (side == 0) add_conv(DOWN_NAMES[i], 64, 0, 64, 1, 3, 0, 2, EPI_BIAS);
            }
        if (is_ddb()) add_ddb("ddb", 64);
        else add_lstm("lstm", "dense", 256);

        pack_framing_tables();
    }

    void pack_framing_tables() {
        // framing tables (tf.signal.hann_window periodic; inverse_stft_window_fn(256); interpreter_proposed.py:21-26)
        std::vector<float> tw(512), win(512), wins(512), inv(512);
        const double PI = 3.14159265358979323846;
        for (int k = 0; k < 256; ++k) {
            tw[2 * k] = (float)cos(2.0 * PI * k / 512.0);
            tw[2 * k + 1] = (float)(-sin(2.0 * PI * k / 512.0));
        }
        for (int i = 0; i < 512; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / 512.0));
        wins = win;
        wins[0] = 1e-7f;
        wins[511] = 1e-7f;
        for (int i = 0; i < 512; ++i) {
            const float a = win[i], b = win[(i + 256) % 512];
            inv[i] = a / (a * a + b * b);
        }
        tw_off = pool.add(tw);
        win_off = pool.add(win);
        win_stream_off = pool.add(wins);
        inv_win_off = pool.add(inv);
    }

    FramingTables tables(bool streaming) const {
        FramingTables t;
        t.tw = reinterpret_cast<const float2*>(pool.at(tw_off));
        t.win = pool.at(streaming ? win_stream_off : win_off);
        t.inv_win = pool.at(inv_win_off);
        return t;
    }

    // -------------------------------------------------------------------------------- launches
    void check_launch(const char* what, double alg_bytes = 0.0) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) fail(NUNET_ECUDA, "launch %s (%s): %s", what, cur_op.c_str(), cudaGetErrorString(e));
        ++launches;
        if (prof_on) {
            ProfEntry pe;
            pe.name = cur_op.empty() ? std::string(what) : cur_op + ":" + what;
            pe.alg_bytes = alg_bytes;
            CUDA_OK(cudaEventCreate(&pe.ev));
            CUDA_OK(cudaEventRecord(pe.ev, prof_stream));
            prof.push_back(pe);
        }
    }
    void prof_begin(cudaStream_t st) {
        prof_stream = st;
        if (!prof_on) return;
        for (auto& pe : prof) cudaEventDestroy(pe.ev);
        prof.clear();
        if (!prof_start) CUDA_OK(cudaEventCreate(&prof_start));
        CUDA_OK(cudaEventRecord(prof_start, st));
    }

    template <int COUT, int CN, int PM, int NT, int EPI>
    void launch_conv_t(const ConvParams& p, int grid, size_t smem, cudaStream_t st) {
        static unsigned long long attr_set = 0;     // bit d: attribute set on device d (function attributes are per device)
        auto kfn = conv_unit_kernel<COUT, CN, PM, NT, EPI>;
        if (!((attr_set >> cfg.device) & 1ull)) {
            CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            attr_set |= 1ull << cfg.device;
        }
        kfn<<<grid, NT, smem, st>>>(p);
        const double frames = (double)p.B * p.T;
        check_launch("conv_unit", frames * 4.0 * ((double)p.F_in * (p.CA + p.CB) + (double)p.F_out * COUT));
    }

    // cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no link-time dependency on libcuda)
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn tc3_encode_tiled() {
        static EncodeTiledFn fn = [] {
            void* f = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
                f = nullptr;
            return reinterpret_cast<EncodeTiledFn>(f);
        }();
        return fn;
    }
    // Tensor map over one sh16 source [B][T][plane][Fp positions][16 bytes].  kind 0: rank 4 = {2 Fp words of 8 bytes, plane, t, b}
    // with a box of {2 P, 1, rows, 1}; kind 1: rank 5 = {16 words, Fp / 8 lines, plane, t, b} with a box of {16, P / 8, 1, rows, 1};
    // kind 2: rank 5 = {2 words, Fp positions, plane, t, b} traversed with a stride of two positions, box {2, 2 P, 1, rows, 1}
    // (P positions land).  Out-of-range coordinates are zero-filled, which is what makes the frequency pads and the causal
    // time pad free.
    std::map<std::tuple<const void*, int, int, int, int, int, int, int>, CUtensorMap> tc3_maps;
    CUtensorMap tc3_tensor_map(const uint8_t* base, int Fp, int planes, int T, int B, size_t row_bytes, int P, int rows, int kind) {
        const int rank = kind == 0 ? 4 : 5;
        const auto key = std::make_tuple((const void*)base, Fp, planes, T, B, P, rows, kind);
        auto it = tc3_maps.find(key);
        if (it != tc3_maps.end()) return it->second;
        CUtensorMap m;
        cuuint64_t dims[5], strides[4];
        cuuint32_t box[5], es[5] = {1, 1, 1, 1, 1};
        if (rank == 4) {
            dims[0] = 2ull * Fp; dims[1] = planes; dims[2] = T; dims[3] = B;
            strides[0] = (cuuint64_t)Fp * 16; strides[1] = row_bytes; strides[2] = (cuuint64_t)T * row_bytes;
            box[0] = 2 * P; box[1] = 1; box[2] = rows; box[3] = 1;
        } else if (kind == 1) {
            dims[0] = 16; dims[1] = Fp / 8; dims[2] = planes; dims[3] = T; dims[4] = B;
            strides[0] = 128; strides[1] = (cuuint64_t)Fp * 16; strides[2] = row_bytes; strides[3] = (cuuint64_t)T * row_bytes;
            box[0] = 16; box[1] = P / 8; box[2] = 1; box[3] = rows; box[4] = 1;
        } else {
            dims[0] = 2; dims[1] = Fp; dims[2] = planes; dims[3] = T; dims[4] = B;
            strides[0] = 16; strides[1] = (cuuint64_t)Fp * 16; strides[2] = row_bytes; strides[3] = (cuuint64_t)T * row_bytes;
            box[0] = 2; box[1] = 2 * P; box[2] = 1; box[3] = rows; box[4] = 1;
            es[1] = 2;
        }
        const CUresult r = tc3_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, const_cast<uint8_t*>(base), dims, strides, box,
                                              es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            fail(NUNET_EINVAL, "cuTensorMapEncodeTiled failed (%d) for Fp=%d planes=%d T=%d B=%d P=%d rows=%d rank=%d", (int)r, Fp, planes, T, B, P,
                 rows, rank);
        if (tc3_maps.size() > 4096) tc3_maps.clear();
        tc3_maps[key] = m;
        return m;
    }

    template <int N, int PC, bool LN, bool PAIR = false>
    void launch_tc3_t(const Tc3Params& p, int grid, size_t smem, cudaStream_t st) {
        static unsigned long long attr_set = 0;     // bit d: attribute set on device d (function attributes are per device)
        auto kfn = conv_tc3_kernel<N, PC, LN, PAIR>;
        if (!((attr_set >> cfg.device) & 1ull)) {
            CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set |= 1ull << cfg.device;
        }
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3((unsigned)grid);
        lc.blockDim = dim3(T3_THREADS);
        lc.dynamicSmemBytes = smem;
        lc.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = tc3_pdl ? 1 : 0;
        lc.attrs = at;
        lc.numAttrs = 1;
        if (p.cluster || p.pair) {
            at[1].id = cudaLaunchAttributeClusterDimension;
            at[1].val.clusterDim.x = 2;
            at[1].val.clusterDim.y = 1;
            at[1].val.clusterDim.z = 1;
            lc.numAttrs = 2;
        }
        if (p.cluster || p.pair) {
            // a persistent grid must be co-resident: clamp to the number of 2-CTA clusters the device can hold at once
            static int max_clusters = -1;
            if (max_clusters < 0) {
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, kfn, &lc) != cudaSuccess || n <= 0) n = num_sms / 2 - 2;
                max_clusters = n;
            }
            if ((int)lc.gridDim.x > 2 * max_clusters) lc.gridDim = dim3((unsigned)(2 * max_clusters));
        }
        {
            const cudaError_t le = cudaLaunchKernelEx(&lc, kfn, p);
            if (le != cudaSuccess)
                fail(NUNET_ECUDA, "conv_tc3 launch (%s): %s [grid %u, smem %zu, cluster %d, pair %d, tiles %d, mt %d, ring %d]", cur_op.c_str(),
                     cudaGetErrorString(le), lc.gridDim.x, smem, p.cluster, p.pair, p.ntiles, p.mt, p.nabuf);
        }
        if (tc3_timing_buf) {   // experiments: per-role cycles of CTA 0, per tile (serialises the stream)
            unsigned long long h[12];
            CUDA_OK(cudaStreamSynchronize(st));
            CUDA_OK(cudaMemcpy(h, tc3_timing_buf, sizeof h, cudaMemcpyDeviceToHost));
            const double tl = (double)std::max<unsigned long long>(h[3], 1);
            fprintf(stderr, "TC3TIMING %-22s tiles %5llu phases %d tma %d | mma total %6.0f wait_data %6.0f wait_acc %5.0f | epilogue total %6.0f wait %6.0f tmem_ld %5.0f | "
                            "loader total %6.0f wait_buf %6.0f table %5.0f zero+fence %5.0f bulk %5.0f  (cycles per tile)\n",
                    cur_op.c_str(), h[3], p.nphase, p.tma, h[0] / tl, h[1] / tl, h[2] / tl, h[6] / tl, h[4] / tl, h[5] / tl, h[9] / tl, h[7] / tl,
                    h[8] / tl, h[10] / tl, h[11] / tl);
        }
        const double frames = (double)p.B * p.T;
        check_launch("conv_tc3", frames * 4.0 * ((double)p.F_in * (p.C0 + p.C1) + (double)p.F_conv * N * p.nhalf));
    }

    // Split-half tensor-core path (offline plans with sh16 tensors): every conv unit of the topology is eligible.
    void launch_conv_tc3(const ConvLayer& L, const float* a_cur, const float* b_cur, const float* a_prev, const float* b_prev,
                         float* out, int B, int T, int F_in, bool src_eo, bool out_eo, cudaStream_t st, bool allow_box = true,
                         bool* probe_two = nullptr, float* out2 = nullptr) {
        Tc3Params p{};
        p.out2 = reinterpret_cast<uint8_t*>(out2);
        p.prev0 = (L.KT == 2) ? reinterpret_cast<const uint8_t*>(a_prev) : nullptr;
        p.prev1 = (L.KT == 2) ? reinterpret_cast<const uint8_t*>(b_prev) : nullptr;
        p.src_eo = src_eo ? 1 : 0;
        p.out_eo = out_eo ? 1 : 0;
        p.src0 = reinterpret_cast<const uint8_t*>(a_cur);
        p.src1 = reinterpret_cast<const uint8_t*>(b_cur);
        p.C0 = L.CA; p.C1 = L.CB;
        p.wpk = reinterpret_cast<const uint8_t*>(pool.at(L.w3));
        p.bias = pool.at(L.b3);
        p.gamma = pool.at(L.gamma); p.beta = pool.at(L.beta); p.alpha = pool.at(L.alpha);
        p.out = reinterpret_cast<uint8_t*>(out);
        p.wscale_inv = L.wscale_inv;
        p.fence_mode = tc3_fence_mode;
        p.dbg = tc3_dbg;
        p.timing = tc3_timing_buf;
        p.B = B; p.T = T; p.F_in = F_in;
        p.F_conv = (L.stride == 2) ? F_in / 2 : F_in;
        p.F_out = p.F_conv * L.COUT / L.PC3;
        if (L.CB != 0 && L.CB != L.CA) fail(NUNET_EINVAL, "conv_tc3: the two sources must have the same width");
        p.ntaps = L.KT * L.KF;
        p.padrow = (L.KT == 2) ? 1 : 0;
        p.nimg = 1;
        p.img_mul[0] = 1; p.img_add[0] = 0; p.img_mul[1] = 1; p.img_add[1] = 0;
        // Tensor-map boxes (tma == 2) need contiguous source positions per image: stride-1 units over bin-ordered sources,
        // stride-2 units over [even | odd] sources.  A box row is P positions: up to 128 it is one run of 8-byte words with
        // the exact pitch; wider rows are made of 128-byte lines, so their pitch is rounded up to 8 positions.
        const int Fp = (L.stride == 2) ? F_in / 2 : F_in;     // storage positions one image row can read
        const int P_min = (L.KT == 1 && L.KF == 1) ? F_in : (L.stride == 1 && L.KF == 3) ? F_in + 2 : Fp + 1;
        // Stride-2 units over bin-ordered sources (the inner convs of a sub-U-Net, whose inputs also feed stride-1 skips) use
        // a traversal stride of two positions instead: the box dimension is 2 P positions, so P <= 128.
        const bool box_strided = L.stride == 2 && !src_eo;
        p.prev_rows = p.prev0 ? 1 : 0;
        // (a streaming step, T = 1 with history, has the carried row in every tile: no boxes there)
        const bool box = allow_box && tc3_tma == 2 && tc3_encode_tiled() && (!p.prev0 || T > 1) && (L.stride == 2 || !src_eo) && p.F_conv >= tc3_box_minf &&
                         (P_min <= 128 || (Fp % 8 == 0 && !box_strided)) && (!box_strided || tc3_box_strided);
        const bool box_lines = box && P_min > 128;
        const int P_pad = box_lines ? (P_min + 7) / 8 * 8 - P_min : 0;    // extra zero positions per flat row
        if (L.stride == 1 && L.KF == 3 && L.padl == 1 && L.KT == 2) {          // spconv
            p.P = F_in + 2 + P_pad; p.img_add[0] = -1; p.lead = p.P + 1; p.xlo = 1;
            for (int kt = 0; kt < 2; ++kt)
                for (int kf = 0; kf < 3; ++kf) { p.tap_img[kt * 3 + kf] = 0; p.tap_off[kt * 3 + kf] = kt * p.P + kf; }
        } else if (L.stride == 2 && L.KF == 3 && L.padl == 1 && L.KT == 2) {   // conv
            p.P = p.F_conv + 1 + P_pad; p.nimg = 2; p.lead = p.P; p.xlo = 0;
            p.img_mul[0] = 2; p.img_add[0] = 0; p.img_mul[1] = 2; p.img_add[1] = -1;
            for (int kt = 0; kt < 2; ++kt) {
                p.tap_img[kt * 3 + 0] = 1; p.tap_off[kt * 3 + 0] = kt * p.P;
                p.tap_img[kt * 3 + 1] = 0; p.tap_off[kt * 3 + 1] = kt * p.P;
                p.tap_img[kt * 3 + 2] = 1; p.tap_off[kt * 3 + 2] = kt * p.P + 1;
            }
        } else if (L.KT == 1 && L.KF == 1) {                                   // inconv 1x1
            p.P = F_in; p.lead = 0; p.xlo = 0; p.tap_img[0] = 0; p.tap_off[0] = 0;
        } else if (L.KT == 1 && L.KF == 3 && L.stride == 2 && L.padl == 0) {   // down_sampling
            p.P = p.F_conv + 1 + P_pad; p.nimg = 2; p.lead = 0; p.xlo = 0;
            p.img_mul[0] = 2; p.img_add[0] = 0; p.img_mul[1] = 2; p.img_add[1] = 1;
            p.tap_img[0] = 0; p.tap_off[0] = 0; p.tap_img[1] = 1; p.tap_off[1] = 0; p.tap_img[2] = 0; p.tap_off[2] = 1;
        } else if (L.KT == 1 && L.KF == 2 && L.stride == 1 && L.padl == 1) {   // up_sampling o inconv
            p.P = F_in + 1 + P_pad; p.img_add[0] = -1; p.lead = 1; p.xlo = 1;
            p.tap_img[0] = 0; p.tap_off[0] = 0; p.tap_img[1] = 0; p.tap_off[1] = 1;
        } else {
            fail(NUNET_EINVAL, "conv_tc3: unsupported unit geometry");
        }
        int maxoff = 0;
        for (int i = 0; i < p.ntaps; ++i) maxoff = std::max(maxoff, p.tap_off[i]);
        // bulk-copy (TMA) row segments need contiguous source bins: stride-1 units over bin-ordered sources, stride-2 units
        // over [even | odd] sources; <= 32 (16 per image) frame rows per tile
        p.tma = (tc3_tma && !p.prev0 && ((p.nimg == 1 && !src_eo && F_in >= tc3_tma_minf) || (p.nimg == 2 && src_eo && p.F_conv >= 32))) ? 1 : 0;
        int box_dmax = 0;
        if (box) {
            p.tma = 2;
            p.tm_rank = (box_lines || box_strided) ? 5 : 4;
            for (int i = 0; i < p.nimg; ++i) {
                const int par = (L.stride == 2 && src_eo) ? (p.img_add[i] & 1) : 0;
                p.tm_par[i] = par;
                p.tm_delta[i] = (L.stride == 2 && src_eo) ? (p.img_add[i] - par) / 2 : p.img_add[i];     // a_i: storage position read by x = 0
            }
            for (int i = 0; i < p.nimg; ++i) {
                const int c = box_lines ? (p.tm_delta[i] < 0 ? -8 : 0) : p.tm_delta[i];     // first storage position of a box row
                p.tm_delta[i] -= c;
                box_dmax = std::max(box_dmax, p.tm_delta[i]);
                if (c + (box_strided ? 2 : 1) * p.P < (box_strided ? F_in : Fp)) fail(NUNET_EINVAL, "conv_tc3: box row does not cover the source row");
                p.tm_c[i] = box_lines ? c / 8 : box_strided ? c : 2 * c;
            }
        }
        p.cluster = (p.tma == 1 && L.nhalf3 == 2 && tc3_cluster) ? 1 : 0;
        p.nphase = (L.CA + L.CB) / T3_KCH;
        p.nhalf = L.nhalf3;
        p.w_half_bytes = p.nphase * p.ntaps * L.N3 * 64;
        const long long total = (long long)B * (T + p.padrow) * p.P;
        if (total >= 0x7fffffffLL - 1024) fail(NUNET_EINVAL, "conv_tc3: more than 2^31 flat positions");
        p.total_flat = (int)total;
        const size_t fixed = (size_t)T3_FIXED_BYTES + (size_t)p.w_half_bytes;
        const size_t limit = 227 * 1024;
        // tile = mt x 128 positions; prefer two tiles per iteration unless that leaves a ring of fewer than three
        // image buffers while one tile would allow it (the ring depth is what hides the HBM latency)
        // a box image starts on a frame-row boundary, (tile * mt * 128 - lead) mod P positions before the tile's first one
        auto box_xoff_max = [&](int mt) {
            int g = mt * 128, b = p.P;
            while (b) { const int r = g % b; g = b; b = r; }
            return p.P - g + (g - p.lead % g) % g;
        };
        // Row tiles (units whose F_conv is a multiple of 128): tile u covers bins [128 j, 128 j + 128) of one frame row, so no pad
        // position is computed and every image starts 128 j positions into its first frame row -- the fewest box rows.
        const int row_tpr = (box && tc3_row_tiles && p.F_conv % 128 == 0) ? p.F_conv / 128 : 0;
        const int tile2_off = (row_tpr == 1) ? p.P : 128;
        auto box_rows = [&](int mt, int slots) {
            if (!row_tpr) return (box_xoff_max(mt) + box_dmax + slots - 1) / p.P + 1;
            int rows = 0;
            for (int k = 0; k < row_tpr; ++k) {      // an iteration starts at tile k * mt of some frame row
                const int j = (k * mt) % row_tpr;
                rows = std::max(rows, (128 * j + box_dmax + slots - 1) / p.P + 1);
            }
            return rows;
        };
        auto geometry = [&](int mt, int& slots, int& plane_bytes, size_t& abuf) {
            slots = 128 + (mt - 1) * tile2_off + maxoff;
            int plane16 = (p.nimg * slots + 31) / 32 * 32;   // both images + the loaders' round-up padding
            if (box) {   // whole frame rows per image, every image and plane on a 128-byte boundary (tensor-copy destination)
                const int rows = box_rows(mt, slots);
                const int img16 = (rows * p.P + 7) / 8 * 8;
                plane16 = std::max(plane16, p.nimg * img16);
                if (rows > 256) return 0;
            } else {
                while (plane16 % 8 != 2) ++plane16;
            }
            plane_bytes = plane16 * 16;
            abuf = (size_t)4 * plane_bytes;
            if (p.nimg * slots > T3_TBL - 32 || fixed + 2 * abuf > limit) return 0;
            return (int)std::min<size_t>((limit - fixed) / abuf, (size_t)T3_MAXNB);
        };
        int slots2, plane2, slots1, plane1;
        size_t abuf2, abuf1;
        const int nb2 = geometry(2, slots2, plane2, abuf2), nb1 = geometry(1, slots1, plane1, abuf1);
        const bool ok = nb2 >= 2 || nb1 >= 2;
        bool two = nb2 >= 2;   // larger tiles beat a deeper ring (measured): per-tile overheads dominate
        if (tc3_force_mt == 1 && nb1 >= 2) two = false;
        if (tc3_force_mt == 2 && nb2 >= 2) two = true;
        // 128-channel units over box images: CTA pairs with cta_group::2 MMAs (one 128-position tile per CTA)
        const bool pair = box && L.nhalf3 == 2 && L.N3 == 64 && L.PC3 == 64 && L.epi != EPI_BIAS && tc3_pair && nb1 >= 2 &&
                          p.F_conv >= tc3_pair_minf;
        if (pair) two = false;
        if (probe_two) {
            *probe_two = two;
            return;
        }
        if (box && !two && !pair) {
            // whole-row boxes are larger than the exact images of the row-segment loader: when only the latter fits two tiles
            // per iteration, it wins (measured: 128-channel stride-2 unit at 128 bins, 2.3 ms against 3.3 ms)
            bool two_rows = false;
            launch_conv_tc3(L, a_cur, b_cur, a_prev, b_prev, out, B, T, F_in, src_eo, out_eo, st, false, &two_rows);
            if (two_rows) {
                launch_conv_tc3(L, a_cur, b_cur, a_prev, b_prev, out, B, T, F_in, src_eo, out_eo, st, false, nullptr, out2);
                return;
            }
        }
        p.mt = two ? 2 : 1;
        p.slots = two ? slots2 : slots1;
        p.plane_bytes = two ? plane2 : plane1;
        p.nabuf = two ? nb2 : nb1;
        const size_t smem = fixed + (size_t)p.nabuf * (two ? abuf2 : abuf1);
        if (!ok) fail(NUNET_EINVAL, "conv_tc3: unit does not fit shared memory");
        if (box) {
            p.tm_rows = box_rows(p.mt, p.slots);
            p.tm_box_bytes = p.tm_rows * p.P * 16;
            p.tm_img_bytes = (p.tm_rows * p.P + 7) / 8 * 8 * 16;
            const int planes = (src_eo ? 2 : 1) * (L.CA / 4);          // hi | lo  x  C/8 chunks (x parity halves)
            const int kind = box_strided ? 2 : box_lines ? 1 : 0;
            const int Fm = box_strided ? F_in : Fp;
            p.tm_map[0] = tc3_tensor_map(p.src0, Fm, planes, T, B, (size_t)F_in * L.CA * 4, p.P, p.tm_rows, kind);
            if (p.src1) p.tm_map[1] = tc3_tensor_map(p.src1, Fm, planes, T, B, (size_t)F_in * L.CA * 4, p.P, p.tm_rows, kind);
        }
        p.row_tpr = row_tpr;
        p.tile2_off = tile2_off;
        {   // reciprocals of the geometry's divisors (conv_tc3.cuh: fast_div)
            auto magic = [](int d) -> unsigned long long { return d <= 1 ? 0ull : (unsigned long long)(~0ull / (unsigned long long)d) + 1ull; };
            p.mg_P = magic(p.P);
            p.mg_Tp = magic(T + p.padrow);
            p.mg_row_tpr = magic(row_tpr);
        }
        // units with many short phases (up_sampling o inconv: 8 phases of 2 taps) are bound by the box issue of one lane per plane
        p.ld_rr = (tc3_ld_rr && box && p.ntaps <= 2 && p.nphase >= 8) ? 1 : 0;
        p.tm_dmin = 0;
        // 128-position tiles of the launch: consecutive runs of the flat axis, or row_tpr per frame row
        const long long units = row_tpr ? (long long)B * (T + p.padrow) * row_tpr : (total + 127) / 128;
        p.ntiles = (int)((units + p.mt - 1) / p.mt);
        if (pair) {
            int g = 128, b = p.P;
            while (b) { const int r = g % b; g = b; b = r; }
            p.pair = 1;
            p.pair_m = row_tpr ? row_tpr : p.P / g;                      // tiles between the two CTAs of a pair: a whole number of frame rows
            p.mg_pair_m = p.pair_m <= 1 ? 0ull : (~0ull / (unsigned long long)p.pair_m) + 1ull;
            p.ntiles = (int)((units + 2 * p.pair_m - 1) / (2 * p.pair_m)) * p.pair_m;   // cluster work units
        }
        const int grid = std::max(1, std::min(p.ntiles, num_sms / p.nhalf)) * p.nhalf;
        const bool ln = (L.epi != EPI_BIAS);
        if (L.N3 == 32 && L.PC3 == 32 && ln) launch_tc3_t<32, 32, true>(p, grid, smem, st);
        else if (L.N3 == 64 && L.PC3 == 64 && ln && p.pair) launch_tc3_t<64, 64, true, true>(p, grid, smem, st);
        else if (L.N3 == 64 && L.PC3 == 64 && ln) launch_tc3_t<64, 64, true>(p, grid, smem, st);
        else if (L.N3 == 64 && L.PC3 == 64 && !ln) launch_tc3_t<64, 64, false>(p, grid, smem, st);
        else if (L.N3 == 64 && L.PC3 == 32 && ln) launch_tc3_t<64, 32, true>(p, grid, smem, st);
        else fail(NUNET_EINVAL, "conv_tc3: no kernel for N=%d PC=%d", L.N3, L.PC3);
    }

    void launch_conv(const ConvLayer& L, const float* a_cur, const float* a_prev, const float* b_cur,
                     const float* b_prev, float* out, int B, int T, bool has_prev, int F_in, cudaStream_t st) {
        ConvParams p;
        p.a_cur = a_cur; p.a_prev = a_prev; p.b_cur = b_cur; p.b_prev = b_prev;
        p.w = pool.at(L.w); p.bias = pool.at(L.bias);
        p.gamma = pool.at(L.gamma); p.beta = pool.at(L.beta); p.alpha = pool.at(L.alpha);
        p.out = out;
        p.CA = L.CA; p.CB = L.CB; p.B = B; p.T = T; p.has_prev = has_prev ? 1 : 0;
        p.F_in = F_in;
        p.F_out = (L.stride == 2) ? F_in / 2 : F_in;
        p.KT = L.KT; p.KF = L.KF; p.padl = L.padl; p.stride = L.stride;
        // tile geometry: 128 pixels = G clips x TT frames x FT bins
        int lFT = 0;
        while ((1 << (lFT + 1)) <= p.F_out && (1 << (lFT + 1)) <= 32) ++lFT;
        const int FT = 1 << lFT;
        int rem = CONV_P / FT;
        int tcap = (FT == 32) ? 4 : 8;
        int lTT = 0;
        while ((1 << (lTT + 1)) <= rem && (1 << (lTT + 1)) <= tcap && (1 << lTT) < T) ++lTT;
        int lG = 0;
        while ((1 << (lG + 1)) <= rem >> lTT) ++lG;
        const int Cmax = L.CA > L.CB ? L.CA : L.CB;
        const size_t wbytes = 2 * CONV_KC * L.COUT * sizeof(float);
        const size_t budget = 200 * 1024;
        while (lG > 0 && wbytes + sizeof(float) * conv_tile_floats(1 << lG, 1 << lTT, FT, L.KT, L.KF, L.stride, Cmax) > budget) --lG;
        // no point in more clip slots than clips
        while (lG > 0 && (1 << (lG - 1)) >= B) --lG;
        p.lFT = lFT; p.lTT = lTT; p.lG = lG;
        const size_t smem = wbytes + sizeof(float) * conv_tile_floats(1 << lG, 1 << lTT, FT, L.KT, L.KF, L.stride, Cmax);
        if (smem > 220 * 1024) fail(NUNET_EINVAL, "conv tile does not fit shared memory (%zu bytes)", smem);
        const int nfb = p.F_out >> lFT, ntb = (T + (1 << lTT) - 1) >> lTT, nbg = (B + (1 << lG) - 1) >> lG;
        const long long grid = (long long)nfb * ntb * nbg;
        if (grid <= 0 || grid > 0x7fffffffLL) fail(NUNET_EINVAL, "conv grid out of range");
        if (L.COUT == 32 && L.epi == EPI_LN) launch_conv_t<32, 4, 8, 128, EPI_LN>(p, (int)grid, smem, st);
        else if (L.COUT == 64 && L.epi == EPI_LN) launch_conv_t<64, 8, 4, 256, EPI_LN>(p, (int)grid, smem, st);
        else if (L.COUT == 64 && L.epi == EPI_BIAS) launch_conv_t<64, 8, 4, 256, EPI_BIAS>(p, (int)grid, smem, st);
        else if (L.COUT == 64 && L.epi == EPI_SHUF32) launch_conv_t<64, 8, 4, 256, EPI_SHUF32>(p, (int)grid, smem, st);
        else if (L.COUT == 128 && L.epi == EPI_SHUF64) launch_conv_t<128, 8, 8, 256, EPI_SHUF64>(p, (int)grid, smem, st);
        else fail(NUNET_EINVAL, "no conv kernel for COUT=%d epi=%d", L.COUT, L.epi);
    }

    MlpW mlpw(const MlpLayer& m) const { return MlpW{pool.at(m.k0), pool.at(m.b0), pool.at(m.k1), pool.at(m.b1)}; }

    // -------------------------------------------------------------------------------- topology -> plan
    // conv unit reading one or two tensors
    Ten* op_conv(Plan& P, const std::string& role, Ten* a, Ten* b, const std::string& out_name, bool persistent,
                 bool out_eo = false, bool want_twin = false) {
        const ConvLayer& L = convs.at(role);
        // a stride-2 unit reads the [even | odd] copies of its sources when their producers wrote one: its tile images are
        // then whole plane rows per tensor-map box instead of a box of 16-byte pieces (traversal stride two)
        if (L.stride == 2 && a->twin && !a->eo && (!b || (b->twin && !b->eo))) {
            a = a->twin;
            if (b) b = b->twin;
        }
        if (a->C != L.CA || (b ? b->C : 0) != L.CB || (b && b->F != a->F)) fail(NUNET_EINVAL, "plan: %s wiring", role.c_str());
        const int F_in = a->F;
        const int F_conv = (L.stride == 2) ? F_in / 2 : F_in;
        const bool shuf = (L.epi == EPI_SHUF32 || L.epi == EPI_SHUF64);
        Ten* o = P.make(out_name, shuf ? 2 * F_conv : F_conv, shuf ? L.COUT / 2 : L.COUT, persistent);
        o->sh = P.sh16;
        o->eo = P.sh16 && out_eo;
        Ten* o2 = nullptr;
        if (want_twin && tc3_twin && !P.streaming && P.sh16 && !o->eo && !fz_open && (o->F % 2) == 0) {
            o2 = P.make("", o->F, o->C, persistent);
            o2->sh = true;
            o2->eo = true;
            o->twin = o2;
        }
        if (b && b->eo != a->eo) fail(NUNET_EINVAL, "plan: %s sources disagree on the bin order", role.c_str());
        const bool src_eo = a->eo, dst_eo = o->eo;
        Plan* pp = &P;
        size_t a_coff = 0, b_coff = 0;
        if (!P.streaming && P.sh16 && L.KT == 2) {   // causal conv: its inputs' last rows are carried between time chunks
            a_coff = P.want_carry(a);
            if (b) b_coff = P.want_carry(b);
        }
        if (fz_open) {
            if (!L.wfz) fail(NUNET_EINVAL, "plan: %s has no fused-kernel weights", role.c_str());
            fz_open->alg_bytes += 4.0 * ((double)F_in * (L.CA + L.CB) * L.KT + (double)o->numel());
            fz_open->fill.push_back([=](Engine& E, const Run& r, FzLayer& f) {
                f.kind = FZ_CONV;
                f.a_cur = reinterpret_cast<const uint8_t*>(pp->cur(a, r.parity));
                f.a_prev = reinterpret_cast<const uint8_t*>(pp->prev(a, r.parity));
                f.b_cur = b ? reinterpret_cast<const uint8_t*>(pp->cur(b, r.parity)) : nullptr;
                f.b_prev = b ? reinterpret_cast<const uint8_t*>(pp->prev(b, r.parity)) : nullptr;
                f.out = reinterpret_cast<uint8_t*>(pp->cur(o, r.parity));
                f.w = reinterpret_cast<const uint4*>(E.pool.at(L.wfz));
                f.bias = E.pool.at(L.b3); f.gamma = E.pool.at(L.gamma); f.beta = E.pool.at(L.beta); f.alpha = E.pool.at(L.alpha);
                f.wscale_inv = L.wscale_inv;
                f.F_in = F_in; f.Ca = L.CA; f.Cb = L.CB; f.F_conv = F_conv; f.N = L.COUT; f.KT = L.KT; f.KF = L.KF; f.padl = L.padl;
                f.stride = L.stride;
                f.epi = L.epi == EPI_LN ? FZE_LN : L.epi == EPI_SHUF32 ? FZE_SHUF32 : L.epi == EPI_SHUF64 ? FZE_SHUF64 : FZE_BIAS;
                f.in_eo = src_eo ? 1 : 0; f.out_eo = dst_eo ? 1 : 0;
            });
            return o;
        }
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = out_name;
            if (pp->sh16) {
                const float* ap = pp->prev(a, r.parity);
                const float* bp = b ? pp->prev(b, r.parity) : nullptr;
                if (r.use_carry && L.KT == 2) {
                    ap = pp->carry_at(a_coff);
                    bp = b ? pp->carry_at(b_coff) : nullptr;
                }
                E.launch_conv_tc3(L, pp->cur(a, r.parity), b ? pp->cur(b, r.parity) : nullptr, ap, bp, pp->cur(o, r.parity), r.B, r.T, F_in,
                                  src_eo, dst_eo, r.st, true, nullptr, o2 ? pp->cur(o2, r.parity) : nullptr);
                return;
            }
            E.launch_conv(L, pp->cur(a, r.parity), pp->prev(a, r.parity), b ? pp->cur(b, r.parity) : nullptr,
                          b ? pp->prev(b, r.parity) : nullptr, pp->cur(o, r.parity), r.B, r.T, pp->streaming, F_in, r.st);
        });
        return o;
    }

    // Reshape [T, F*C] -> LSTM(21) -> Dense(F*C) -> Reshape (models/proposed.py:305-309)
    Ten* op_lstm(Plan& P, const std::string& lstm, Ten* x, const std::string& out_name, const std::string& state_name,
                 bool persistent) {
        const LstmLayer& L = lstms.at(lstm);
        const int D = x->F * x->C;
        if (D != L.D) fail(NUNET_EINVAL, "plan: %s width", lstm.c_str());
        Ten *hst = nullptr, *cst = nullptr;
        if (P.streaming) {
            hst = P.make("", 1, LSTM_UNITS, true, false);
            cst = P.make("", 1, LSTM_UNITS, true, false);
            Plan::StateRef sh, sc;
            sh.name = state_name + "_h"; sh.is_lstm = true; sh.a = hst;
            sc.name = state_name + "_c"; sc.is_lstm = true; sc.a = cst;
            P.states.push_back(sh);
            P.states.push_back(sc);
        }
        Ten* o = P.make(out_name, x->F, x->C, persistent);
        o->sh = P.sh16;
        Plan* pp = &P;
        const size_t hc_off = P.streaming ? 0 : P.carry_alloc(2 * 32);     // offline: h | c carried between time chunks, 32 floats each
        if (fz_open) {
            const int xC = x->C;
            fz_open->alg_bytes += 8.0 * D;
            fz_open->fill.push_back([=](Engine& E, const Run& r, FzLayer& f) {
                f.kind = FZ_LSTM;
                f.a_cur = reinterpret_cast<const uint8_t*>(pp->cur(x, r.parity));
                f.out = reinterpret_cast<uint8_t*>(pp->cur(o, r.parity));
                f.wk = E.pool.at(L.wk); f.wr = E.pool.at(L.wr); f.bk = E.pool.at(L.wb); f.wd = E.pool.at(L.dk); f.bd = E.pool.at(L.db);
                f.h = pp->cur(hst, 0); f.c = pp->cur(cst, 0);
                f.D = D; f.C = xC;
            });
            return o;
        }
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = out_name;
            const long long rows = (long long)r.B * r.T;
            const int xC = x->C;
            const size_t smem = (size_t)LSTM_TB * 2 * (D + LSTM_GATES + 24) * sizeof(float);
            float* hp = hst ? pp->cur(hst, 0) : nullptr;
            float* cp = cst ? pp->cur(cst, 0) : nullptr;
            int zero_init = 0;
            if (!pp->streaming && (r.use_carry || r.save_carry)) {
                hp = pp->carry_at(hc_off);
                cp = hp + (size_t)pp->carry_cap * LSTM_UNITS;
                zero_init = r.use_carry ? 0 : 1;
            }
            if (pp->streaming && r.T == 1 && hp && E.lstm_stream) {
                // a streaming step: groups of 16 streams per CTA share the weight reads
                const size_t sm2 = (size_t)LSTM_TB * (D + LSTM_GATES + 24) * sizeof(float);
                const int grid = (r.B + LSTM_TB - 1) / LSTM_TB;
                if (pp->sh16)
                    lstm_stream_kernel<true><<<grid, LSTM_THREADS, sm2, r.st>>>(pp->cur(x, r.parity), E.pool.at(L.wk4), E.pool.at(L.wr), E.pool.at(L.wb),
                                                                           E.pool.at(L.dk), E.pool.at(L.db), hp, cp, pp->cur(o, r.parity), r.B, D, xC);
                else
                    lstm_stream_kernel<false><<<grid, LSTM_THREADS, sm2, r.st>>>(pp->cur(x, r.parity), E.pool.at(L.wk4), E.pool.at(L.wr), E.pool.at(L.wb),
                                                                            E.pool.at(L.dk), E.pool.at(L.db), hp, cp, pp->cur(o, r.parity), r.B, D, xC);
                E.check_launch("lstm", rows * 4.0 * (2.0 * D + 2 * LSTM_UNITS));
                return;
            }
            // one CTA per clip / stream: projection, recurrence and Dense in one kernel (lstm_kernels.cuh)
            if (pp->sh16)
                lstm_block_kernel<true><<<r.B, LSTM_THREADS, smem, r.st>>>(pp->cur(x, r.parity), E.pool.at(L.wk4), E.pool.at(L.wr), E.pool.at(L.wb),
                                                                  E.pool.at(L.dk), E.pool.at(L.db), hp, cp, zero_init, pp->cur(o, r.parity), r.T, D, xC);
            else
                lstm_block_kernel<false><<<r.B, LSTM_THREADS, smem, r.st>>>(pp->cur(x, r.parity), E.pool.at(L.wk4), E.pool.at(L.wr), E.pool.at(L.wb),
                                                                   E.pool.at(L.dk), E.pool.at(L.db), hp, cp, zero_init, pp->cur(o, r.parity), r.T, D, xC);
            E.check_launch("lstm", rows * 4.0 * (2.0 * D + 2 * LSTM_UNITS));
        });
        return o;
    }

    // Dilated dense block bottleneck (models/nunet_tls.py:383-410; one-frame form converter_nunet_tls.py:373-411).
    // Offline: seven dense fp32 intermediates [frame][F][h].  Streaming: out_0..out_5 are per-stream rings of DDB_RING
    // steps (layer k looks d = 2^(k-1) steps back), out_6 and the block input are ping-ponged like every activation.
    Ten* op_ddb(Plan& P, const std::string& role, Ten* x, const std::string& out_name, bool persistent) {
        const DdbLayer L = ddbs.at(role);
        const int C = x->C, h = C / 2, F = x->F;
        if (C != L.C) fail(NUNET_EINVAL, "plan: %s width", role.c_str());
        Ten* mid[7];
        for (int i = 0; i < 6; ++i) mid[i] = P.streaming ? P.make("", F * DDB_RING, h, true, false) : P.make("", F, h, false, false);
        mid[6] = P.make("", F, h, P.streaming, true);
        Ten* o = P.make(out_name, F, C, persistent);
        o->sh = P.sh16;
        if (P.streaming) {
            Plan::StateRef si, so;
            si.name = role + "_in"; si.a = x;
            P.states.push_back(si);
            for (int k = 1; k <= 6; ++k) {
                Plan::StateRef s;
                s.name = role + "_" + std::to_string(k);
                s.ddb_k = k; s.ddb_d = 1 << (k - 1);
                for (int j = 0; j < k; ++j) s.rings[j] = mid[j];
                s.a = mid[0];
                P.states.push_back(s);
            }
            so.name = role + "_out"; so.a = mid[6];
            P.states.push_back(so);
        }
        Plan* pp = &P;
        const bool sh = P.sh16;
        // offline: history of a time chunk -- 32 frames of out_0..out_5, one row of the block input (fp32) and of out_6, per clip
        size_t hist_off[8] = {};
        if (!P.streaming) {
            for (int i = 0; i < 6; ++i) hist_off[i] = P.carry_alloc((size_t)DDB_HIST * F * h);
            hist_off[6] = P.carry_alloc((size_t)F * C);
            hist_off[7] = P.carry_alloc((size_t)F * h);
        }
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = out_name;
            const long long units = (long long)r.B * r.T;
            const bool chunked = !pp->streaming && (r.use_carry || r.save_carry);
            if (chunked && !pp->carry) fail(NUNET_EINVAL, "time chunking needs the carry arena");
            if (!pp->streaming && !chunked) {
                // Whole clips offline: one launch per layer (in, six dilated layers, out).  Measured at 256 clips x 249 frames the
                // one-CTA-per-clip kernel below takes 12.4 ms per step against 8.8 ms for these eight grid-wide launches per block:
                // a clip's 996 pixels are too few threads to hide the L2 latency of the dense block's gather.  The fused kernel
                // serves streaming (13 instead of 104 launches per step) and time-chunked calls (it reads the carried history);
                // both paths evaluate every output with the same operation sequence (bit-identical, tests/test_gpu_chunking.py).
                const long long nin = units * F * h;
                DdbGeom g{pp->streaming ? 1 : 0, r.step & (DDB_RING - 1), r.T};
                float* m[7];
                for (int i = 0; i < 6; ++i) m[i] = pp->cur(mid[i], 0);
                m[6] = pp->cur(mid[6], r.parity);
                const void* xin = pp->cur(x, r.parity);
                const void* xprev = pp->prev(x, r.parity);
                {
                    const int blocks = (int)((units * F + 31) / 32);
                    const size_t smem = (size_t)(6 * C * h + 6 * C * 32) * sizeof(float);
                    auto go = [&](auto kfn) {
                        CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
                        kfn<<<blocks, 128, smem, r.st>>>(xin, xprev, E.pool.at(L.w_in), E.pool.at(L.b_in), E.pool.at(L.a_in), m[0], units, g, F);
                    };
                    if (C == 64) { if (sh) go(ddb_conv23_kernel<64, 32, true, false, true>); else go(ddb_conv23_kernel<64, 32, true, false, false>); }
                    else { if (sh) go(ddb_conv23_kernel<32, 16, true, false, true>); else go(ddb_conv23_kernel<32, 16, true, false, false>); }
                }
                E.check_launch("ddb_in", units * 4.0 * F * (C + h));
                DdbOuts src{};
                for (int i = 0; i < 6; ++i) src.o[i] = m[i];
                for (int k = 1; k <= 6; ++k) {
                    const int d = 1 << (k - 1);
                    const int blocks = (int)((nin + 127) / 128);
                    const int ring_out = (pp->streaming && k < 6) ? 1 : 0;
                    if (h == 16)
                        ddb_layer_kernel<16><<<blocks, 128, 0, r.st>>>(src, k, d, E.pool.at(L.w0[k - 1]), E.pool.at(L.b0[k - 1]), E.pool.at(L.w1[k - 1]),
                                                                      E.pool.at(L.b1[k - 1]), E.pool.at(L.gamma[k - 1]), E.pool.at(L.beta[k - 1]),
                                                                      E.pool.at(L.alpha[k - 1]), m[k], ring_out, units, g, F);
                    else
                        ddb_layer_kernel<32><<<blocks, 128, 0, r.st>>>(src, k, d, E.pool.at(L.w0[k - 1]), E.pool.at(L.b0[k - 1]), E.pool.at(L.w1[k - 1]),
                                                                      E.pool.at(L.b1[k - 1]), E.pool.at(L.gamma[k - 1]), E.pool.at(L.beta[k - 1]),
                                                                      E.pool.at(L.alpha[k - 1]), m[k], ring_out, units, g, F);
                    E.check_launch("ddb_layer", units * 4.0 * F * h * (k + 1));
                }
                void* yo = pp->cur(o, r.parity);
                const float* o6prev = pp->prev(mid[6], r.parity);
                {
                    const int blocks = (int)((units * F + 31) / 32);
                    const size_t smem = (size_t)(6 * h * C + 6 * h * 32) * sizeof(float);
                    auto go = [&](auto kfn) {
                        CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
                        kfn<<<blocks, 128, smem, r.st>>>(m[6], o6prev, E.pool.at(L.w_out), E.pool.at(L.b_out), E.pool.at(L.a_out), yo, units, g, F);
                    };
                    if (C == 64) { if (sh) go(ddb_conv23_kernel<32, 64, false, true, true>); else go(ddb_conv23_kernel<32, 64, false, true, false>); }
                    else { if (sh) go(ddb_conv23_kernel<16, 32, false, true, true>); else go(ddb_conv23_kernel<16, 32, false, true, false>); }
                }
                E.check_launch("ddb_out", units * 4.0 * F * (C + h));
                return;
            }
            DdbFused f{};
            f.g = DdbGeom{pp->streaming ? 1 : 0, r.step & (DDB_RING - 1), r.T};
            f.x = pp->cur(x, r.parity);
            f.x_prev = pp->prev(x, r.parity);
            f.y = pp->cur(o, r.parity);
            for (int i = 0; i < 6; ++i) f.mid[i] = pp->cur(mid[i], 0);
            f.mid[6] = pp->cur(mid[6], r.parity);
            f.mid6_prev = pp->prev(mid[6], r.parity);
            if (!pp->streaming && r.use_carry) {
                for (int i = 0; i < 6; ++i) f.hist_mid[i] = pp->carry_at(hist_off[i]);
                f.hist_x = pp->carry_at(hist_off[6]);
                f.hist_mid6 = pp->carry_at(hist_off[7]);
            }
            f.w_in = E.pool.at(L.w_in); f.b_in = E.pool.at(L.b_in); f.a_in = E.pool.at(L.a_in);
            f.w_out = E.pool.at(L.w_out); f.b_out = E.pool.at(L.b_out); f.a_out = E.pool.at(L.a_out);
            for (int k = 0; k < 6; ++k) {
                f.w0[k] = E.pool.at(L.w0[k]); f.b0[k] = E.pool.at(L.b0[k]); f.w1[k] = E.pool.at(L.w1[k]); f.b1[k] = E.pool.at(L.b1[k]);
                f.gamma[k] = E.pool.at(L.gamma[k]); f.beta[k] = E.pool.at(L.beta[k]); f.alpha[k] = E.pool.at(L.alpha[k]);
            }
            f.F = F;
            f.units = units;
            f.units_per_cta = pp->streaming ? 8 : r.T;          // one clip per CTA offline (its layers depend only on its own frames)
            const int grid = (int)((units + f.units_per_cta - 1) / f.units_per_cta);
            const int NG = (C == 32) ? 4 : 2;
            const size_t smem = (size_t)(6 * C * h + NG * 6 * C * 32) * sizeof(float);
            auto go = [&](auto kfn) {
                CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                kfn<<<grid, DDBF_THREADS, smem, r.st>>>(f);
            };
            if (C == 64) { if (sh) go(ddb_block_kernel<64, true>); else go(ddb_block_kernel<64, false>); }
            else { if (sh) go(ddb_block_kernel<32, true>); else go(ddb_block_kernel<32, false>); }
            E.check_launch("ddb_block", units * 4.0 * F * (2.0 * C + 2.0 * h * 7 + 21.0 * h));
            if (!pp->streaming && r.save_carry) {
                DdbHistW hw{};
                for (int i = 0; i < 6; ++i) hw.mid[i] = pp->carry_at(hist_off[i]);
                hw.x = pp->carry_at(hist_off[6]);
                hw.mid6 = pp->carry_at(hist_off[7]);
                if (sh) ddb_hist_update_kernel<true><<<dim3(r.B, 7), 256, 0, r.st>>>(f, hw, C, r.use_carry ? 1 : 0);
                else ddb_hist_update_kernel<false><<<dim3(r.B, 7), 256, 0, r.st>>>(f, hw, C, r.use_carry ? 1 : 0);
                E.check_launch("ddb_hist", 0.0);
            }
        });
        return o;
    }

    // One nested sub-U-Net (MSFE): returns ctfa(de_1) + en_in; fills des_out[k-1] = de_k (k = 1..n, F0 >> (k-1) bins)
    Ten* op_msfe(Plan& P, const std::string& blk, int n, Ten* en_in, Ten* const* skips, Ten** des_out,
                 bool des_persistent, bool out_persistent, bool out_eo, bool fuse_out_conv = false) {
        std::string pc, ps;
        state_prefixes(blk, pc, ps);
        std::vector<Ten*> ens;
        Ten* cur = en_in;
        for (int k = 1; k <= n; ++k) {
            Ten* sk = skips ? skips[k - 1] : nullptr;
            if (P.streaming) {
                Plan::StateRef s;
                s.name = pc + "_" + std::to_string(k);
                s.a = cur;
                s.b = sk;
                P.states.push_back(s);
            }
            // streaming: from the first conv whose input has <= 32 bins on, the layers of this sub-U-Net are collected into one
            // fused_tail_kernel launch (closed below after the last sub-pixel conv that still writes <= 32 bins)
            if (P.streaming && P.sh16 && stream_fuse && !is_ddb() && !fz_open && cur->F <= 32) fz_open.reset(new FzGroup());
            // (its output feeds the next stride-2 conv and, in bin order, the sub-pixel conv of the way up)
            cur = op_conv(P, blk + "_conv" + std::to_string(k), cur, sk, blk + "_conv" + std::to_string(k), false, false, /*want_twin=*/k < n);
            ens.push_back(cur);
        }
        Ten* bb = is_ddb() ? op_ddb(P, blk + "_ddb", cur, blk + "_bb", false) : op_lstm(P, blk + "_lstm", cur, blk + "_bb", blk, false);
        cur = bb;
        auto fz_close = [&]() {
            if (!fz_open) return;
            std::shared_ptr<FzGroup> grp(fz_open.release());
            if ((int)grp->fill.size() > FZ_MAXL) fail(NUNET_EINVAL, "plan: %s fuses %zu layers", blk.c_str(), grp->fill.size());
            Plan* pq = &P;
            P.ops.push_back([=](Engine& E, const Run& r) {
                E.cur_op = blk + "_tail";
                FzParams fp{};
                fp.nl = (int)grp->fill.size();
                for (int i = 0; i < fp.nl; ++i) grp->fill[i](E, r, fp.L[i]);
                fp.S = r.B;
                fp.dbg = E.fz_dbg;
                fp.G = std::max(1, (r.B + FZ_CTAS_PER_SM * E.num_sms - 1) / (FZ_CTAS_PER_SM * E.num_sms));
                (void)pq;
                static unsigned long long attr_set = 0;
                if (!((attr_set >> E.cfg.device) & 1ull)) {
                    CUDA_OK(cudaFuncSetAttribute(fused_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
                    attr_set |= 1ull << E.cfg.device;
                }
                fused_tail_kernel<<<(r.B + fp.G - 1) / fp.G, FZ_THREADS, FZ_SMEM, r.st>>>(fp);
                E.check_launch("fused_tail", (double)r.B * grp->alg_bytes);
            });
        };
        std::vector<Ten*> des;
        for (int k = 1; k <= n; ++k) {
            Ten* sk = ens[n - k];
            if (P.streaming) {
                Plan::StateRef s;
                s.name = ps + "_" + std::to_string(k);
                s.a = cur;
                s.b = sk;
                P.states.push_back(s);
            }
            if (fz_open && cur->F > 16) fz_close();        // this sub-pixel conv would write more than 32 bins: it runs on its own
            // (second-level skips de_2 .. de_n of an encoder block are read by the stride-2 convs of the paired decoder block)
            cur = op_conv(P, blk + "_spconv" + std::to_string(k), cur, sk, blk + "_spconv" + std::to_string(k), des_persistent,
                          /*out_eo=*/k == n, /*want_twin=*/des_out != nullptr && des_persistent && k < n);
            des.push_back(cur);
        }
        fz_close();
        if (des_out)
            for (int k = 1; k <= n; ++k) des_out[k - 1] = des[n - k];
        // CTFA + residual
        Ten* x = cur;
        const int F0 = x->F;
        Ten* ta = P.make(blk + "_ta", 1, 64, false, false);
        Ten* gate = P.make(blk + "_gate", 1, 64, false, false);
        Ten* ring = nullptr;
        if (P.streaming && cfg.stream_ctfa_history) {
            ring = P.make("", CTFA_WINDOW, 64, true, false);
            P.rings.push_back(ring);
        }
        Ten* out = P.make(blk + "_out", F0, 64, out_persistent);
        out->sh = P.sh16;
        out->eo = P.sh16 && out_eo;
        if (P.sh16 && (x->eo != en_in->eo)) fail(NUNET_EINVAL, "plan: %s gate operands disagree on the bin order", blk.c_str());
        const int in_eo = x->eo ? 1 : 0, o_eo = out->eo ? 1 : 0;
        const MlpLayer mta = mlps.at(blk + "_ta"), mfa = mlps.at(blk + "_fa");
        Plan* pp = &P;
        const int off_mode = cfg.ctfa_mode;
        // offline: TA rows of the 31 frames in front of a time chunk, and the rows this sub-U-Net must hand to the next chunk
        const size_t hist_off = P.streaming ? 0 : P.carry_alloc((CTFA_WINDOW - 1) * 64);
        std::vector<std::pair<Ten*, size_t>> saves;
        saves.swap(P.pending_save);
        if ((int)saves.size() > CARRY_MAX) fail(NUNET_EINVAL, "plan: %s carries %zu rows", blk.c_str(), saves.size());
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = blk;
            const int frames = r.B * r.T;
            if (pp->streaming && pp->sh16 && !ring && !fuse_out_conv) {
                // one-frame graph: mean, both MLPs, gate and residual of a stream in one launch
                ctfa_stream_sh_kernel<<<r.B, 256, 0, r.st>>>(reinterpret_cast<const uint8_t*>(pp->cur(x, r.parity)),
                                                            reinterpret_cast<const uint8_t*>(pp->cur(en_in, r.parity)), E.mlpw(mta), E.mlpw(mfa),
                                                            reinterpret_cast<uint8_t*>(pp->cur(out, r.parity)), F0, in_eo, o_eo);
                E.check_launch("ctfa_stream", frames * 4.0 * (3.0 * F0 * 64));
                return;
            }
            if (pp->sh16 && F0 <= 64)
                ctfa_ta_warp_sh_kernel<<<std::min((frames + 31) / 32, E.num_sms * 3), 256, 0, r.st>>>(
                    reinterpret_cast<const uint8_t*>(pp->cur(x, r.parity)), E.mlpw(mta), pp->cur(ta, 0), F0, (long long)frames);
            else if (pp->sh16)
                ctfa_ta_sh_kernel<<<(frames + CTFA_FPB - 1) / CTFA_FPB, 256, 0, r.st>>>(
                    reinterpret_cast<const uint8_t*>(pp->cur(x, r.parity)), E.mlpw(mta), pp->cur(ta, 0), F0, (long long)frames);
            else
                ctfa_ta_kernel<<<frames, 256, 0, r.st>>>(pp->cur(x, r.parity), E.mlpw(mta), pp->cur(ta, 0), F0);
            E.check_launch("ctfa_ta", frames * 4.0 * (F0 * 64 + 64));
            const int div32 = pp->streaming ? 1 : (off_mode == NUNET_CTFA_FRAME_DIV32);
            if (!pp->streaming && (r.use_carry || r.save_carry) && !pp->sh16) fail(NUNET_EINVAL, "time chunking needs the tensor-core plan");
            if (!pp->streaming && r.save_carry && saves.size()) {
                CarrySave cs{};
                cs.n = (int)saves.size();
                for (int i = 0; i < cs.n; ++i) {
                    cs.src[i] = reinterpret_cast<const uint8_t*>(pp->cur(saves[i].first, r.parity));
                    cs.dst[i] = reinterpret_cast<uint8_t*>(pp->carry_at(saves[i].second));
                    cs.row16[i] = (int)(saves[i].first->numel() / 4);
                }
                carry_save_kernel<<<dim3(cs.n, r.B), 128, 0, r.st>>>(cs, r.T);
                E.check_launch("carry_save", 0.0);
            }
            const float* hist = (!pp->streaming && r.use_carry && !div32) ? pp->carry_at(hist_off) : nullptr;
            if (!ring && (frames >= 64 || !pp->streaming))
                if (E.ctfa_gate4)
                    ctfa_gate_warp4_kernel<<<std::min((frames + 31) / 32, E.num_sms * 4), 256, 0, r.st>>>(pp->cur(ta, 0), E.mlpw(mfa), pp->cur(gate, 0),
                                                                                                         r.T, div32, (long long)frames, hist, r.t0);
                else
                ctfa_gate_warp_kernel<<<std::min((frames + 7) / 8, E.num_sms * 8), 256, 0, r.st>>>(pp->cur(ta, 0), E.mlpw(mfa), pp->cur(gate, 0),
                                                                                                    r.T, div32, (long long)frames, hist, r.t0);
            else
                ctfa_gate_kernel<<<frames, 64, 0, r.st>>>(pp->cur(ta, 0), E.mlpw(mfa), pp->cur(gate, 0), r.T, div32,
                                                         ring ? pp->cur(ring, 0) : nullptr, r.ring_pos);
            E.check_launch("ctfa_gate", frames * 4.0 * (64 + 64));
            if (!pp->streaming && r.save_carry && !div32) {
                ctfa_hist_update_kernel<<<r.B, 64, 0, r.st>>>(pp->cur(ta, 0), pp->carry_at(hist_off), r.T, r.use_carry ? 1 : 0);
                E.check_launch("ctfa_hist", 0.0);
            }
            const long long n4 = (long long)frames * F0 * 16;
            if (pp->sh16 && fuse_out_conv) {
                // last decoder block: the 64 -> 1 out_conv is applied on the fly, the block output is never written
                const long long npix = (long long)frames * F0;
                gate_residual_out_conv_sh_kernel<<<(int)((npix + 127) / 128), 128, 0, r.st>>>(
                    reinterpret_cast<const uint8_t*>(pp->cur(x, r.parity)), reinterpret_cast<const uint8_t*>(pp->cur(en_in, r.parity)),
                    pp->cur(gate, 0), E.pool.at(E.out_layer.w), E.pool.at(E.out_layer.b), r.est_out, npix, F0, in_eo, r.est_stride,
                    r.est_off);
                E.check_launch("gate_residual_out_conv", frames * 4.0 * (3.0 * F0 * 64 + 64) + npix * 4.0 * 65);
                return;
            }
            if (pp->sh16) {
                const long long n8 = n4 / 2;
                const int blocks8 = (int)std::min<long long>((n8 + 255) / 256, 148LL * 16);
                gate_residual_sh_kernel<<<blocks8, 256, 0, r.st>>>(reinterpret_cast<const uint8_t*>(pp->cur(x, r.parity)),
                                                                  reinterpret_cast<const uint8_t*>(pp->cur(en_in, r.parity)),
                                                                  pp->cur(gate, 0),
                                                                  reinterpret_cast<uint8_t*>(pp->cur(out, r.parity)), n8, F0,
                                                                  in_eo, o_eo);
            } else {
                const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148LL * 16);
                gate_residual_kernel<<<blocks, 256, 0, r.st>>>(reinterpret_cast<const float4*>(pp->cur(x, r.parity)),
                                                              reinterpret_cast<const float4*>(pp->cur(en_in, r.parity)),
                                                              reinterpret_cast<const float4*>(pp->cur(gate, 0)),
                                                              reinterpret_cast<float4*>(pp->cur(out, r.parity)), n4, F0);
            }
            E.check_launch("gate_residual", frames * 4.0 * (3.0 * F0 * 64 + 64));
        });
        return out;
    }

    void build_plan(Plan& P) {
        Plan* pp = &P;
        const bool recycle = !P.streaming && !no_recycle;
        Ten* x0 = P.make("input_layer", 256, 64, false);
        x0->sh = P.sh16;
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = "input_layer";
            const long long npix = (long long)r.B * r.T * 256;
            const int blocks = (int)((npix * 8 + 255) / 256);
            const VecLayer& v = E.in_layer;
            if (pp->sh16)
                input_layer_sh_kernel<<<(int)((npix + 127) / 128), 128, 0, r.st>>>(
                    r.mag_in, E.pool.at(v.w), E.pool.at(v.b), E.pool.at(v.gamma), E.pool.at(v.beta), E.pool.at(v.alpha),
                    reinterpret_cast<uint8_t*>(pp->cur(x0, r.parity)), npix, 256);
            else
                input_layer_kernel<<<blocks, 256, 0, r.st>>>(r.mag_in, E.pool.at(v.w), E.pool.at(v.b), E.pool.at(v.gamma),
                                                            E.pool.at(v.beta), E.pool.at(v.alpha), pp->cur(x0, r.parity), npix);
            E.check_launch("input_layer", npix * 4.0 * 65);
        });
        Ten* x = x0;
        Ten* enc_des[6][6] = {};
        Ten* enc_out[6] = {};
        for (int i = 0; i < 6; ++i) {
            const std::string blk = ENC_NAMES[i];
            const size_t m = P.mark();
            Ten* en_in = op_conv(P, blk + "_in", x, nullptr, blk + "_in", false, /*out_eo=*/true);
            Ten* out = op_msfe(P, blk, ENC_N[i], en_in, nullptr, enc_des[i], true, false, /*out_eo=*/true);
            x = op_conv(P, DOWN_NAMES[i], out, nullptr, DOWN_NAMES[i], true);
            enc_out[i] = x;
            if (recycle) P.release(m);
        }
        Ten* y = is_ddb() ? op_ddb(P, "ddb", x, "bb_main", true) : op_lstm(P, "lstm", x, "bb_main", "state", true);
        for (int i = 0; i < 6; ++i) {
            const std::string blk = DEC_NAMES[i];
            const int j = 5 - i;
            const size_t m = P.mark();
            Ten* en_in = op_conv(P, blk + "_in", y, enc_out[j], blk + "_in", false, /*out_eo=*/true);
            // the last block's gate + residual also applies out_conv (sh16 plans; kept apart when every tensor is retained
            // for tests/layer_report.py)
            const bool fuse = (i == 5) && P.sh16 && !no_recycle;
            y = op_msfe(P, blk, DEC_N[i], en_in, enc_des[j], nullptr, false, true, /*out_eo=*/false, fuse);
            if (recycle) P.release(m);
        }
        if (!(P.sh16 && !no_recycle))
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = "out_conv";
            const long long npix = (long long)r.B * r.T * 256;
            const int blocks = (int)((npix * 16 + 255) / 256);
            if (pp->sh16)
                out_conv_sh_kernel<<<(int)((npix + 127) / 128), 128, 0, r.st>>>(
                    reinterpret_cast<const uint8_t*>(pp->cur(y, r.parity)), E.pool.at(E.out_layer.w), E.pool.at(E.out_layer.b),
                    r.est_out, npix, 256, r.est_stride, r.est_off);
            else
                out_conv_kernel<<<blocks, 256, 0, r.st>>>(pp->cur(y, r.parity), E.pool.at(E.out_layer.w), E.pool.at(E.out_layer.b),
                                                         r.est_out, npix, 256, r.est_stride, r.est_off);
            E.check_launch("out_conv", npix * 4.0 * 65);
        });
    }

    // ---- int8-hybrid variant: the one-frame graph operator by operator (hybrid_kernels.cuh), fp32 tensors, streaming only
    Ten* hq_conv_op(Plan& P, const std::string& role, Ten* a, Ten* b, const std::string& out_name) {
        const HqConvLayer L = hconvs.at(role);
        if (a->C + (b ? b->C : 0) != L.Ct || (b && b->F != a->F)) fail(NUNET_EINVAL, "plan: %s wiring", role.c_str());
        const int F_in = a->F;
        const int F_conv = (L.stride == 2) ? F_in / 2 : F_in;
        const bool shuf = (L.epi == HQ_SHUF32 || L.epi == HQ_SHUF64);
        Ten* o = P.make(out_name, shuf ? 2 * F_conv : F_conv, shuf ? L.Cout / 2 : L.Cout, true);
        Plan* pp = &P;
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = out_name;
            int8_t* qb = reinterpret_cast<int8_t*>(pp->cur(E.hq_qbuf, 0));
            HqParams* qp = reinterpret_cast<HqParams*>(pp->cur(E.hq_qp, 0));
            hq_quantize_kernel<<<r.B, 256, 0, r.st>>>(pp->prev(a, r.parity), pp->cur(a, r.parity), b ? pp->prev(b, r.parity) : nullptr,
                                                      b ? pp->cur(b, r.parity) : nullptr, F_in, a->C, b ? b->C : 0, L.KT, qb, qp);
            E.check_launch("hq_quantize", 0.0);
            HqConv c{};
            c.q = qb; c.qp = qp;
            c.w = reinterpret_cast<const int*>(E.pool.at(L.w));
            c.wtap = reinterpret_cast<const int*>(E.pool.at(L.wtap));
            c.wscale = E.pool.at(L.wscale); c.bias = E.pool.at(L.bias);
            c.gamma = E.pool.at(L.gamma); c.beta = E.pool.at(L.beta); c.alpha = E.pool.at(L.alpha);
            c.out = pp->cur(o, r.parity);
            c.F_in = F_in; c.Ct = L.Ct; c.Cout = L.Cout; c.KT = L.KT; c.KF = L.KF; c.padl = L.padl; c.stride = L.stride; c.F_conv = F_conv; c.epi = L.epi;
            const int PT = 128 / L.Cout;
            hq_conv_kernel<<<dim3(r.B, (F_conv + PT - 1) / PT), 128, 128 * sizeof(float), r.st>>>(c);
            E.check_launch("hq_conv", (double)r.B * (F_in * L.Ct * L.KT + 4.0 * o->numel()));
        });
        return o;
    }

    Ten* hq_lstm_op(Plan& P, const std::string& lstm, Ten* x, const std::string& out_name, const std::string& state_name) {
        const HqLstmLayer L = hlstms.at(lstm);
        const int D = x->F * x->C;
        if (D != L.D) fail(NUNET_EINVAL, "plan: %s width", lstm.c_str());
        Ten* hst = P.make("", 1, LSTM_UNITS, true, false);
        Ten* cst = P.make("", 1, LSTM_UNITS, true, false);
        Plan::StateRef sh, sc;
        sh.name = state_name + "_h"; sh.is_lstm = true; sh.a = hst;
        sc.name = state_name + "_c"; sc.is_lstm = true; sc.a = cst;
        P.states.push_back(sh);
        P.states.push_back(sc);
        Ten* o = P.make(out_name, x->F, x->C, true);
        Plan* pp = &P;
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = out_name;
            HqLstm l{};
            l.x = pp->cur(x, r.parity); l.y = pp->cur(o, r.parity);
            l.h = pp->cur(hst, 0); l.c = pp->cur(cst, 0);
            l.wk = reinterpret_cast<const int8_t*>(E.pool.at(L.wk));
            l.wr = reinterpret_cast<const int8_t*>(E.pool.at(L.wr));
            l.wd = L.dense_q ? reinterpret_cast<const int8_t*>(E.pool.at(L.wd)) : nullptr;
            l.wd_f = L.dense_q ? nullptr : E.pool.at(L.wd_f);
            l.bk = E.pool.at(L.bk); l.bd = E.pool.at(L.bd);
            l.sk = L.sk; l.sr = L.sr; l.sd = L.sd; l.D = D;
            hq_lstm_kernel<<<r.B, 128, 0, r.st>>>(l);
            E.check_launch("hq_lstm", (double)r.B * 8.0 * D);
        });
        return o;
    }

    HqMlp hq_mlp_of(const HqMlpLayer& m) const {
        HqMlp o;
        o.k0 = reinterpret_cast<const int8_t*>(pool.at(m.k0)); o.k1 = reinterpret_cast<const int8_t*>(pool.at(m.k1));
        o.s0 = pool.at(m.s0); o.s1 = pool.at(m.s1); o.b0 = pool.at(m.b0); o.b1 = pool.at(m.b1);
        return o;
    }

    Ten* hq_msfe(Plan& P, const std::string& blk, int n, Ten* en_in, Ten* const* skips, Ten** des_out) {
        std::string pc, ps;
        state_prefixes(blk, pc, ps);
        std::vector<Ten*> ens;
        Ten* cur = en_in;
        for (int k = 1; k <= n; ++k) {
            Ten* sk = skips ? skips[k - 1] : nullptr;
            Plan::StateRef st;
            st.name = pc + "_" + std::to_string(k); st.a = cur; st.b = sk;
            P.states.push_back(st);
            cur = hq_conv_op(P, blk + "_conv" + std::to_string(k), cur, sk, blk + "_conv" + std::to_string(k));
            ens.push_back(cur);
        }
        cur = hq_lstm_op(P, blk + "_lstm", cur, blk + "_bb", blk);
        std::vector<Ten*> des;
        for (int k = 1; k <= n; ++k) {
            Ten* sk = ens[n - k];
            Plan::StateRef st;
            st.name = ps + "_" + std::to_string(k); st.a = cur; st.b = sk;
            P.states.push_back(st);
            cur = hq_conv_op(P, blk + "_spconv" + std::to_string(k), cur, sk, blk + "_spconv" + std::to_string(k));
            des.push_back(cur);
        }
        if (des_out)
            for (int k = 1; k <= n; ++k) des_out[k - 1] = des[n - k];
        Ten* x = cur;
        Ten* out = P.make(blk + "_out", x->F, 64, true);
        const HqMlpLayer mta = hmlps.at(blk + "_ta"), mfa = hmlps.at(blk + "_fa");
        const int F0 = x->F;
        Plan* pp = &P;
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = blk;
            hq_ctfa_kernel<<<r.B, 64, 0, r.st>>>(pp->cur(x, r.parity), pp->cur(en_in, r.parity), E.hq_mlp_of(mta), E.hq_mlp_of(mfa),
                                                 pp->cur(out, r.parity), F0);
            E.check_launch("hq_ctfa", (double)r.B * 4.0 * 3 * F0 * 64);
        });
        return out;
    }

    void build_hybrid_plan(Plan& P) {
        Plan* pp = &P;
        hq_qbuf = P.make("", 1, 2 * 256 * 128 / 4, true, false);      // int8 [rows 2][F 256][C 128] at most
        hq_qp = P.make("", 1, 64, true, false);
        Ten* x0 = P.make("input_layer", 256, 64, true);
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = "input_layer";
            const long long npix = (long long)r.B * 256;
            const VecLayer& v = E.in_layer;
            input_layer_kernel<<<(int)((npix * 8 + 255) / 256), 256, 0, r.st>>>(r.mag_in, E.pool.at(v.w), E.pool.at(v.b), E.pool.at(v.gamma),
                                                                              E.pool.at(v.beta), E.pool.at(v.alpha), pp->cur(x0, r.parity), npix);
            E.check_launch("input_layer", npix * 4.0 * 65);
        });
        Ten* x = x0;
        Ten* enc_des[6][6] = {};
        Ten* enc_out[6] = {};
        for (int i = 0; i < 6; ++i) {
            const std::string blk = ENC_NAMES[i];
            Ten* en_in = hq_conv_op(P, blk + "_in", x, nullptr, blk + "_in");
            Ten* out = hq_msfe(P, blk, ENC_N[i], en_in, nullptr, enc_des[i]);
            x = hq_conv_op(P, DOWN_NAMES[i], out, nullptr, DOWN_NAMES[i]);
            enc_out[i] = x;
        }
        Ten* y = hq_lstm_op(P, "lstm", x, "bb_main", "state");
        for (int i = 0; i < 6; ++i) {
            const std::string blk = DEC_NAMES[i];
            const int j = 5 - i;
            // up_sampling(cat[y, enc_out]): float transpose conv on the dequantised kernel, then the hybrid 1x1 inconv
            Ten* a = y;
            Ten* b = enc_out[j];
            Ten* u = P.make(std::string(UP_NAMES[i]), 2 * a->F, 128, true);
            const auto up = hups.at(UP_NAMES[i]);
            const int F_in = a->F;
            const std::string un = UP_NAMES[i];
            P.ops.push_back([=](Engine& E, const Run& r) {
                E.cur_op = un;
                hq_upsample_kernel<<<dim3(r.B, 2 * F_in), 128, 0, r.st>>>(pp->cur(a, r.parity), pp->cur(b, r.parity), E.pool.at(up.first),
                                                                         E.pool.at(up.second), pp->cur(u, r.parity), F_in);
                E.check_launch("hq_upsample", (double)r.B * 4.0 * 3 * F_in * 128);
            });
            Ten* en_in = hq_conv_op(P, blk + "_in", u, nullptr, blk + "_in");
            y = hq_msfe(P, blk, DEC_N[i], en_in, enc_des[j], nullptr);
        }
        P.ops.push_back([=](Engine& E, const Run& r) {
            E.cur_op = "out_conv";
            const long long npix = (long long)r.B * 256;
            out_conv_kernel<<<(int)((npix * 16 + 255) / 256), 256, 0, r.st>>>(pp->cur(y, r.parity), E.pool.at(E.out_layer.w), E.pool.at(E.out_layer.b),
                                                                             r.est_out, npix, 256, r.est_stride, r.est_off);
            E.check_launch("out_conv", npix * 4.0 * 65);
        });
    }

    void alloc_plan(Plan& P, int cap, bool streaming) {
        P.streaming = streaming;
        P.sh16 = use_tc && (!streaming || stream_tc3) && !is_hybrid();
        P.cap = cap;
        if (is_hybrid()) {
            if (!streaming) fail(NUNET_EINVAL, "the int8-hybrid variant is the deployed ONE-FRAME graph: streaming entry points only (max_frames must be 0)");
            build_hybrid_plan(P);
        } else {
            build_plan(P);
        }
        if (streaming) {
            s_mag = P.make("", 1, 256, true, false);
            s_ph = P.make("", 1, 2 * NBINS, true, false);
            s_est = P.make("", 1, 256, true, false);
            s_inbuf = P.make("", 1, NFFT, true, false);
            s_outbuf = P.make("", 1, NFFT, true, false);
        } else {
            o_mag = P.make("mag", 1, 256, true);
            o_ph = P.make("", 1, 2 * NBINS, true);
            o_est = P.make("est", 1, NBINS, true);
            o_frames = P.make("", 1, NFFT, true);
        }
        P.finalize();
        const size_t bytes = P.unit_floats * (size_t)cap * sizeof(float);
        cudaError_t e = cudaMalloc(&P.arena, bytes);
        if (e != cudaSuccess) fail(NUNET_ENOMEM, "cudaMalloc of the %s arena (%.2f GB) failed: %s",
                                   streaming ? "streaming" : "offline", bytes / 1e9, cudaGetErrorString(e));
        CUDA_OK(cudaMemset(P.arena, 0, bytes));
        if (!streaming && P.sh16 && P.ctop) {
            // clips whose history can be carried at once: a forced chunk length packs max_frames / chunk_frames clips into a
            // chunk; without one, time chunking only happens for clips longer than max_frames, one clip at a time
            P.carry_cap = cfg.chunk_frames > 0 ? std::max(1, cap / cfg.chunk_frames) : 1;
            const size_t cbytes = P.ctop * (size_t)P.carry_cap * sizeof(float);
            e = cudaMalloc(&P.carry, cbytes);
            if (e != cudaSuccess) fail(NUNET_ENOMEM, "cudaMalloc of the carry arena (%.2f GB) failed: %s", cbytes / 1e9, cudaGetErrorString(e));
            CUDA_OK(cudaMemset(P.carry, 0, cbytes));
        }
    }

    // -------------------------------------------------------------------------------- public operations
    void run_plan(Plan& P, const Run& r) {
        for (auto& op : P.ops) op(*this, r);
    }

    // How a call of B clips x T frames is cut to fit the arena (max_frames): sub-batches of Bs whole clips when a clip fits,
    // and time chunks of Tc frames with carried history (conv rows, LSTM h / c, the 31-frame TA window) when it does not or
    // when cfg.chunk_frames asks for it.  Chunked results are bit-identical to unchunked ones: every output position is
    // computed by the same instruction sequence whatever tile it lands in.
    struct Cut {
        int Bs, Tc;
    };
    Cut cut_call(int B, int T) const {
        if (!offline.arena) fail(NUNET_EINVAL, "offline path disabled (max_frames = 0)");
        if (B <= 0 || T <= 0) fail(NUNET_EINVAL, "bad shape B=%d T=%d", B, T);
        Cut c;
        c.Tc = cfg.chunk_frames > 0 ? std::min(cfg.chunk_frames, T) : T;
        c.Tc = std::min(c.Tc, offline.cap);
        c.Bs = std::max(1, std::min(B, offline.cap / c.Tc));
        if (c.Tc < T) {
            if (!offline.carry) fail(NUNET_ENOMEM, "B*T = %lld exceeds max_frames = %d and this plan cannot carry history between time chunks",
                                     (long long)B * T, offline.cap);
            c.Bs = std::min(c.Bs, offline.carry_cap);
        }
        return c;
    }

    // one chunk of the network: mag [Bc*Tc][256] dense -> out (frame-dense with stride / offset)
    void net_chunk(const float* mag, int Bc, int Tc, int t0, bool more, float* out, int out_stride, int out_off, cudaStream_t st) {
        Run r;
        r.B = Bc; r.T = Tc; r.st = st; r.mag_in = mag; r.est_out = out; r.est_stride = out_stride; r.est_off = out_off;
        r.t0 = t0; r.use_carry = t0 > 0; r.save_carry = more;
        last_B = Bc; last_T = Tc;
        run_plan(offline, r);
    }

    void forward_mag(const float* mag, int B, int T, float* out, cudaStream_t st) {
        const Cut c = cut_call(B, T);
        if (c.Bs >= B && c.Tc >= T) {
            net_chunk(mag, B, T, 0, false, out, 256, 0, st);
            return;
        }
        float* mg = offline.cur(o_mag, 0);    // [frames][256] staging of a chunk's input
        float* es = offline.cur(o_est, 0);    // [frames][257] (column 0 unused)
        for (int b0 = 0; b0 < B; b0 += c.Bs) {
            const int Bc = std::min(c.Bs, B - b0);
            for (int t0 = 0; t0 < T; t0 += c.Tc) {
                const int Tc = std::min(c.Tc, T - t0);
                CUDA_OK(cudaMemcpy2DAsync(mg, (size_t)Tc * 256 * 4, mag + ((size_t)b0 * T + t0) * 256, (size_t)T * 256 * 4, (size_t)Tc * 256 * 4,
                                          Bc, cudaMemcpyDeviceToDevice, st));
                net_chunk(mg, Bc, Tc, t0, t0 + Tc < T, es, NBINS, 1, st);
                for (int b = 0; b < Bc; ++b)      // drop the DC column: [Tc][257] -> [Tc][256] of clip b0 + b
                    CUDA_OK(cudaMemcpy2DAsync(out + ((size_t)(b0 + b) * T + t0) * 256, 256 * 4, es + (size_t)b * Tc * NBINS + 1, NBINS * 4, 256 * 4, Tc,
                                              cudaMemcpyDeviceToDevice, st));
            }
        }
    }

    void forward_wav(const float* wav, int B, int n, float* out_wav, float* out_mag, cudaStream_t st) {
        const int T = nunet_num_frames(n);
        if (B <= 0) fail(NUNET_EINVAL, "bad batch size B=%d", B);
        if (T <= 0) fail(NUNET_EINVAL, "clip shorter than one 512-sample frame");
        const Cut c = cut_call(B, T);
        launches = 0;
        order_begin(st);
        prof_begin(st);
        float* mag = offline.cur(o_mag, 0);
        float2* ph = reinterpret_cast<float2*>(offline.cur(o_ph, 0));
        float* est = offline.cur(o_est, 0);   // [frames][257]
        float* fr = offline.cur(o_frames, 0);
        const long long n_out = (long long)(T - 1) * HOP + NFFT;
        for (int b0 = 0; b0 < B; b0 += c.Bs) {
            const int Bc = std::min(c.Bs, B - b0);
            for (int t0 = 0; t0 < T; t0 += c.Tc) {
                const int Tc = std::min(c.Tc, T - t0);
                const long long frames = (long long)Bc * Tc;
                const int fblocks = (int)((frames + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA);
                cur_op = "framing";
                stft_kernel<<<fblocks, 32 * FRAMES_PER_CTA, 0, st>>>(wav + (size_t)b0 * n + (size_t)t0 * HOP, tables(false), mag, ph, Bc, Tc, n);
                check_launch("stft", frames * 4.0 * (256 + 256 + 514));
                net_chunk(mag, Bc, Tc, t0, t0 + Tc < T, est, NBINS, 1, st);
                if (out_mag)
                    CUDA_OK(cudaMemcpy2DAsync(out_mag + ((size_t)b0 * T + t0) * NBINS, (size_t)T * NBINS * 4, est, (size_t)Tc * NBINS * 4,
                                              (size_t)Tc * NBINS * 4, Bc, cudaMemcpyDeviceToDevice, st));
                cur_op = "framing";
                if (out_wav) {
                    istft_frames_kernel<<<fblocks, 32 * FRAMES_PER_CTA, 0, st>>>(est, ph, tables(false), fr, frames);
                    check_launch("istft_frames", frames * 4.0 * (257 + 514 + 512));
                    // samples [t0 * 256, (t0 + Tc) * 256 + 256) of every clip: the first hop of a later chunk adds to what the
                    // previous chunk's last frame left there
                    const long long span = (long long)Tc * HOP + HOP;
                    const int blocks = (int)std::min<long long>((Bc * span + 255) / 256, 148LL * 16);
                    overlap_add_kernel<<<blocks, 256, 0, st>>>(fr, out_wav + (size_t)b0 * n_out + (size_t)t0 * HOP, Bc, Tc, span, n_out, t0 > 0 ? 1 : 0);
                    check_launch("overlap_add", frames * 4.0 * (512 + 256));
                }
            }
        }
        order_end(st);
    }

    void check_streams(int S) {
        if (!stream.arena) fail(NUNET_EINVAL, "streaming path disabled (max_streams = 0)");
        if (S != stream.cap) fail(NUNET_EINVAL, "a step must cover all max_streams = %d streams (got S = %d)", stream.cap, S);
    }

    // S may be a sub-range of the streams (graph chains); the entry points check the full count
    void stream_step_mag(const float* mag, int S, float* out, cudaStream_t st) {
        Run r;
        r.B = S; r.T = 1; r.st = st; r.mag_in = mag; r.est_out = out; r.est_stride = 256; r.est_off = 0;
        r.parity = stream_parity ^ 1;
        r.ring_pos = stream_steps & (CTFA_WINDOW - 1);
        r.step = stream_steps;
        run_plan(stream, r);
    }
    void stream_advance() {
        stream_parity ^= 1;
        ++stream_steps;
        ++state_gen;
    }

    // ---- CUDA graphs for the streaming step.  One step is ~200 small kernels whose arguments depend only on the step parity
    // (and on the ring slot for the DDB / attention-history rings), the stream count and the caller's buffers: each such
    // combination is captured once (on a private stream, so that callers may use the legacy default stream) and replayed
    // with one cudaGraphLaunch.  The first two steps run eagerly (lazy function attributes, occupancy queries).
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        int launches = 0;
    };
    std::map<std::tuple<int, int, int, int>, StepGraph> step_graphs;
    int stream_graphs = 1;          // NUNET_STREAM_GRAPH=0: always launch kernel by kernel
    int stream_split = 1;           // NUNET_STREAM_SPLIT=2..4 (experiments): parallel chains per captured step, from 64 streams per
                                    // chain.  Measured at 1024 streams: 2.69 ms (1 chain), 2.90 ms (2), 4.48 ms (4) -- the 1-CTA-per-SM
                                    // conv kernels of different chains do not overlap, so more chains only add launches
    int eager_steps = 0;
    cudaStream_t cap_stream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t cap_event[4] = {nullptr, nullptr, nullptr, nullptr};

    // body(stream, first, count) enqueues one step for streams [first, first + count) WITHOUT advancing the step counters.
    template <typename Body>
    void graph_step(int kind, int S, cudaStream_t st, Body&& body) {
        const bool rings = is_ddb() || cfg.stream_ctfa_history;
        if (!stream_graphs || prof_on || tc3_timing_buf || eager_steps < 2) {
            ++eager_steps;
            body(st, 0, S);
            stream_advance();
            return;
        }
        const auto key = std::make_tuple(kind, S, stream_parity, rings ? (stream_steps & (DDB_RING - 1)) : 0);
        auto it = step_graphs.find(key);
        if (it == step_graphs.end()) {
            // Streams are independent and most kernels of a step are far too small to fill the GPU, so a step of many
            // streams is captured as several parallel chains over disjoint stream ranges: the graph lets kernels of
            // different chains run side by side.
            const int nsplit = (S >= 64 * stream_split) ? std::min(stream_split, 4) : 1;
            if (!cap_stream[0])
                for (int i = 0; i < 4; ++i) {
                    CUDA_OK(cudaStreamCreateWithFlags(&cap_stream[i], cudaStreamNonBlocking));
                    CUDA_OK(cudaEventCreateWithFlags(&cap_event[i], cudaEventDisableTiming));
                }
            cudaGraph_t g = nullptr;
            StepGraph sg;
            std::string why;
            bool ok = cudaStreamBeginCapture(cap_stream[0], cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                try {
                    launches = 0;
                    if (nsplit > 1) CUDA_OK(cudaEventRecord(cap_event[0], cap_stream[0]));
                    for (int i = 1; i < nsplit; ++i) CUDA_OK(cudaStreamWaitEvent(cap_stream[i], cap_event[0], 0));
                    for (int i = 0; i < nsplit; ++i) {
                        const int first = (int)((long long)S * i / nsplit), last = (int)((long long)S * (i + 1) / nsplit);
                        stream.unit0 = first;
                        body(cap_stream[i], first, last - first);
                    }
                    for (int i = 1; i < nsplit; ++i) {
                        CUDA_OK(cudaEventRecord(cap_event[i], cap_stream[i]));
                        CUDA_OK(cudaStreamWaitEvent(cap_stream[0], cap_event[i], 0));
                    }
                    sg.launches = launches;
                } catch (const std::exception& ex) {
                    why = ex.what();
                    ok = false;
                }
                stream.unit0 = 0;
                const cudaError_t ce = cudaStreamEndCapture(cap_stream[0], &g);
                if (ce != cudaSuccess || !g) {
                    if (why.empty()) why = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce);
                    ok = false;
                }
            } else {
                why = "cudaStreamBeginCapture failed";
            }
            if (ok) {
                const cudaError_t ie = cudaGraphInstantiate(&sg.exec, g, 0);
                if (ie != cudaSuccess) {
                    why = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie);
                    ok = false;
                }
            }
            if (g) cudaGraphDestroy(g);
            if (!ok) {   // capture is an optimisation, never a requirement: fall back to plain launches for good
                fprintf(stderr, "nunet_b200: streaming step not captured as a CUDA graph (%s); launching kernel by kernel\n", why.c_str());
                cudaGetLastError();
                stream_graphs = 0;
                body(st, 0, S);
                stream_advance();
                return;
            }
            it = step_graphs.emplace(key, sg).first;
        }
        CUDA_OK(cudaGraphLaunch(it->second.exec, st));
        launches = it->second.launches;
        stream_advance();
    }

    void stream_step_wav(const float* hop, int S, float* out_hop, float* out_mag, cudaStream_t st) {
        check_streams(S);
        launches = 0;
        cur_op = "framing";
        order_begin(st);
        prof_begin(st);
        if (stream_graphs && !prof_on && !tc3_timing_buf) {
            // the graph works on the engine's own staging buffers, so it does not depend on the caller's pointers
            const size_t n = (size_t)S * HOP * sizeof(float);
            if (hop != h_in) CUDA_OK(cudaMemcpyAsync(h_in, hop, n, cudaMemcpyDeviceToDevice, st));
            graph_step(1, S, st, [&](cudaStream_t s, int first, int count) {
                stream_step_wav_body(h_in + (size_t)first * HOP, count, h_out + (size_t)first * HOP, nullptr, s);
            });
            if (out_hop != h_out) CUDA_OK(cudaMemcpyAsync(out_hop, h_out, n, cudaMemcpyDeviceToDevice, st));
            if (out_mag) CUDA_OK(cudaMemcpyAsync(out_mag, stream.cur(s_est, 0), (size_t)S * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else {
            stream_step_wav_body(hop, S, out_hop, out_mag, st);
            stream_advance();
        }
        order_end(st);
    }
    // one hop of the network alone (magnitudes in, magnitudes out), staged the same way
    void stream_step_mag_api(const float* mag, int S, float* out, cudaStream_t st) {
        check_streams(S);
        if (stream_graphs && !prof_on && !tc3_timing_buf) {
            const size_t n = (size_t)S * 256 * sizeof(float);
            CUDA_OK(cudaMemcpyAsync(stream.cur(s_mag, 0), mag, n, cudaMemcpyDeviceToDevice, st));
            graph_step(0, S, st, [&](cudaStream_t s, int, int count) { stream_step_mag(stream.cur(s_mag, 0), count, stream.cur(s_est, 0), s); });
            CUDA_OK(cudaMemcpyAsync(out, stream.cur(s_est, 0), n, cudaMemcpyDeviceToDevice, st));
        } else {
            stream_step_mag(mag, S, out, st);
            stream_advance();
        }
    }

    void stream_step_wav_body(const float* hop, int S, float* out_hop, float* out_mag, cudaStream_t st) {
        cur_op = "framing";
        float* mag = stream.cur(s_mag, 0);
        float2* ph = reinterpret_cast<float2*>(stream.cur(s_ph, 0));
        float* est = stream.cur(s_est, 0);
        const int fblocks = (S + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA;
        stream_analysis_kernel<<<fblocks, 32 * FRAMES_PER_CTA, 0, st>>>(hop, stream.cur(s_inbuf, 0), tables(true), mag, ph, S);
        check_launch("stream_analysis", S * 4.0 * (256 + 512 + 512 + 256 + 514));
        stream_step_mag(mag, S, est, st);
        if (out_mag) CUDA_OK(cudaMemcpyAsync(out_mag, est, (size_t)S * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        cur_op = "framing";
        stream_synthesis_kernel<<<fblocks, 32 * FRAMES_PER_CTA, 0, st>>>(est, ph, tables(true), stream.cur(s_outbuf, 0), out_hop, S,
                                                                       cfg.dc_mode == NUNET_DC_EDGE);
        check_launch("stream_synthesis", S * 4.0 * (256 + 514 + 512 + 512 + 256));
    }

    void stream_reset(int first, int count, cudaStream_t st) {
        if (!stream.arena) fail(NUNET_EINVAL, "streaming path disabled (max_streams = 0)");
        if (first < 0 || count < 0 || first + count > stream.cap) fail(NUNET_EINVAL, "stream range out of bounds");
        ++state_gen;
        order_begin(st);
        if (first == 0 && count == stream.cap) {
            CUDA_OK(cudaMemsetAsync(stream.arena, 0, stream.unit_floats * (size_t)stream.cap * sizeof(float), st));
            order_end(st);
            return;
        }
        // every per-stream region is [cap][n] at offset off*cap: zero rows [first, first+count) of each
        for (auto& t : stream.tens) {
            const size_t n = t->numel();
            CUDA_OK(cudaMemsetAsync(stream.ptr(t->off[0]) + (size_t)first * n, 0, (size_t)count * n * sizeof(float), st));
            if (t->pingpong)
                CUDA_OK(cudaMemsetAsync(stream.ptr(t->off[1]) + (size_t)first * n, 0, (size_t)count * n * sizeof(float), st));
        }
        order_end(st);
    }

    const Plan::StateRef* find_state(const std::string& name) const {
        for (auto& s : stream.states)
            if (s.name == name) return &s;
        return nullptr;
    }
    int state_numel(const Plan::StateRef& s) const {
        if (s.is_lstm) return LSTM_UNITS;
        if (s.ddb_k) return s.ddb_d * (s.a->F / DDB_RING) * s.a->C * s.ddb_k;
        return s.a->F * (s.a->C + (s.b ? s.b->C : 0));
    }
    // Framing state that is not part of the signature's tensor list but belongs to a stream checkpoint: the analysis
    // window `in_buffer` and the overlap-add accumulator `out_buffer` of the frame loop (interpreter_proposed.py:30-31),
    // and, when the engine keeps real attention history (stream_ctfa_history), the 32-frame rings "ctfa_ring<i>"
    // (oldest frame first).  Returns the element count, or -1 for other names.
    int extra_state(const std::string& name, const Ten** ten, int* ring) const {
        *ring = 0;
        if (name == "in_buffer") { *ten = s_inbuf; return s_inbuf ? NFFT : -1; }
        if (name == "out_buffer") { *ten = s_outbuf; return s_outbuf ? NFFT : -1; }
        if (name.rfind("ctfa_ring", 0) == 0) {
            const int i = atoi(name.c_str() + 9);
            if (i < 0 || i >= (int)stream.rings.size() || name != "ctfa_ring" + std::to_string(i)) return -1;
            *ten = stream.rings[i];
            *ring = 1;
            return CTFA_WINDOW * 64;
        }
        return -1;
    }
    void state_xfer(int sid, const std::string& name, float* buf, bool to_host) {
        if (!stream.arena) fail(NUNET_EINVAL, "streaming path disabled (max_streams = 0)");
        if (sid < 0 || sid >= stream.cap) fail(NUNET_EINVAL, "stream id out of range");
        if (!to_host) ++state_gen;
        const Ten* xt = nullptr;
        int xring = 0;
        const int xn = extra_state(name, &xt, &xring);
        if (xn > 0) {
            CUDA_OK(cudaDeviceSynchronize());
            float* dev = stream.cur(xt, 0) + (size_t)sid * xn;
            for (int r = 0; r < (xring ? CTFA_WINDOW : 1); ++r) {
                // ring row r (oldest first) lives in slot (steps + r) mod 32: slot steps mod 32 is overwritten next
                const int slot = xring ? ((stream_steps + r) & (CTFA_WINDOW - 1)) : 0;
                const size_t n = xring ? 64 : (size_t)xn;
                if (to_host) CUDA_OK(cudaMemcpy(buf + (size_t)r * n, dev + (size_t)slot * n, n * sizeof(float), cudaMemcpyDeviceToHost));
                else CUDA_OK(cudaMemcpy(dev + (size_t)slot * n, buf + (size_t)r * n, n * sizeof(float), cudaMemcpyHostToDevice));
            }
            return;
        }
        const Plan::StateRef* s = find_state(name);
        if (!s) fail(NUNET_ESTATE, "unknown state tensor '%s'", name.c_str());
        CUDA_OK(cudaDeviceSynchronize());
        auto xfer = [&](float* dev, float* host, size_t width, size_t hpitch, size_t rows) {
            if (to_host) CUDA_OK(cudaMemcpy2D(host, hpitch * 4, dev, width * 4, width * 4, rows, cudaMemcpyDeviceToHost));
            else CUDA_OK(cudaMemcpy2D(dev, width * 4, host, hpitch * 4, width * 4, rows, cudaMemcpyHostToDevice));
        };
        if (s->is_lstm) {
            xfer(stream.cur(s->a, 0) + (size_t)sid * LSTM_UNITS, buf, LSTM_UNITS, LSTM_UNITS, 1);
            return;
        }
        if (s->ddb_k) {
            // reference tensor [d rows (oldest first)][F][k*h], channels = cat[out_{k-1}, .., out_0]; row r is step
            // last - (d - 1 - r), the last finished step sits in slot (stream_steps - 1)
            const int k = s->ddb_k, d = s->ddb_d, h = s->a->C, F = s->a->F / DDB_RING;
            for (int r = 0; r < d; ++r) {
                const int slot = (stream_steps - 1 - (d - 1 - r)) & (DDB_RING - 1);
                for (int m = 0; m < k; ++m) {
                    float* ring = stream.cur(s->rings[k - 1 - m], 0) + ((size_t)sid * DDB_RING + slot) * F * h;
                    xfer(ring, buf + (size_t)r * F * k * h + (size_t)m * h, h, (size_t)k * h, F);
                }
            }
            return;
        }
        const int CA = s->a->C, CB = s->b ? s->b->C : 0, F = s->a->F;
        auto one = [&](const Ten* tn, int coff) {
            float* dev = stream.cur(tn, stream_parity) + (size_t)sid * tn->numel();
            const int C = tn->C;
            if (!tn->sh) {
                xfer(dev, buf + coff, C, CA + CB, F);
                return;
            }
            // sh16 frame row: planar [hi|lo][chunk][position][8 halves], positions natural or [even | odd]
            std::vector<__half> row((size_t)2 * F * C);
            if (!to_host) CUDA_OK(cudaMemcpy(row.data(), dev, (size_t)4 * F * C, cudaMemcpyDeviceToHost));   // keep nothing stale
            if (to_host) CUDA_OK(cudaMemcpy(row.data(), dev, (size_t)4 * F * C, cudaMemcpyDeviceToHost));
            for (int f = 0; f < F; ++f)
                for (int c = 0; c < C; ++c) {
                    const int pos = tn->eo ? (f & 1) * (F >> 1) + (f >> 1) : f;
                    const size_t hi = ((size_t)(c >> 3) * F + pos) * 8 + (c & 7);
                    const size_t lo = ((size_t)((C >> 3) + (c >> 3)) * F + pos) * 8 + (c & 7);
                    float& v = buf[(size_t)f * (CA + CB) + coff + c];
                    if (to_host) {
                        v = __half2float(row[hi]) + __half2float(row[lo]);
                    } else {
                        const __half h = __float2half_rn(v);
                        row[hi] = h;
                        row[lo] = __float2half_rn(v - __half2float(h));
                    }
                }
            if (!to_host) CUDA_OK(cudaMemcpy(dev, row.data(), (size_t)4 * F * C, cudaMemcpyHostToDevice));
        };
        one(s->a, 0);
        if (s->b) one(s->b, CA);
    }
};

}  // namespace nunet

using namespace nunet;

struct nunet_engine {
    Engine e;
};

// Every handle-taking entry point runs with the engine's device current and restores the caller's device on exit, so
// engines on different GPUs can be used alternately from any host thread.
struct DeviceGuard {
    int prev = -1, dev = -1;
    explicit DeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

template <typename Fn>
static int guarded(Fn&& fn) {
    try {
        fn();
        g_err.clear();
        return NUNET_OK;
    } catch (const Error& e) {
        g_err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_err = e.what();
        return NUNET_EINVAL;
    }
}
// guarded() for entry points that take a handle: null check + device guard
template <typename Fn>
static int guarded_h(nunet_engine* h, Fn&& fn) {
    if (!h) {
        g_err = "null engine handle";
        return NUNET_EINVAL;
    }
    DeviceGuard dg(h->e.cfg.device);
    return guarded(fn);
}

extern "C" {

const char* nunet_last_error(void) { return g_err.c_str(); }
int nunet_abi_version(void) { return NUNET_ABI_VERSION; }

int nunet_num_frames(int n_samples) { return n_samples < NFFT ? 0 : 1 + (n_samples - NFFT) / HOP; }

long long nunet_blob_validate(const void* blob, size_t blob_bytes, int variant) {
    long long n = 0;
    int rc = guarded([&] {
        if (!blob) fail(NUNET_EINVAL, "null argument");
        Engine E;
        E.cfg.variant = variant;
        E.blob.parse(blob, blob_bytes);
        E.pack_params();
        n = (long long)E.pool.host.size();
    });
    return rc == NUNET_OK ? n : rc;
}

int nunet_create(const nunet_config* cfg, const void* blob, size_t blob_bytes, nunet_engine** out) {
    if (out) *out = nullptr;
    std::unique_ptr<nunet_engine> h;
    int rc = guarded([&] {
        if (!cfg || !blob || !out) fail(NUNET_EINVAL, "nunet_create: null argument");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) fail(NUNET_ENODEV, "no CUDA device");
        if (cfg->device < 0 || cfg->device >= ndev) fail(NUNET_ENODEV, "device %d not present (%d devices)", cfg->device, ndev);
        DeviceGuard dg(cfg->device);     // the caller's current device is restored on return
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
        if (prop.major != 10) fail(NUNET_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
        if (cfg->max_frames < 0 || cfg->max_streams < 0 || (cfg->max_frames == 0 && cfg->max_streams == 0))
            fail(NUNET_EINVAL, "max_frames / max_streams");
        h.reset(new nunet_engine());
        Engine& E = h->e;
        E.cfg = *cfg;
        E.num_sms = prop.multiProcessorCount;
        // Experiment switches (ablations for tools/ and the graph-vs-plain-launch tests).  They are ignored unless
        // NUNET_DEBUG_KNOBS=1 is set, so a stray environment variable cannot change kernels or numerics in production.
        const bool knobs = getenv("NUNET_DEBUG_KNOBS") && atoi(getenv("NUNET_DEBUG_KNOBS")) != 0;
        auto knob = [&](const char* name) -> const char* { return knobs ? getenv(name) : nullptr; };
        E.no_recycle = knob("NUNET_NO_RECYCLE") != nullptr;
        if (const char* v = knob("NUNET_TC3_TWIN")) E.tc3_twin = atoi(v) != 0;
        if (const char* v = knob("NUNET_CTFA_GATE4")) E.ctfa_gate4 = atoi(v) != 0;
        if (const char* v = knob("NUNET_LSTM_STREAM")) E.lstm_stream = atoi(v) != 0;
        if (const char* c = knob("NUNET_CONV")) {
            E.use_tc = strcmp(c, "simt") != 0;
        }
        if (const char* c = knob("NUNET_TC3_FENCE")) E.tc3_fence_mode = atoi(c);
        if (const char* c = knob("NUNET_TC3_DBG")) E.tc3_dbg = atoi(c);
        if (knob("NUNET_TC3_TIMING")) {
            CUDA_OK(cudaMalloc(&E.tc3_timing_buf, 16 * sizeof(unsigned long long)));
            CUDA_OK(cudaMemset(E.tc3_timing_buf, 0, 16 * sizeof(unsigned long long)));
        }
        if (const char* c = knob("NUNET_TC3_TMA")) E.tc3_tma = atoi(c);
        if (const char* c = knob("NUNET_TC3_CLUSTER")) E.tc3_cluster = atoi(c);
        if (const char* c = knob("NUNET_TC3_BOX_MINF")) E.tc3_box_minf = atoi(c);
        if (const char* c = knob("NUNET_TC3_PAIR")) E.tc3_pair = atoi(c);
        if (const char* c = knob("NUNET_TC3_ROW_TILES")) E.tc3_row_tiles = atoi(c);
        if (const char* c = knob("NUNET_STREAM_GRAPH")) E.stream_graphs = atoi(c);
        if (const char* c = knob("NUNET_STREAM_FUSE")) E.stream_fuse = atoi(c);
        if (const char* c = knob("NUNET_FZ_DBG")) E.fz_dbg = atoi(c);
        if (const char* c = knob("NUNET_STREAM_SPLIT")) E.stream_split = std::max(1, std::min(4, atoi(c)));
        if (const char* c = knob("NUNET_TC3_PAIR_MINF")) E.tc3_pair_minf = atoi(c);
        if (const char* c = knob("NUNET_TC3_LD_RR")) E.tc3_ld_rr = atoi(c) != 0;
        if (const char* c = knob("NUNET_TC3_BOX_STRIDED")) E.tc3_box_strided = atoi(c);
        if (const char* c = knob("NUNET_TC3_TMA_MINF")) E.tc3_tma_minf = std::max(8, atoi(c));
        if (const char* c = knob("NUNET_TC3_MT")) E.tc3_force_mt = atoi(c);
        if (const char* c = knob("NUNET_STREAM_CONV")) E.stream_tc3 = strcmp(c, "simt") != 0;
        if (const char* c = knob("NUNET_TC3_PDL")) E.tc3_pdl = atoi(c) != 0;
        E.blob.parse(blob, blob_bytes);
        E.pack_params();
        E.pool.upload();
        E.blob.m.clear();   // the host blob is not referenced after packing
        if (cfg->max_frames > 0) E.alloc_plan(E.offline, cfg->max_frames, false);
        if (cfg->max_streams > 0) E.alloc_plan(E.stream, cfg->max_streams, true);
        CUDA_OK(cudaStreamCreateWithFlags(&E.own_stream, cudaStreamNonBlocking));
        // staging for the host entry points
        size_t in_f = 0, out_f = 0;
        if (cfg->max_frames > 0) {
            in_f = (size_t)cfg->max_frames * 3 * HOP;      // a clip of T frames has < (T + 2) * HOP <= 3 T HOP samples
            out_f = in_f;
        }
        if (cfg->max_streams > 0) {
            in_f = std::max(in_f, (size_t)cfg->max_streams * HOP);
            out_f = std::max(out_f, (size_t)cfg->max_streams * HOP);
        }
        CUDA_OK(cudaMalloc(&E.h_in, in_f * sizeof(float)));
        CUDA_OK(cudaMalloc(&E.h_out, out_f * sizeof(float)));
        E.h_in_cap = in_f;
        E.h_out_cap = out_f;
        CUDA_OK(cudaDeviceSynchronize());
    });
    if (rc == NUNET_OK) *out = h.release();
    return rc;
}

void nunet_destroy(nunet_engine* h) {
    if (!h) return;
    DeviceGuard dg(h->e.cfg.device);
    cudaDeviceSynchronize();
    delete h;
}

int nunet_forward_wav_dev(nunet_engine* h, const float* wav, int B, int n_samples, float* out_wav, float* out_mag,
                          void* stream) {
    return guarded_h(h, [&] {
        if (!h || !wav) fail(NUNET_EINVAL, "null argument");
        h->e.forward_wav(wav, B, n_samples, out_wav, out_mag, static_cast<cudaStream_t>(stream));
    });
}

int nunet_forward_wav_host(nunet_engine* h, const float* wav, int B, int n_samples, float* out_wav, float* out_mag) {
    return guarded_h(h, [&] {
        if (!h || !wav) fail(NUNET_EINVAL, "null argument");
        Engine& E = h->e;
        const int T = nunet_num_frames(n_samples);
        if (B <= 0) fail(NUNET_EINVAL, "bad batch size B=%d", B);
        if (T <= 0) fail(NUNET_EINVAL, "clip shorter than one 512-sample frame");
        // The call is staged through the handle's device buffers in sub-batches of whole clips.  The buffers are sized for
        // max_frames frames at create time; a single clip that is longer than that (a clip cut into time chunks) makes them
        // grow -- the one place the library allocates after nunet_create.
        const size_t clip_in = (size_t)n_samples, clip_out = (size_t)(T - 1) * HOP + NFFT;
        cudaStream_t st = E.own_stream;
        if (clip_in > E.h_in_cap || clip_out > E.h_out_cap) {
            CUDA_OK(cudaStreamSynchronize(st));
            if (E.last_done) CUDA_OK(cudaEventSynchronize(E.last_done));
            cudaFree(E.h_in);
            cudaFree(E.h_out);
            E.h_in = E.h_out = nullptr;
            E.h_in_cap = E.h_out_cap = 0;
            CUDA_OK(cudaMalloc(&E.h_in, clip_in * sizeof(float)));
            CUDA_OK(cudaMalloc(&E.h_out, clip_out * sizeof(float)));
            E.h_in_cap = clip_in;
            E.h_out_cap = clip_out;
        }
        const int Bh = (int)std::max<size_t>(1, std::min(E.h_in_cap / clip_in, E.h_out_cap / clip_out));
        if (out_mag && Bh < B) fail(NUNET_EINVAL, "nunet_forward_wav_host: out_mag needs the whole call to fit the staging buffers (B <= %d here)", Bh);
        for (int b0 = 0; b0 < B; b0 += Bh) {
            const int Bc = std::min(Bh, B - b0);
            E.order_begin(st);
            CUDA_OK(cudaMemcpyAsync(E.h_in, wav + (size_t)b0 * clip_in, (size_t)Bc * clip_in * sizeof(float), cudaMemcpyHostToDevice, st));
            const int before = E.launches;
            E.forward_wav(E.h_in, Bc, n_samples, out_wav ? E.h_out : nullptr, nullptr, st);
            if (b0 > 0) E.launches += before;
            if (out_wav)
                CUDA_OK(cudaMemcpyAsync(out_wav + (size_t)b0 * clip_out, E.h_out, (size_t)Bc * clip_out * sizeof(float), cudaMemcpyDeviceToHost, st));
            if (out_mag) {
                const Engine::Cut c = E.cut_call(B, T);
                if (c.Bs < B || c.Tc < T) fail(NUNET_EINVAL, "nunet_forward_wav_host: out_mag needs the call to fit max_frames in one piece");
                CUDA_OK(cudaMemcpyAsync(out_mag, E.offline.cur(E.o_est, 0), (size_t)B * T * NBINS * sizeof(float), cudaMemcpyDeviceToHost, st));
            }
        }
        CUDA_OK(cudaStreamSynchronize(st));
    });
}

int nunet_forward_mag_dev(nunet_engine* h, const float* mag, int B, int T, float* out_mag, void* stream) {
    return guarded_h(h, [&] {
        if (!h || !mag || !out_mag) fail(NUNET_EINVAL, "null argument");
        h->e.launches = 0;
        h->e.order_begin(static_cast<cudaStream_t>(stream));
        h->e.prof_begin(static_cast<cudaStream_t>(stream));
        h->e.forward_mag(mag, B, T, out_mag, static_cast<cudaStream_t>(stream));
        h->e.order_end(static_cast<cudaStream_t>(stream));
    });
}

int nunet_stream_reset(nunet_engine* h, int first, int count, void* stream) {
    return guarded_h(h, [&] {
        if (!h) fail(NUNET_EINVAL, "null argument");
        h->e.stream_reset(first, count, static_cast<cudaStream_t>(stream));
    });
}

int nunet_stream_step_mag_dev(nunet_engine* h, const float* mag, int S, float* out_mag, void* stream) {
    return guarded_h(h, [&] {
        if (!h || !mag || !out_mag) fail(NUNET_EINVAL, "null argument");
        h->e.launches = 0;
        h->e.order_begin(static_cast<cudaStream_t>(stream));
        h->e.prof_begin(static_cast<cudaStream_t>(stream));
        h->e.stream_step_mag_api(mag, S, out_mag, static_cast<cudaStream_t>(stream));
        h->e.order_end(static_cast<cudaStream_t>(stream));
    });
}

int nunet_stream_step_wav_dev(nunet_engine* h, const float* hop, int S, float* out_hop, float* out_mag, void* stream) {
    return guarded_h(h, [&] {
        if (!h || !hop || !out_hop) fail(NUNET_EINVAL, "null argument");
        h->e.stream_step_wav(hop, S, out_hop, out_mag, static_cast<cudaStream_t>(stream));
    });
}

int nunet_stream_step_wav_host(nunet_engine* h, const float* hop, int S, float* out_hop) {
    return guarded_h(h, [&] {
        if (!h || !hop || !out_hop) fail(NUNET_EINVAL, "null argument");
        Engine& E = h->e;
        E.check_streams(S);
        cudaStream_t st = E.own_stream;
        const size_t n = (size_t)S * HOP;
        E.order_begin(st);
        CUDA_OK(cudaMemcpyAsync(E.h_in, hop, n * sizeof(float), cudaMemcpyHostToDevice, st));
        E.stream_step_wav(E.h_in, S, E.h_out, nullptr, st);
        CUDA_OK(cudaMemcpyAsync(out_hop, E.h_out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    });
}

int nunet_state_count(nunet_engine* h) { return h ? (int)h->e.stream.states.size() : NUNET_EINVAL; }

int nunet_state_name(nunet_engine* h, int index, char* name_out, int cap) {
    return guarded_h(h, [&] {
        if (!h || !name_out || cap <= 0) fail(NUNET_EINVAL, "null argument");
        if (index < 0 || index >= (int)h->e.stream.states.size()) fail(NUNET_EINVAL, "state index out of range");
        snprintf(name_out, (size_t)cap, "%s", h->e.stream.states[index].name.c_str());
    });
}

int nunet_state_numel(nunet_engine* h, const char* name) {
    int n = 0;
    int rc = guarded_h(h, [&] {
        if (!h || !name) fail(NUNET_EINVAL, "null argument");
        const Ten* xt = nullptr;
        int xring = 0;
        n = h->e.extra_state(name, &xt, &xring);
        if (n > 0) return;
        const Plan::StateRef* s = h->e.find_state(name);
        if (!s) fail(NUNET_ESTATE, "unknown state tensor '%s'", name);
        n = h->e.state_numel(*s);
    });
    return rc == NUNET_OK ? n : rc;
}

int nunet_state_export(nunet_engine* h, int stream_id, const char* name, float* buf) {
    return guarded_h(h, [&] {
        if (!h || !name || !buf) fail(NUNET_EINVAL, "null argument");
        h->e.state_xfer(stream_id, name, buf, true);
    });
}

int nunet_state_import(nunet_engine* h, int stream_id, const char* name, const float* buf) {
    return guarded_h(h, [&] {
        if (!h || !name || !buf) fail(NUNET_EINVAL, "null argument");
        h->e.state_xfer(stream_id, name, const_cast<float*>(buf), false);
    });
}

long long nunet_state_generation(nunet_engine* h) { return h ? h->e.state_gen : (long long)NUNET_EINVAL; }

int nunet_last_launch_count(nunet_engine* h) { return h ? h->e.launches : NUNET_EINVAL; }

int nunet_profile_enable(nunet_engine* h, int on) {
    if (!h) return NUNET_EINVAL;
    h->e.prof_on = on != 0;
    return NUNET_OK;
}

int nunet_profile_count(nunet_engine* h) { return h ? (int)h->e.prof.size() : NUNET_EINVAL; }

int nunet_profile_entry(nunet_engine* h, int index, char* name_out, int cap, float* ms_out, double* alg_bytes_out) {
    return guarded_h(h, [&] {
        if (!h || !name_out || !ms_out || !alg_bytes_out) fail(NUNET_EINVAL, "null argument");
        Engine& E = h->e;
        if (index < 0 || index >= (int)E.prof.size()) fail(NUNET_EINVAL, "profile index out of range");
        CUDA_OK(cudaEventSynchronize(E.prof[index].ev));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, index == 0 ? E.prof_start : E.prof[index - 1].ev, E.prof[index].ev));
        *ms_out = ms;
        *alg_bytes_out = E.prof[index].alg_bytes;
        snprintf(name_out, (size_t)cap, "%s", E.prof[index].name.c_str());
    });
}

long long nunet_debug_read(nunet_engine* h, const char* tensor_name, float* buf, long long cap) {
    long long n = 0;
    int rc = guarded_h(h, [&] {
        if (!h || !tensor_name) fail(NUNET_EINVAL, "null argument");
        Engine& E = h->e;
        auto it = E.offline.named.find(tensor_name);
        if (it == E.offline.named.end()) fail(NUNET_ESTATE, "unknown tensor '%s'", tensor_name);
        const Ten* t = it->second;
        n = (long long)E.last_B * E.last_T * (long long)t->numel();
        if (buf) {
            if (cap < n) fail(NUNET_EINVAL, "buffer too small");
            CUDA_OK(cudaDeviceSynchronize());
            CUDA_OK(cudaMemcpy(buf, E.offline.cur(t, 0), (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
            if (t->sh) {   // planar frame rows [hi|lo][chunk][f][8 halves] -> floats [f][c], in place
                const int C = t->C, F = t->F;
                std::vector<__half> row((size_t)2 * F * C);
                for (long long fr = 0; fr < n / ((long long)F * C); ++fr) {
                    float* dst = buf + fr * (long long)F * C;
                    memcpy(row.data(), dst, (size_t)4 * F * C);
                    for (int f = 0; f < F; ++f)
                        for (int c = 0; c < C; ++c) {
                            const int pos = t->eo ? (f & 1) * (F >> 1) + (f >> 1) : f;
                            const size_t hi = ((size_t)(c >> 3) * F + pos) * 8 + (c & 7);
                            const size_t lo = ((size_t)((C >> 3) + (c >> 3)) * F + pos) * 8 + (c & 7);
                            dst[(size_t)f * C + c] = __half2float(row[hi]) + __half2float(row[lo]);
                        }
                }
            }
        }
    });
    return rc == NUNET_OK ? n : rc;
}

}  // extern "C"
