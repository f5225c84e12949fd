/* nunet_b200.h -- C ABI of the B200-native NUNet-TLS / NUNet-TLS-LSTM inference path.
 *
 * The reference has no FFI of its own: its hot path is reached through two Python call surfaces
 * (all paths relative to the reference checkout, dnn_model/):
 *   surface 1 (offline)   models/proposed.py:627  NUTLS_LSTM(opt).build_model() -> model(wav[B,N]) -> wav
 *                         (graph: models/proposed.py:284-625; same pair in models/nunet_tls.py:361-1017)
 *   surface 2 (streaming) interpreter_proposed.py:374-380  Interpreter(...).get_signature_runner('nutls_lstm_sm')
 *                         called once per hop with `input` + 130 history tensors (interpreter_proposed.py:215-350),
 *                         graph spec converter_proposed.py:188-867; frame loop interpreter_proposed.py:15-370.
 * Each entry point below names the reference interface it replaces.  The Python host
 * (nunet_b200/models.py, nunet_b200/interpreter.py) mirrors those two surfaces on top of this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / Python types.
 *   - `*_dev` functions take DEVICE pointers owned by the caller and are asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - `*_host` functions take HOST pointers, do the H2D / D2H copies themselves on the handle's own
 *     stream and return after the result is in the host buffer (the end-to-end call a user makes).
 *   - every function returns 0 on success or a negative NUNET_E* code; nunet_last_error() gives the text.
 *   - no allocation after nunet_create(): capacity is fixed by max_frames / max_streams.
 *   - all tensors are float32.  Spectrogram layout is the reference's [B, T, F] (F fastest).
 */
#ifndef NUNET_B200_H
#define NUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define NUNET_ABI_VERSION 1

enum {
    NUNET_OK = 0,
    NUNET_EINVAL = -1,   /* bad argument / shape / blob */
    NUNET_ENOMEM = -2,   /* capacity exceeded or cudaMalloc failed */
    NUNET_ECUDA = -3,    /* CUDA runtime error (text in nunet_last_error) */
    NUNET_ENODEV = -4,   /* no sm_100 device */
    NUNET_ESTATE = -5    /* unknown state tensor name */
};

/* variant: which bottleneck block the nested U-Net uses */
enum { NUNET_VARIANT_LSTM = 0,   /* models/proposed.py   NUTLS_LSTM  */
       NUNET_VARIANT_DDB = 1,    /* models/nunet_tls.py  NUTLS (dilated dense block) */
       NUNET_VARIANT_LSTM_HYBRID = 2 };
       /* NUTLS_LSTM with the arithmetic of the DEPLOYED artefact: the dynamic-range-quantised one-frame .tflite
          (converter_proposed.py:901 Optimize.DEFAULT) as the TFLite runtime executes it -- int8 weights, every hybrid
          CONV_2D / FULLY_CONNECTED input quantised to int8 per call, int32 accumulation (interpreter_proposed.py:374-380,
          RTSE_NUTLS_LSTM.java:571).  Streaming entry points only (the reference has no offline int8 graph); the blob carries
          the int8 tensors and their scales (nunet_b200.weights.hybrid_weight_set). */

/* CTFA time pooling */
enum { NUNET_CTFA_CAUSAL_AVG32 = 0,  /* offline graph: models/proposed.py:125 `ctfa` (mean of last 32 TA) */
       NUNET_CTFA_FRAME_DIV32 = 1 }; /* one-frame graph: models/proposed.py:162 `ctfa_rt` (TA/32)         */

/* DC-bin restore before the inverse FFT */
enum { NUNET_DC_ZERO = 0,   /* models/proposed.py:617 tf.pad; RTSE_NUTLS_LSTM.java:677 */
       NUNET_DC_EDGE = 1 }; /* interpreter_proposed.py:352 np.pad(mode='edge')         */

typedef struct nunet_engine nunet_engine;

typedef struct nunet_config {
    int32_t variant;       /* NUNET_VARIANT_* (must match the blob) */
    int32_t device;        /* CUDA device ordinal */
    int32_t max_frames;    /* offline capacity: frames resident at once (sizes the activation arena; 0 = offline disabled).
                              Calls with more than max_frames frames are processed in sub-batches of whole clips and, for
                              clips longer than max_frames, in time chunks (LSTM variant). */
    int32_t max_streams;   /* streaming capacity: concurrent streams (0 = streaming disabled) */
    int32_t ctfa_mode;     /* NUNET_CTFA_* used by the OFFLINE path; streaming is always frame_div32
                              unless stream_ctfa_history != 0 */
    int32_t dc_mode;       /* NUNET_DC_* used by the streaming wav path (offline always pads zero) */
    int32_t stream_ctfa_history; /* extension: carry 31 frames of TA per stream so that streaming == offline */
    int32_t chunk_frames;  /* offline: process every clip in time chunks of at most this many frames, carrying the conv history
                              rows, LSTM h / c and the 31-frame attention window from chunk to chunk (the reference bounds its
                              memory the same way: options.py:42 chunk_size, and the one-frame graph converter_proposed.py:
                              188-867 is the chunk = 1 limit).  0 = whole clips; a clip longer than max_frames is still cut into
                              chunks of max_frames.  Chunked and unchunked results are bit-identical for chunks of two or more frames
                              (one-frame chunks use the streaming kernel variants: equal to fp32 rounding). */
} nunet_config;

/* Text of the last error on the calling thread ("" if none). */
const char* nunet_last_error(void);
int nunet_abi_version(void);

/* Host-only check of a weight blob (no GPU needed): parses it, verifies every tensor the variant needs is present
 * with the reference shape and packs the parameters.  Returns the packed float count or a negative error. */
long long nunet_blob_validate(const void* blob, size_t blob_bytes, int variant);

/* Replaces: NUTLS_LSTM(opt) + model.load_weights(path) (test_interface.py:45, converter_proposed.py:13) and
 * tf.lite.Interpreter(model_path) + allocate_tensors() (interpreter_proposed.py:374-375).
 * `blob` is the packed role-named weight set produced by nunet_b200.weights.pack_blob (HOST memory). */
int nunet_create(const nunet_config* cfg, const void* blob, size_t blob_bytes, nunet_engine** out);
void nunet_destroy(nunet_engine* h);

/* Number of STFT frames tf.signal.stft(frame_length=512, frame_step=256, pad_end=False) yields for n samples
 * (models/proposed.py:285): 1 + (n-512)/256, or 0 if n < 512. */
int nunet_num_frames(int n_samples);

/* Replaces model(noisy_wav, training=False) (test_interface.py:58; graph models/proposed.py:284-625).
 *   wav      [B, n_samples]
 *   out_wav  [B, (T-1)*256+512]           (nullable)
 *   out_mag  [B, T, 257] estimated magnitudes, DC bin zero  (nullable)            */
int nunet_forward_wav_dev(nunet_engine* h, const float* wav, int B, int n_samples,
                          float* out_wav, float* out_mag, void* stream);
int nunet_forward_wav_host(nunet_engine* h, const float* wav, int B, int n_samples,
                           float* out_wav, float* out_mag);

/* The network alone (models/proposed.py:293-615): mag [B,T,256] (DC already dropped) -> est [B,T,256]. */
int nunet_forward_mag_dev(nunet_engine* h, const float* mag, int B, int T, float* out_mag, void* stream);

/* ---- streaming (converter_proposed.py:188-867 one-frame graph; interpreter_proposed.py:200-366 loop) ---- */

/* Zero the history of streams [first, first+count): the zero dict of interpreter_proposed.py:36-198. */
int nunet_stream_reset(nunet_engine* h, int first, int count, void* stream);

/* One signature call for S streams at once: mag [S,256] -> est [S,256]; history advances in place.
 * Replaces nutls_lstm_sm(input=..., *_prevK=..., *_h/_c=...) (interpreter_proposed.py:215-350). */
int nunet_stream_step_mag_dev(nunet_engine* h, const float* mag, int S, float* out_mag, void* stream);

/* One iteration of the frame loop of real_time_speech_enhancer (interpreter_proposed.py:200-366) for S
 * streams: hop [S,256] new samples -> out_hop [S,256] enhanced samples (16 ms algorithmic latency).
 * out_mag [S,256] (nullable) receives the enhanced magnitudes of this frame. */
int nunet_stream_step_wav_dev(nunet_engine* h, const float* hop, int S, float* out_hop, float* out_mag,
                              void* stream);
int nunet_stream_step_wav_host(nunet_engine* h, const float* hop, int S, float* out_hop);

/* History wire format: the tensors of the reference signature (converter_proposed.py:26-187 inputs,
 * :729-867 outputs), addressed by their reference names without the _prev/_cur infix, e.g.
 * "msfe6_ee_1" <-> msfe6_ee_prev1/msfe6_ee_cur1, "msfe6_en_h", "state_c".
 * buf is a HOST buffer of nunet_state_numel(name) floats.
 * Not enumerated by nunet_state_count/name but accepted by numel/export/import, for stream checkpoints: the frame
 * loop's "in_buffer" and "out_buffer" (512 floats each, interpreter_proposed.py:30-31) and, with
 * stream_ctfa_history, the attention rings "ctfa_ring<i>" (32 x 64 floats, oldest frame first). */
int nunet_state_count(nunet_engine* h);
int nunet_state_name(nunet_engine* h, int index, char* name_out, int cap);
int nunet_state_numel(nunet_engine* h, const char* name);
int nunet_state_export(nunet_engine* h, int stream_id, const char* name, float* buf);
int nunet_state_import(nunet_engine* h, int stream_id, const char* name, const float* buf);
/* Counter that changes whenever the resident history changes (a step, a reset, an import).  The signature runner
 * (interpreter.py) uses it to decide whether the arrays a caller feeds back (interpreter_proposed.py:215 replaces the
 * whole dict every hop) are still what the engine holds, or must be imported. */
long long nunet_state_generation(nunet_engine* h);

/* Introspection used by bench.py / tests: kernels launched by the most recent forward/step call. */
int nunet_last_launch_count(nunet_engine* h);
/* Per-launch profile of the NEXT forward/step calls (bench.py roofline leg): when enabled, a CUDA event is
 * recorded on the launching stream after every kernel.  Entry i = (kernel name, device milliseconds since the
 * previous event, algorithmic bytes = what that launch must read + write once). */
int nunet_profile_enable(nunet_engine* h, int on);
int nunet_profile_count(nunet_engine* h);
int nunet_profile_entry(nunet_engine* h, int index, char* name_out, int cap, float* ms_out, double* alg_bytes_out);

/* Debug tap: copy an intermediate tensor of the most recent OFFLINE forward to host (tests only).
 * Returns the element count (per call) or a negative error; buf may be NULL to query the size. */
long long nunet_debug_read(nunet_engine* h, const char* tensor_name, float* buf, long long cap);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* NUNET_B200_H */
