/* gtcrn_b200.h -- C ABI of the B200-native GTCRN-Micro inference path (libgtcrn_b200.so).
 *
 * The reference (bglid/GTCRN-Micro) has no FFI: its hot path is the Python nn.Module API
 *   GTCRNMicro.forward            gtcrn_micro/models/gtcrn_micro.py:506-532
 *   StreamGTCRNMicro.forward      gtcrn_micro/streaming/gtcrn_micro_stream.py:541-574
 *   torch.stft / torch.istft      gtcrn_micro/infer.py:60-76
 * executed as ~60 ATen ops per forward.  This library replaces everything below that API with
 * hand-written sm_100a kernels; `gtcrn_micro_b200/modules.py` keeps the nn.Module surface and calls
 * these entry points through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer, *_host a host pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     asynchronous on that stream unless stated otherwise; a handle may be used from one stream
 *     at a time (its activation workspace is shared);
 *   - return value 0 = success, otherwise one of GTCRNB_E_*; gtcrnb_last_error() gives a message
 *     (thread-local).  No exceptions cross the ABI;
 *   - all audio / spectra / weights are IEEE fp32.  16 kHz, n_fft 512, hop 256, sqrt-Hann.
 */
#ifndef GTCRN_B200_H_
#define GTCRN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTCRNB_OK 0
#define GTCRNB_E_INVALID 1   /* bad argument (NULL pointer, bad shape, L <= 256, ...)            */
#define GTCRNB_E_BLOB 2      /* weight blob has the wrong size / version                         */
#define GTCRNB_E_CUDA 3      /* a CUDA runtime call or kernel launch failed                      */
#define GTCRNB_E_NOMEM 4     /* workspace allocation failed                                      */
#define GTCRNB_E_ARCH 5      /* device is not compute capability 10.x (kernels are sm_100a only) */

typedef struct gtcrnb_handle gtcrnb_handle;

/* Library / blob versioning. */
int gtcrnb_abi_version(void);
/* Number of fp32 values `gtcrnb_create` expects (layout: gtcrn_micro_b200/pack.py). */
size_t gtcrnb_blob_floats(void);

/* Create a model instance on the current CUDA device from a packed weight blob (host pointer).
 * Replaces: GTCRNMicro().load_state_dict(ckpt["model"]).eval()   (infer.py:37-40). */
int gtcrnb_create(const float* blob_host, size_t n_floats, gtcrnb_handle** out);
/* Same with an explicit model variant (SURVEY 8(b) proposal):
 *   GTCRNB_VARIANT_MICRO  the reference's GTCRNMicro (SFE-Lite, TRALite, GTCN) -- what gtcrnb_create builds;
 *   GTCRNB_VARIANT_FULL   the original-GTCRN-style architecture of BASELINE configs[3] (unfold SFE, groups = 2, GTConv dilations
 *                         1/2/5, GRU-based TRA, DPGRNN), which the reference only keeps as commented-out lines
 *                         (gtcrn_micro/models/gtcrn_micro.py:187,199,346,361,381,391,417-418,429,448,494): parity UNPINNED, checked
 *                         against the builder-written oracle/gtcrn_full_oracle.py; blob layout gtcrn_micro_b200/pack_full.py.
 * Every other entry point takes either kind of handle; gtcrnb_stream_import/export_caches are Micro-only. */
#define GTCRNB_VARIANT_MICRO 0
#define GTCRNB_VARIANT_FULL 1
int gtcrnb_create_variant(const float* blob_host, size_t n_floats, int variant, gtcrnb_handle** out);
size_t gtcrnb_blob_floats_variant(int variant);
void gtcrnb_destroy(gtcrnb_handle* h);

/* Drop-in for GTCRNMicro.forward (gtcrn_micro.py:506-532): spec [B,257,T,2] -> enhanced spec.
 * Both tensors are addressed with element strides {b, f, t, c}, so torch.stft's non-contiguous
 * output can be passed as is.  Zero temporal history (start of utterance). */
int gtcrnb_forward_spec(gtcrnb_handle* h, const float* spec_dev, const int64_t in_strides[4],
                        int64_t B, int64_t T, float* out_dev, const int64_t out_strides[4], void* stream);

/* Fused replacement for infer.py:60-76: wav [B,L] -> STFT -> model -> iSTFT -> wav [B, 256*(L/256)].
 * L > 256.  Rows are contiguous with the given leading strides (in elements). */
int gtcrnb_enhance(gtcrnb_handle* h, const float* wav_dev, int64_t B, int64_t L, int64_t in_stride,
                   float* out_dev, int64_t out_stride, void* stream);

/* Same computation on HOST buffers (ideally pinned): chunked H2D -> kernels -> D2H pipeline on
 * internal streams; returns after the last chunk has landed in out_host (synchronous). */
int gtcrnb_enhance_host(gtcrnb_handle* h, const float* wav_host, int64_t B, int64_t L,
                        float* out_host);

/* The same pipeline with PCM16 samples on the wire (wav files are PCM16; half the PCIe bytes): the device decodes
 * x / 32768 -- what the reference's sf.read(..., dtype="float32") yields, infer.py:52 -- and encodes
 * floor(y * 32768) clamped to int16, the quantisation of the reference's sf.write(enh_path, enhanced, fs) outputs
 * (infer.py:107; the shipped *_enh.wav examples).  out_host [B, 256*(L/256)] int16. */
int gtcrnb_enhance_host_pcm16(gtcrnb_handle* h, const int16_t* wav_host, int64_t B, int64_t L,
                              int16_t* out_host);

/* STFT / iSTFT on their own (torch.stft / torch.istft call sites infer.py:60-67, :73-76).
 * spec is contiguous [B][T][257][2], T = 1 + L/256; istft writes [B][256*(T-1)]. */
int gtcrnb_stft(gtcrnb_handle* h, const float* wav_dev, int64_t B, int64_t L, int64_t in_stride,
                float* spec_dev, void* stream);
int gtcrnb_istft(gtcrnb_handle* h, const float* spec_dev, int64_t B, int64_t T, float* wav_dev,
                 int64_t out_stride, void* stream);

/* ---- streaming (StreamGTCRNMicro.forward, gtcrn_micro_stream.py:541-574) ------------------------
 * The state of N concurrent streams (conv / TRA / TCN caches, plus STFT/OLA tails for the wav
 * entry point) lives in one device buffer owned by the caller, `gtcrnb_stream_state_bytes(N)` big. */
size_t gtcrnb_stream_state_bytes(int64_t N);                       /* GTCRNB_VARIANT_MICRO */
size_t gtcrnb_stream_state_bytes_variant(int variant, int64_t N);
int gtcrnb_stream_reset(gtcrnb_handle* h, void* state_dev, int64_t N, void* stream);
/* One 16 ms hop for N streams on spectra: spec_frame [N][257][2] -> out [N][257][2] (contiguous). */
int gtcrnb_stream_step_spec(gtcrnb_handle* h, void* state_dev, const float* spec_dev, float* out_dev,
                            int64_t N, void* stream);
/* One hop on waveforms: hop_in [N][256] new samples -> hop_out [N][256] enhanced samples.  The
 * output lags the input by 512 samples (32 ms: one frame of look-ahead for the centred STFT plus
 * one hop of overlap-add); the first two hops are zeros; see DESIGN.md. */
int gtcrnb_stream_step_wav(gtcrnb_handle* h, void* state_dev, const float* hop_in_dev, float* hop_out_dev,
                           int64_t N, void* stream);
/* End of stream for the wav entry point: after K calls of gtcrnb_stream_step_wav (K >= 2 full hops) plus `tail_len` (0..255)
 * trailing samples tail_dev [N][tail_len] (NULL when 0), emit the two hops the 32 ms latency still holds back:
 * out_dev [2][N][256] = offline output hops K-2 and K-1 of a signal of L = 256 K + tail_len samples, including torch.stft's
 * reflect padding at the end of the signal (infer.py:60-67) -- with it the streamed output equals gtcrnb_enhance's
 * 256*(L/256) samples exactly.  The state is finished afterwards (gtcrnb_stream_reset before reuse). */
int gtcrnb_stream_flush_wav(gtcrnb_handle* h, void* state_dev, const float* tail_dev, int64_t tail_len,
                            float* out_dev, int64_t N, void* stream);
/* Import / export the model caches of a stream batch in the reference's own tensor layout
 * (gtcrn_micro_stream.py:618-623): conv_cache [2,N,16,6,33], tra_cache [2,3,N,8,2],
 * tcn_cache[g][k] [N,16,2*d_k,33] passed as 8 pointers (g-major).  All device pointers. */
int gtcrnb_stream_import_caches(gtcrnb_handle* h, void* state_dev, int64_t N, const float* conv_cache_dev,
                                const float* tra_cache_dev, const float* const tcn_cache_dev[8], void* stream);
int gtcrnb_stream_export_caches(gtcrnb_handle* h, const void* state_dev, int64_t N, float* conv_cache_dev,
                                float* tra_cache_dev, float* const tcn_cache_dev[8], void* stream);

/* PCM16 <-> float32 on device buffers of n samples (the conversions gtcrnb_enhance_host_pcm16 applies): decode x / 32768,
 * encode floor(y * 32768) clamped to [-32768, 32767]. */
int gtcrnb_pcm16_to_f32(gtcrnb_handle* h, const int16_t* in_dev, float* out_dev, int64_t n, void* stream);
int gtcrnb_f32_to_pcm16(gtcrnb_handle* h, const float* in_dev, int16_t* out_dev, int64_t n, void* stream);

/* ---- on-device quality metrics ---------------------------------------------------
 * SI-SNR and SDR of B (reference, enhanced) waveform pairs of L samples, as defined by the reference's evaluation script
 * (gtcrn_micro/eval/eval_intrusive_metrics.py:74-82 sisnr_metric, :85-94 sdr_metric).  out_dev [B][2] = (si_snr_db, sdr_db). */
int gtcrnb_pair_metrics(gtcrnb_handle* h, const float* ref_dev, int64_t ref_stride, const float* inf_dev, int64_t inf_stride,
                        int64_t B, int64_t L, float* out_dev, void* stream);

/* ---- introspection -------------------------------------------------------------------------- */
/* Copy an intermediate activation of the LAST gtcrnb_forward_spec / gtcrnb_enhance call (first
 * workspace chunk) to a device buffer, channel-last [B][T][F][16].  name: "en0" (F=65), "en1".."en4",
 * "gtcn1", "gtcn2", "de0".."de2" (F=33).  Used by the parity tests to localise an error. */
int gtcrnb_debug_activation(gtcrnb_handle* h, const char* name, float* out_dev, size_t n_floats, void* stream);
/* Per-kernel-class device timing: when enabled every launch is bracketed by CUDA events on the launching
 * stream.  gtcrnb_profile_read waits for the recorded events and returns, per class, accumulated
 * milliseconds, launch count and frames processed (arrays of gtcrnb_profile_num_classes() entries). */
int gtcrnb_profile_enable(gtcrnb_handle* h, int on);
int gtcrnb_profile_read(gtcrnb_handle* h, int reset, double* ms_out, int64_t* launches_out, int64_t* frames_out);
int gtcrnb_profile_num_classes(void);
const char* gtcrnb_profile_class_name(int i);
/* Which layers of a GTCRNB_VARIANT_MICRO handle run on the tensor cores (tcgen05, fp32-class accuracy through hi/lo operand splits):
 * bit 0 the GTCN pointwise convolutions, bit 1 the decoder GTConv 3x3 convolution, bit 2 en1 inside the front kernel, bit 3 de3 inside
 * the back kernel.  Default 15; the environment variable GTCRNB_TC (read at gtcrnb_create) overrides it, 0 = FFMA kernels only, whose offline results are bit-identical to the streaming path. */
int gtcrnb_tc_mask(const gtcrnb_handle* h);
/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t gtcrnb_launch_count(const gtcrnb_handle* h);
/* Pinned host memory helpers for the *_host entry points. */
void* gtcrnb_host_alloc(size_t bytes);
/* write-combined page-locked memory for INPUT buffers the CPU only writes sequentially (never reads back): the DMA engine
 * reads it without snooping the CPU caches.  Free with gtcrnb_host_free. */
void* gtcrnb_host_alloc_wc(size_t bytes);
void gtcrnb_host_free(void* p);
const char* gtcrnb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GTCRN_B200_H_ */
