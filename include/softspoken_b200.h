/*
 * softspoken_b200 — C ABI of the B200-native voice-detector batch path.
 *
 * The reference (AVianEco/Softspoken) has no FFI: its hot path is duck-typed
 * Python.  Each entry point below replaces one reference
 * function or group of functions, cited as file:line relative to the reference
 * tree; INTEGRATION.md shows the ctypes stub a maintainer adds on the
 * reference side.  Conventions:
 *
 *   - extern "C", plain pointers and sizes, no C++/torch types;
 *   - every function returns 0 on success or a negative SS_E_* code and never
 *     throws; ss_last_error() gives the message of the calling thread's last
 *     failure;
 *   - "_dev" pointers are device pointers on the context's GPU, owned by the
 *     caller; "_host" pointers are host memory (pinned or pageable);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream); device entry points only enqueue work, host entry points
 *     synchronise before returning;
 *   - a context allocates device memory only in ss_ctx_create (weights, classifier
 *     workspace) and ss_ctx_reserve (file-level scratch); the compute entry points
 *     never allocate; a context is used by one host thread at a time;
 *   - there is no CPU fallback: without a CUDA device ss_ctx_create fails.
 */
#ifndef SOFTSPOKEN_B200_H
#define SOFTSPOKEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_ABI_VERSION 2

#if defined(__GNUC__)
#define SS_API __attribute__((visibility("default")))
#else
#define SS_API
#endif

/* status codes */
#define SS_OK 0
#define SS_E_ARG (-1)      /* bad argument */
#define SS_E_CUDA (-2)     /* CUDA runtime error (message in ss_last_error) */
#define SS_E_BLOB (-3)     /* malformed / incompatible weight blob */
#define SS_E_CAPACITY (-4) /* output capacity or context limit exceeded */
#define SS_E_NODEVICE (-5) /* no usable CUDA device */
#define SS_E_RANGE (-6)    /* an activation left the fp16 range in an fp16-operand mode: result invalid */

/* classifier arithmetic (ss_classify `mode`) */
#define SS_MODE_FP32 0  /* CUDA-core float32 direct convolution (reference arithmetic, slow) */
#define SS_MODE_BF16 1  /* tcgen05 / TMEM implicit GEMM, bf16 operands, fp32 accumulate: throughput mode */
#define SS_MODE_F16 2   /* same kernel, fp16 operands (weights pre-scaled per layer): 8x finer than bf16 */
#define SS_MODE_F16X3 3 /* same kernel, fp16 hi/lo split operands, 3 MMAs per product: fp32-grade logits on
                           tensor cores; the default (parity) mode of the drop-in path */

typedef struct ss_ctx ss_ctx;

/* Half-open range [begin, end) of float32 elements inside one device buffer. */
typedef struct ss_interval {
  int64_t begin;
  int64_t end;
} ss_interval;

SS_API int ss_abi_version(void);
SS_API const char* ss_last_error(void);

/* Compile-time constant of the kernels by name ("sample_rate", "window_samples",
 * "step_samples", "pad_samples", "n_frames", "n_mels", "gap_bins", ...), so the host side can
 * assert it matches root/code/backend/settings.py:4-16 and NNDetector.py:67-75. */
SS_API int ss_get_constant(const char* name, double* value);

SS_API int ss_device_count(int* count);
/* Number of CUDA kernels this library has launched in the process so far (bench.py `gpu_launches`). */
SS_API int ss_launch_count(uint64_t* count);

/* Replaces NNDetector.__init__ model construction + load_checkpoint
 * (root/code/frontend/NNDetector.py:21-53): `blob` is the packed, BN-folded state dict
 * (softspoken_b200/checkpoint.py:pack_blob).  `max_batch_windows` bounds the windows one
 * classifier pass holds in its workspace (larger calls are split internally). */
SS_API int ss_ctx_create(int device, const void* blob, size_t blob_bytes, int max_batch_windows, ss_ctx** ctx);
SS_API int ss_ctx_destroy(ss_ctx* ctx);
SS_API int ss_ctx_device_bytes(ss_ctx* ctx, size_t* bytes);
/* Size the file-level scratch of ss_detect_* for clips of up to `max_samples` samples and up to
 * `region_cap` regions per clip.  Grows only; a detect call beyond the reservation fails with
 * SS_E_CAPACITY instead of allocating. */
SS_API int ss_ctx_reserve(ss_ctx* ctx, int64_t max_samples, int region_cap);

/* Margin-guided refinement of the file-level calls (ss_detect_*).  A detection is the comparison `avg > 0.1`
 * (NNDetector.py:120) on an average of up to five window logits; the tensor-core classifier's logits carry more
 * rounding noise than the reference's float32 (operand precisions), so a bin whose average falls inside the
 * noise band could land on the other side of the threshold.  With eps > 0 every window covering a bin with
 * |avg - 0.1| < eps is classified again in `refine_mode` (SS_MODE_FP32: the reference's own arithmetic class) and
 * K5 / K6 run on the patched logits.  The calls then synchronise their stream once per clip (the host needs the
 * number of marked windows).  eps = 0 switches the refinement off; a call whose `mode` equals refine_mode skips it.
 * Default: see DESIGN.md.  ss_ctx_refine_stats: {windows classified, windows refined, clips, clips with at least one
 * refined window} since creation or the last reset. */
SS_API int ss_ctx_set_refine(ss_ctx* ctx, double eps, int refine_mode);
SS_API int ss_ctx_refine_stats(ss_ctx* ctx, uint64_t* stats4, int reset);

/* NNDetector.plan_detection_job (NNDetector.py:65-80): number of 3 s windows of a clip of
 * `n_samples` samples at 22,050 Hz once padded (worker.py:58-62); starts are i * 13230. */
SS_API int64_t ss_plan_windows(int64_t n_samples);
/* NNDetector.average_overlapping_detections `output_length` (NNDetector.py:168) for a padded length. */
SS_API int64_t ss_timeline_bins(int64_t n_padded);

/* worker.py:58-62 on the device: dst[0:pad)=0, dst[pad:pad+n)=src, dst[pad+n:n+2pad)=0. */
SS_API int ss_pad(ss_ctx* ctx, const float* pcm_dev, int64_t n_samples, float* padded_dev, void* stream);

/* K1 — window gather + MelSpectrogram + sqrt(log10(x+1)) + [:, :, :256]
 * (NNDetector.py:90-96; pytorch_neural_nets.py:92-99,144-153).
 * pcm_dev: padded clip, n_padded floats.  win_start_dev: n_windows int64 sample offsets (each
 * start + 65536 <= n_padded is required; samples beyond 65535 of a window never matter).
 * mel_out_dev: [n_windows][128 mel][256 frames] float32. */
SS_API int ss_features(ss_ctx* ctx, const float* pcm_dev, int64_t n_padded, const int64_t* win_start_dev,
                int n_windows, float* mel_out_dev, void* stream);

/* K9 — sample-rate conversion to the detector's rate: stands where load_audio calls
 * librosa.resample (root/code/backend/voice_activity.py:44-66; soxr "HQ" there — NOT restated here: this is a
 * polyphase FIR with a caller-designed table, parity with the reference unpinned for resampled files).
 *   out[m] = sum_{j=-T..T} pcm[(m * down) div up - j] * table[(j + T) * up + m mod up],  pcm = 0 outside;
 *   column q = m mod up of the table holds the filter phase (q * down) mod up (visit order)
 * n_out must be ceil(n_in * up / down) (librosa's output length); table_dev: [2 * taps_half + 1][up] float32
 * (softspoken_b200/resample.py designs it: Kaiser-windowed sinc, unit DC gain per phase).
 * *_pcm16 reads the int16 samples of a PCM_16 file (value / 32768). */
SS_API int ss_resample(ss_ctx* ctx, const float* pcm_dev, int64_t n_in, float* out_dev, int64_t n_out, int up, int down,
                int taps_half, const float* table_dev, void* stream);
SS_API int ss_resample_pcm16(ss_ctx* ctx, const int16_t* pcm_dev, int64_t n_in, float* out_dev, int64_t n_out, int up,
                      int down, int taps_half, const float* table_dev, void* stream);

/* K8 — review-screen spectrogram: voice_activity.wav_to_spec(data, trim_edges=False)
 * (root/code/backend/voice_activity.py:148-154) = np.abs(librosa.stft(data, n_fft=512, win_length=512,
 * hop_length=256)): periodic Hann, centred frames, zero padding.  mag_dev: [257 bins][ss_spectrogram_frames(n)]
 * float32, frequency-major as librosa returns it.  max_dev: NULL, or one float that receives the largest
 * magnitude (what the dB stage needs).  *_pcm16 reads the int16 samples of a PCM_16 file (value / 32768). */
SS_API int64_t ss_spectrogram_frames(int64_t n_samples);
SS_API int ss_spectrogram(ss_ctx* ctx, const float* pcm_dev, int64_t n_samples, float* mag_dev, float* max_dev,
                   void* stream);
SS_API int ss_spectrogram_pcm16(ss_ctx* ctx, const int16_t* pcm_dev, int64_t n_samples, float* mag_dev, float* max_dev,
                         void* stream);
/* The display transform of ReviewDetectionsScreen.display_spectrogram (root/code/frontend/review_detections.py:880-881),
 * in place on mag_dev: np.abs(librosa.amplitude_to_db(S ** 2, ref=np.max)) in float32 (0 at the loudest cell, 80 at
 * the floor; librosa squares its argument once more, so this is |40 log10(S / max S)| clipped at 80).
 * max_dev: the float ss_spectrogram wrote. */
SS_API int ss_spectrogram_db(ss_ctx* ctx, float* mag_dev, int64_t n_elems, const float* max_dev, void* stream);

/* K2-K4 — SpecUNet_2D.forward after the front end (pytorch_neural_nets.py:156-195).
 * mel_dev: [n_windows][128][256].  logits_dev: [n_windows][256] raw mask logits (no sigmoid).
 * spec_out_dev: NULL, or [n_windows][2][128][256] for the separation head the reference
 * computes and discards (worker.py:78-79). */
SS_API int ss_classify(ss_ctx* ctx, const float* mel_dev, int n_windows, float* logits_dev,
                float* spec_out_dev, int mode, void* stream);

/* K5 — NNDetector.average_overlapping_detections (NNDetector.py:168-186): float64 overlap-add
 * mean in window order.  avg_dev: out_len doubles (NaN where count == 0), count_dev: out_len int32. */
SS_API int ss_average(ss_ctx* ctx, const float* logits_dev, int n_windows, int64_t out_len, double* avg_dev,
               int32_t* count_dev, void* stream);

/* K6 — NNDetector.find_speech_regions (NNDetector.py:109-141) in bin space: avg > threshold runs
 * (end inclusive), merged when next_start - cur_end <= gap_bins.  regions_dev: [cap][2] int32
 * (start_bin, end_bin); n_regions_dev: one int32 = number found (may exceed cap; only cap stored). */
SS_API int ss_regions(ss_ctx* ctx, const double* avg_dev, const int32_t* count_dev, int64_t out_len,
               double threshold, int gap_bins, int32_t* regions_dev, int32_t* n_regions_dev, int cap,
               void* stream);

/* K7 — SilenceWorker.run inner loop `audio[:, s:e] = 0.0` (silencer_ui.py:974-985) for a table of
 * element ranges inside one packed float32 buffer of n_elems elements. */
SS_API int ss_silence(ss_ctx* ctx, float* pcm_dev, int64_t n_elems, const ss_interval* intervals_dev,
               int n_intervals, void* stream);

/* File-level path on device-resident audio: pad -> K1 -> K2/K3 -> K5 -> K6
 * (ProcessWorker.run per-file body, worker.py:57-97).  pcm_dev is the UNPADDED clip.
 * Results stay on the device; regions as in ss_regions.  logits_out_dev may be NULL.  Enqueues only, except that an
 * active refinement (ss_ctx_set_refine) synchronises `stream` once inside the call. */
SS_API int ss_detect_device(ss_ctx* ctx, const float* pcm_dev, int64_t n_samples, int mode,
                     int32_t* regions_dev, int32_t* n_regions_dev, int cap, float* logits_out_dev,
                     void* stream);

/* Same with HOST buffers (the reference-facing call: load_audio output in, region bins out).
 * The clip is streamed to the device in chunks of windows (chunk k+1 uploads on a copy stream while
 * chunk k computes; consecutive chunks overlap by 52,920 samples because window starts stay on the
 * global i * 13230 grid), logits accumulate on the device, K5/K6 run once on the whole timeline and
 * the regions are copied back.  Synchronises.  n_regions receives the number found; at most `cap`
 * pairs are written.  logits_host: NULL or [n_windows][256]. */
SS_API int ss_detect_host(ss_ctx* ctx, const float* pcm_host, int64_t n_samples, int mode,
                   int32_t* regions_host, int cap, int* n_regions, float* logits_host);

/* The file loop of ProcessWorker.run (worker.py:49-136) for `n_clips` host clips in one call: clip k+1 is uploaded
 * while clip k computes, region lists come back through pinned staging, and the host synchronises once per group of
 * 8 clips instead of once per clip.  regions_host: [n_clips][cap][2]; n_regions: [n_clips] (numbers found; at most
 * cap pairs per clip are written).  Results are identical to n_clips calls of ss_detect_host. */
SS_API int ss_detect_host_batch(ss_ctx* ctx, int n_clips, const float* const* pcm_host, const int64_t* n_samples,
                                int mode, int32_t* regions_host, int cap, int* n_regions);

/* ---- PCM_16 sample path (the decode of `voice_activity.load_audio`,
 * voice_activity.py:32-69, and the encode of `sf.write`, silencer_ui.py:998, moved onto the device) -------------
 * The `_pcm16` variants of the three file-level calls take the int16 samples of a mono PCM_16 file as stored and
 * decode them inside K1 exactly as libsndfile's float read does (sample / 32768, exact in float32): results are
 * bit-identical to the float32 calls on `pcm / 32768`, with half the bytes crossing PCIe and HBM. */
SS_API int ss_detect_device_pcm16(ss_ctx* ctx, const int16_t* pcm_dev, int64_t n_samples, int mode,
                                  int32_t* regions_dev, int32_t* n_regions_dev, int cap, float* logits_out_dev,
                                  void* stream);
SS_API int ss_detect_host_pcm16(ss_ctx* ctx, const int16_t* pcm_host, int64_t n_samples, int mode,
                                int32_t* regions_host, int cap, int* n_regions, float* logits_host);
SS_API int ss_detect_host_batch_pcm16(ss_ctx* ctx, int n_clips, const int16_t* const* pcm_host,
                                      const int64_t* n_samples, int mode, int32_t* regions_host, int cap,
                                      int* n_regions);

/* `sf.read(dtype='float32')` + `librosa.to_mono` (voice_activity.py:37,61-62) for interleaved PCM_16 frames
 * [n_frames][channels] -> float32 mono [n_frames]: sample / 32768, float32 sum over channels (exact for PCM_16),
 * one float32 division by the channel count. */
SS_API int ss_decode_pcm16(ss_ctx* ctx, const int16_t* interleaved_dev, int64_t n_frames, int channels,
                           float* mono_dev, void* stream);

/* `sf.write` float32 -> PCM_16 (silencer_ui.py:998): short(lrintf(x * 32767.0f)) as libsndfile's f2les_array with
 * normalisation on; out-of-range samples saturate.  Restated from libsndfile's source, not pinned by a golden
 * vector (the library is not in the build image).  Both buffers 16-byte aligned. */
SS_API int ss_encode_pcm16(ss_ctx* ctx, const float* src_dev, int64_t n_elems, int16_t* dst_dev, void* stream);

/* K7 on int16 samples: "Silence Voices" for a PCM_16 file that never leaves its storage format.  Zeroes the
 * element ranges like ss_silence; with requantize != 0 every sample first goes through the reference's
 * read -> write round trip, encode(decode(k)) (see ss_encode_pcm16), so that the buffer equals what
 * `sf.write(librosa.load(...))` would store.  The _host variant streams a host buffer through the device. */
SS_API int ss_silence_pcm16(ss_ctx* ctx, int16_t* pcm_dev, int64_t n_elems, const ss_interval* intervals_dev,
                            int n_intervals, int requantize, void* stream);
SS_API int ss_silence_pcm16_host(ss_ctx* ctx, int16_t* pcm_host, int64_t n_elems, const ss_interval* intervals_host,
                                 int n_intervals, int requantize);

/* Device-level entry points (ss_classify, ss_detect_device) only enqueue work, so they cannot report what the
 * kernels found: this call synchronises `stream` and returns SS_E_CUDA if a tcgen05 pipeline wait timed out or
 * SS_E_RANGE if an fp16-operand mode saturated an activation since the last check (the host-level calls
 * ss_detect_host / ss_detect_host_batch check by themselves). */
SS_API int ss_check_health(ss_ctx* ctx, void* stream);

/* SilenceWorker.run on one HOST buffer `(channels, n)` float32: zero [begin, end) of every
 * interval (element offsets into the flattened buffer) on the device and copy the result back. */
SS_API int ss_silence_host(ss_ctx* ctx, float* pcm_host, int64_t n_elems, const ss_interval* intervals_host,
                    int n_intervals);

/* Test instrumentation (parity localisation, not part of the drop-in surface): copy internal activation
 * `which` of the last tensor-core ss_classify call to out_dev as NCHW float32 and report its shape.
 * which: 0-3 conv1..conv4, 4 bottleneck, 5-8 up(encoder_out) .. up(conv8), 9 conv9, 10 conv1_1's intermediate,
 * 11 the im2col'd mel operand (legacy form), 12 / 13 MaxPool(conv1) / MaxPool(conv2); + 0x100: twice the hi operands
 * alone, + 0x200: twice the lo operands alone (split precision).
 * Also surfaces a tcgen05 pipeline time-out of that call as SS_E_CUDA. */
SS_API int ss_debug_activation(ss_ctx* ctx, int which, int n_windows, float* out_dev, int* C, int* H, int* W,
                               void* stream);

/* Test instrumentation: every device allocation of the context (activation tensors, scratch, staging) sits between
 * guard bands holding a byte pattern; this call synchronises the device and counts the guard bytes that no longer
 * hold it (0 = no kernel wrote outside its buffers).  n_bands (optional) receives the number of bands checked. */
SS_API int ss_debug_check_guards(ss_ctx* ctx, uint64_t* bad_bytes, int* n_bands);

/* Test instrumentation: choose which tcgen05 conv launch of the next tensor-core ss_classify call (same mode as the last one) records
 * per-CTA role timers (-1: none) and read back the previous capture ([148][8] int64 cycles; NULL to skip). */
SS_API int ss_debug_tc_profile(ss_ctx* ctx, int select_launch, long long* out_host);

#ifdef __cplusplus
}
#endif
#endif /* SOFTSPOKEN_B200_H */
