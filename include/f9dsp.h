/* =============================================================================
 * f9dsp.h -- C ABI of the B200-native DSP path of F9 Batch Resampler.
 *
 * This is the drop-in boundary.  Plain C: pointers, sizes, POD
 * structs, int status codes; no C++ exceptions, no torch types.  The reference is
 * a single-process JUCE/C++ app whose DSP helpers are private members of
 * MainComponent called on the message thread; every entry point below names the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Threading: one f9_context per host thread / GPU.  A context is not thread-safe;
 * the library is re-entrant across contexts (process-wide state is limited to driver entry points
 * resolved under std::call_once and an atomic per-device attribute mask; tests/test_gpu_multi.py runs two
 * contexts from two threads).  Calls block until the result is in
 * the caller's buffers unless the name ends in _async or takes device pointers
 * (section F), which only enqueue on the context's stream.
 *
 * Ownership: the caller owns and allocates every input and output buffer.  The
 * library never frees caller memory.  Handles are opaque.
 *
 * There is NO CPU fallback: every compute entry point runs CUDA kernels built for
 * sm_100a and fails with F9_ERR_NO_DEVICE / F9_ERR_CUDA when it cannot.
 * ============================================================================= */
#ifndef F9DSP_H
#define F9DSP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define F9_API __attribute__((visibility("default")))
#else
#define F9_API
#endif

#define F9_VERSION_MAJOR 0
#define F9_VERSION_MINOR 2

/* ---- status codes (reference: sentinels + log lines, Source/MainComponent.cpp:974, :777-782;
 *      Swift typed errors AudioProcessingService.swift:16-52) ---------------------------- */
enum {
    F9_OK              = 0,
    F9_ERR_INVALID     = -1,  /* bad argument                                   */
    F9_ERR_CUDA        = -2,  /* a CUDA call failed; see f9_last_error          */
    F9_ERR_NOMEM       = -3,  /* host or device allocation failed               */
    F9_ERR_NO_DEVICE   = -4,  /* no usable sm_100 device                        */
    F9_ERR_UNSUPPORTED = -5   /* valid request this build does not implement    */
};

/* interpolator kinds: juce::Interpolators::{WindowedSinc, Lagrange, CatmullRom, Linear, ZeroOrderHold}
 * (JUCE 8.0.10 juce_audio_basics/utilities/juce_Interpolators.h; linked by the reference at
 * JuceLibraryCode/JuceHeader.h:16) */
enum { F9_WINDOWED_SINC = 0, F9_LAGRANGE = 1, F9_CATMULL_ROM = 2, F9_LINEAR = 3, F9_ZERO_ORDER_HOLD = 4 };

/* PCM sample formats of the deinterleave / format-convert stage (JUCE AudioFormatReader::read at
 * Source/MainComponent.cpp:734-739; 24-bit writer at :784-801) */
enum { F9_PCM_U8 = 1, F9_PCM_S16LE = 2, F9_PCM_S24LE = 3, F9_PCM_S32LE = 4, F9_PCM_F32LE = 5 };

/* tail-silence predicate: C++ RMS form (Source/MainComponent.cpp:863-882) or Swift peak form
 * (AudioProcessingService.swift:710-737) */
enum { F9_TAIL_RMS = 0, F9_TAIL_PEAK = 1 };

typedef struct f9_context f9_context;
typedef struct f9_interp  f9_interp;
typedef struct f9_plan    f9_plan;
typedef struct f9_resampling_source f9_resampling_source;

/* =============================== A. context ================================== */
/* device: CUDA ordinal.  Fails with F9_ERR_NO_DEVICE when there is no GPU. */
F9_API int  f9_context_create(int device, f9_context** out);
F9_API void f9_context_destroy(f9_context* ctx);
/* Human-readable text of the last failure on this context (ctx may be NULL for creation errors). */
F9_API const char* f9_last_error(const f9_context* ctx);
/* Run on a caller-owned CUDA stream (a cudaStream_t, e.g. torch's current stream); NULL restores
 * the context's own stream. */
F9_API int  f9_set_stream(f9_context* ctx, void* cuda_stream);
F9_API int  f9_synchronize(f9_context* ctx);
/* Variant switches for tests and development: which kernel generation / feed a plan created afterwards takes (names as in
 * "F9_NO_UMMA", "F9_UMMA_NB", "F9_SHORT_UMMA", "F9_UMMA_NOTMA", "F9_UMMA_NOCTA2", "F9_NO_HANKEL", "F9_BATCH_CHUNK_MB" ...).
 * Every variant computes the same results within the stated tolerances; the defaults are the measured-fastest choices.  The
 * library never reads the environment. */
F9_API int  f9_context_set_option(f9_context* ctx, const char* name, int value);
F9_API int  f9_context_clear_options(f9_context* ctx);
/* Kernels launched by this context since creation (bench.py's gpu_launches claim). */
F9_API long long f9_launch_count(const f9_context* ctx);
/* Pinned host memory for full-speed H2D/D2H of caller buffers (optional; any host pointer works). */
F9_API int  f9_host_alloc(f9_context* ctx, void** out, size_t bytes);
F9_API int  f9_host_free(f9_context* ctx, void* p);
F9_API int  f9_version(void);                 /* major*100 + minor */
F9_API int  f9_device_count(void);            /* 0 when there is no usable device; never fails */
/* Host-only diagnostic of the tensor-core FIR tables (no GPU needed): plans the rational ratio p/q for `kind` and checks
 * that every (output slot, tap) sits in exactly one position of the fp16 weight tiles and that head + tail/2048
 * reproduces the fp32 polyphase weight.  Returns the largest reconstruction error, or < 0: -1 bad arguments, -2 no
 * tensor-core plan fits (the CUDA-core kernels serve the ratio), -3..-5 table defects.  info (8 ints, may be NULL):
 * scale m, slots per group, groups, groups per block, blocks, accumulator pool slots, split step, shared-memory bytes. */
F9_API double f9_umma_selfcheck(int kind, long long p, long long q, int* info);
/* Same for the Hankel-operand FIR that serves integer upsampling 1:L (L = 2, 4, 8, 16): largest reconstruction error of the
 * weight image, -1 bad arguments, -3 table defect.  info (4 ints, may be NULL): K steps, fp16 elements per tile buffer, buffer
 * bytes, shared-memory bytes of the kernel. */
F9_API double f9_hankel_selfcheck(int kind, int L, int* info);

/* ============================ B. settings math =============================== */
/* Host scalars; mirror ProcessingSettings (Source/AppState.h:183-259). */
F9_API int    f9_recording_length(int source_frames, int latency_frames);              /* AppState.h:240-243 */
F9_API float  f9_noise_floor_threshold_db(int has_nf, float nf_db, float margin_pct);  /* AppState.h:252-258 */
F9_API float  f9_threshold_linear(float threshold_db);                                 /* AppState.h:246-249 */
F9_API double f9_latency_ms(int measured_latency_samples, double sample_rate);         /* AppState.h:231-237 */
F9_API int    f9_needs_latency_remeasurement(int measured_latency_samples,
                                             int last_buffer_size, int buffer_size);   /* AppState.h:221-228 */

/* ============== C. reference-shaped helpers on HOST planar buffers =========== */
/* `ch` = numCh pointers to numFrames float32 each (juce::AudioBuffer<float> read pointers). */

/* MainComponent::findPeakPosition, Source/MainComponent.cpp:950-975.  *out_pos = frame index of the
 * first maximum of |x| in channel-major scan order, or -1 when max <= threshold. */
F9_API int f9_find_peak_position(f9_context* ctx, const float* const* ch, int numCh, int numFrames,
                                 float threshold, int* out_pos);
/* Swift analyzeCapturedAudio, LatencyMeasurementService.swift:147-171 (interleaved index, default 0;
 * *out_found = 0 where Swift throws noImpulseDetected). */
/* The two calls of the latency measurement (Source/MainComponent.cpp:266-284) as one: findPeakPosition(buffer, threshold) and
 * calculateNoiseFloorDb(buffer) from one upload and one read of the capture. */
F9_API int f9_measure_latency(f9_context* ctx, const float* const* ch, int numCh, int numFrames, float threshold,
                              int* out_pos, float* out_noise_floor_db);
F9_API int f9_find_peak_interleaved(f9_context* ctx, const float* audio, long long n, float threshold,
                                    long long* out_index, int* out_found);
/* MainComponent::calculateRMS, Source/MainComponent.cpp:983-1004. */
F9_API int f9_calculate_rms(f9_context* ctx, const float* const* ch, int numCh, int numFrames, float* out_rms);
/* MainComponent::calculateNoiseFloorDb, Source/MainComponent.cpp:977-981. */
F9_API int f9_calculate_noise_floor_db(f9_context* ctx, const float* const* ch, int numCh, int numFrames,
                                       float* out_db);
/* MainComponent::isReverbTailBelowNoiseFloor, Source/MainComponent.cpp:863-882 (threshold from
 * AppState.h:252-258). */
F9_API int f9_is_reverb_tail_below_noise_floor(f9_context* ctx, const float* const* ch, int numCh, int numFrames,
                                               int has_nf, float nf_db, float margin_pct, int* out_below);
/* Swift isReverbTailBelowNoiseFloor, AudioProcessingService.swift:710-737 (peak based, interleaved). */
F9_API int f9_is_reverb_tail_below_noise_floor_swift(f9_context* ctx, const float* window, long long n,
                                                     int has_nf, float nf_db, float margin_pct, int* out_below);
/* MainComponent::trimLatency, Source/MainComponent.cpp:824-861.  latency_samples is INTERLEAVED samples
 * (:828); out = numCh x original_length, zero padded; *out_copied = frames actually copied. */
F9_API int f9_trim_latency(f9_context* ctx, const float* const* captured, int numCh, int captured_frames,
                           int latency_samples, int original_length, float* const* out, int* out_copied);
/* Swift trimLatency, AudioProcessingService.swift:681-703 (interleaved, no padding).  out must hold
 * source_frames*channels floats; *out_count = samples written. */
F9_API int f9_trim_latency_swift(f9_context* ctx, const float* captured, long long count, long long latency_samples,
                                 long long source_frames, int channels, float* out, long long* out_count);
/* MainComponent::removeDCOffset, Source/MainComponent.cpp:884-902 (in place).  The reference accumulates the channel's sum in ONE
 * float, sample after sample (:892-896): on a long quiet capture that accumulator drifts from the true mean by far more than
 * 2^-20 (a 60 s, 44.1 kHz capture with DC 0.003 and 1e-4 of noise: 2.5e-5).  f9_remove_dc_offset reproduces the reference's
 * arithmetic BIT FOR BIT (the chain is walked in order by one thread per channel, ~2 ms per 10^6 frames, all channels at once).
 * f9_remove_dc_offset_ex(..., reference_order = 0) subtracts the exactly rounded mean instead (parallel double sum: faster, and
 * closer to the true mean than the reference is -- a deliberate accuracy deviation, bounded only by the reference's own drift). */
F9_API int f9_remove_dc_offset(f9_context* ctx, float* const* ch, int numCh, int numFrames);
F9_API int f9_remove_dc_offset_ex(f9_context* ctx, float* const* ch, int numCh, int numFrames, int reference_order);

/* MainComponent::generateImpulse, Source/MainComponent.cpp:934-945 (Swift sendImpulse, LatencyMeasurementService.swift:130-145):
 * clears the buffer and writes 0.9 to sample 0 of every channel -- the stimulus of the latency measurement. */
F9_API int f9_generate_impulse(f9_context* ctx, float* const* ch, int numCh, int numFrames);
/* MainComponent::generateSineWave, Source/MainComponent.cpp:907-932: data[i] = amplitude * sin(phase) with a FLOAT phase chain
 * (phase += (frequency * 2 pi) / sample_rate, wrapped at 2 pi), the same chain on every channel.  *phase_io is sinePhase: read at
 * entry, updated as :929-931 does (phase + inc * numSamples, one wrap).  callback_form != 0: the audio callback's variant
 * (:141-167), where sinePhase is the chain itself.  The reference fixes amplitude 0.5 and frequency 1000 (MainComponent.h:153-154).
 * The phase chain is bit-exact; samples are sin() correctly rounded (the host libm's sinf may differ by one ulp in rare cases). */
F9_API int f9_generate_sine_wave(f9_context* ctx, float* const* ch, int numCh, int numSamples, float frequency, float sample_rate,
                                 float amplitude, float* phase_io, int callback_form);
/* Swift SineWaveGenerator.generateSineWave, SineWaveGenerator.swift:35-59: DOUBLE phase, interleaved output, every channel of a
 * frame gets Float(sin(phase)) * amplitude; *phase_io is the generator's phase member. */
F9_API int f9_generate_sine_wave_swift(f9_context* ctx, float* buffer, int frame_count, int channel_count, double frequency,
                                       double sample_rate, float amplitude, double* phase_io);

/* Offline form of the reverb-mode stop loop (Swift AudioProcessingService.swift:423-453; C++ intent
 * claude.md:346-367).  Poll i tests the last `window` frames before e_i = start_frame + (i+1)*hop;
 * the first poll that makes `required` consecutive silent polls gives *out_stop_frame = e_i (else -1).
 * flags (optional, capacity max_flags) receives -1 skipped / 0 / 1 per poll; *out_polls the poll count. */
F9_API int f9_tail_scan(f9_context* ctx, const float* const* ch, int numCh, long long numFrames,
                        long long start_frame, int window, int hop, int required, int mode,
                        int has_nf, float nf_db, float margin_pct,
                        long long* out_stop_frame, int* flags, int max_flags, int* out_polls);

/* Bounded-lag cross-correlation with argmax (north_star (b); the reference peak-picks an impulse,
 * LatencyMeasurementService.swift:164).  r_c[lag] = sum_i x[i]*y_c[i+lag] over lag in [lag_min, lag_max],
 * scan order channel-major / lag ascending, strict '>' on |r| (ties keep the earliest lag of the lowest
 * channel, as findPeakPosition).  *out_found = max|r| > threshold*||x||_2. */
F9_API int f9_xcorr_peak(f9_context* ctx, const float* const* y, int numCh, int numFrames,
                         const float* x, int stim_len, int lag_min, int lag_max, float threshold,
                         int* out_found, int* out_lag, int* out_ch, double* out_value);

/* ============ D. juce::Interpolators-shaped stateful objects ================ */
/* [JUCE 8.0.10 juce_GenericInterpolator.h]  One object per channel, as JUCE.  State (the last
 * memorySize inputs and the sub-sample position) lives on the host; each process() runs the FIR on
 * the GPU.  `in` must hold at least the returned number of samples. */
F9_API int   f9_interp_create(f9_context* ctx, int kind, f9_interp** out);
F9_API void  f9_interp_destroy(f9_interp* h);
F9_API int   f9_interp_reset(f9_interp* h);                                     /* reset()          */
F9_API float f9_interp_base_latency(const f9_interp* h);                        /* getBaseLatency() */
/* int process(double speedRatio, const float* in, float* out, int numOut) -> inputs consumed (<0: error) */
F9_API int   f9_interp_process(f9_interp* h, double speed_ratio, const float* in, float* out, int num_out);
/* processAdding(..., gain) */
F9_API int   f9_interp_process_adding(f9_interp* h, double speed_ratio, const float* in, float* out,
                                      int num_out, float gain);
/* process(..., numInputSamplesAvailable, wrapAround) */
F9_API int   f9_interp_process_wrap(f9_interp* h, double speed_ratio, const float* in, float* out,
                                    int num_out, int num_in_available, int wrap_around);
/* WindowedSincTraits::lookupTable[10001].  JUCE's literal table is not in the reference; the built-in
 * default is sinc*Hann.  A host that has JUCE can install the real table here.  Call it before creating plans
 * and interpolators: those created earlier keep the table they were built with (their polyphase tables stay cached until the
 * context is destroyed); everything created afterwards uses the new one. */
F9_API int   f9_sinc_table_set(f9_context* ctx, const float* table10001);
F9_API int   f9_sinc_table_get(const f9_context* ctx, float* table10001);

/* ============ D2. juce::ResamplingAudioSource-shaped object ============ */
/* [JUCE 8.0.10 juce_audio_basics/sources/juce_ResamplingAudioSource.{h,cpp}; named by north_star next to the
 * Interpolators, no call site in the reference]  Linear interpolation with a double sub-sample position plus a 2nd-order
 * Butterworth low-pass (double state): on the pulled input when ratio > 1.0001, on the output when ratio < 0.9999.
 * One object carries numChannels channels, as JUCE.  The caller plays the role of the input AudioSource: `in` holds the
 * next samples that source would deliver; past num_in_available it delivers zeros (AudioFormatReaderSource past the end
 * of a file).  Control flow on the host as in JUCE, filters and interpolation on the GPU in JUCE's operation order. */
F9_API int    f9_ras_create(f9_context* ctx, int num_channels, f9_resampling_source** out);
F9_API void   f9_ras_destroy(f9_resampling_source* h);
F9_API int    f9_ras_set_resampling_ratio(f9_resampling_source* h, double samples_in_per_output_sample); /* setResamplingRatio */
F9_API double f9_ras_get_resampling_ratio(const f9_resampling_source* h);                                 /* getResamplingRatio */
F9_API int    f9_ras_prepare_to_play(f9_resampling_source* h, int samples_per_block_expected, double sample_rate);
F9_API int    f9_ras_flush_buffers(f9_resampling_source* h);                                              /* flushBuffers      */
F9_API int    f9_ras_release_resources(f9_resampling_source* h);                                          /* releaseResources  */
/* applyFilter's JUCE_INTEL branch (filter outputs within +-1e-8 are flushed to 0): on by default (x86 build of the reference) */
F9_API int    f9_ras_set_intel_denormal_flush(f9_resampling_source* h, int on);
/* getNextAudioBlock(info): fills out[c][0 .. num_samples).  Returns the number of samples pulled from the input source
 * (round(num_samples * ratio) + 3 - samples still buffered; the caller advances its read position by min(that,
 * num_in_available)), or a negative status. */
/* how many samples the next getNextAudioBlock(num_samples) will pull from the input source */
F9_API int    f9_ras_num_samples_to_pull(const f9_resampling_source* h, int num_samples);
F9_API int    f9_ras_get_next_audio_block(f9_resampling_source* h, const float* const* in, int num_in_available,
                                          float* const* out, int num_samples);
/* Whole channels from reset state (prepareToPlay + getNextAudioBlock until num_out), all channels in one pass; the IIR is
 * evaluated chunk-parallel (tolerance parity).  Host buffers / device buffers (planar, strides in floats;
 * d_scratch: num_streams x scratch_stride floats with scratch_stride >= f9_ras_scratch_frames(ratio, num_out)). */
F9_API int    f9_ras_convert(f9_context* ctx, const float* const* in, int numCh, long long num_in, double ratio,
                             float* const* out, long long num_out);
F9_API long long f9_ras_scratch_frames(double ratio, long long num_out);
F9_API int    f9_dev_ras_convert(f9_context* ctx, const float* d_in, long long in_stride, int num_streams, long long num_in,
                                 double ratio, float* d_out, long long out_stride, long long num_out,
                                 float* d_scratch, long long scratch_stride);

/* ======================= E. batch job flow (host buffers) =================== */
/* One job = one file of the MainComponent/AppState batch flow (Source/MainComponent.cpp:705-805;
 * Swift processFiles AudioProcessingService.swift:66-113, :339-536): captured recording ->
 * [tail-silence scan] -> trimLatency -> [removeDCOffset] -> [sample-rate conversion] -> planar float
 * (and optionally interleaved 24-bit PCM, the WAV payload). */
typedef struct f9_job {
    const float* const* captured;   /* numCh planar channels, captured_frames each (host)          */
    int   numCh;
    int   captured_frames;
    int   latency_samples;          /* interleaved samples, as measuredLatencySamples (AppState.h:197) */
    int   original_length;          /* source length in frames (trimLatency's originalLength)       */
    double fs_in, fs_out;           /* conversion ratio = fs_in / fs_out; equal => no conversion     */
    int   interp_kind;              /* F9_WINDOWED_SINC / F9_LAGRANGE / ...                          */
    int   flags;                    /* F9_JOB_* below                                               */
    /* tail scan (used when F9_JOB_TAIL_SCAN): */
    int   tail_window, tail_hop, tail_required, tail_mode;
    int   has_nf; float nf_db; float margin_pct;
    /* outputs (caller-allocated): */
    float* const* out;              /* numCh channels, out_capacity frames each; NULL: no float download (F9_JOB_PCM24 only) */
    int   out_capacity;
    unsigned char* out_pcm24;       /* optional: numCh*out_frames*3 bytes interleaved, or NULL      */
    /* The capture as the file / device holds it (optional; when src_pcm != NULL it replaces `captured`, which may then be NULL):
     * captured_frames frames of src_ch interleaved channels in format src_fmt (F9_PCM_*), i.e. what reader->read consumes at
     * Source/MainComponent.cpp:734-739.  The deinterleave / int->float stage then runs on the device and the upload is the file's
     * own 2 or 3 bytes per sample instead of 4.  Plane c of the job reads source channel min(c, src_ch - 1). */
    const void* src_pcm;
    int   src_fmt;
    int   src_ch;
} f9_job;

enum { F9_JOB_TAIL_SCAN = 1, F9_JOB_REMOVE_DC = 2, F9_JOB_PCM24 = 4,
       F9_JOB_DC_REFERENCE_ORDER = 8 };  /* with F9_JOB_REMOVE_DC: the reference's sequential float accumulator (bit-exact dcOffset; ~2 ms per
                                          * 10^6 frames) instead of the exactly rounded mean (see f9_remove_dc_offset) */

typedef struct f9_result {
    int status;                     /* F9_OK or an error for this job                               */
    int latency_frames;             /* latency_samples / numCh (MainComponent.cpp:835)              */
    int trim_start;                 /* first captured frame used                                    */
    int frames_copied;              /* frames taken from the capture (rest of original_length is 0) */
    int out_frames;                 /* frames written per output channel                            */
    long long tail_stop_frame;      /* stop frame of the tail scan, -1 none / not requested         */
    int tail_polls;
} f9_result;

/* Output length of a conversion of n_in frames: ceil(n_in * fs_out / fs_in) in exact integers. */
F9_API long long f9_resampled_length(long long n_in, double fs_in, double fs_out);
/* Blocking.  The batch is cut into chunks of device memory and pipelined: uploads, kernels and downloads of different chunks overlap
 * (three streams of the context, event hand-offs, no host wait between chunks).  Host buffers given as pinned (page-locked) memory
 * are transferred by DMA straight from / into them; all of them must stay valid and untouched until the call returns.  Returns
 * the worst job status; results[i] is filled for every job. */
F9_API int f9_process_batch(f9_context* ctx, const f9_job* jobs, int n_jobs, f9_result* results);

/* ============== F. device-resident entry points (pointers in HBM) =========== */
/* Same operations on buffers that already live on the device; they enqueue on the context's stream
 * and return without synchronising.  Used for pipelining, for sharding one batch over several GPUs
 * (one context per GPU) and by bench.py's HBM-resident leg. */
typedef struct f9_dev_buffer {      /* planar float32 on the device                                 */
    const float* base;              /* channel c starts at base + c*ch_stride                       */
    long long ch_stride;            /* in floats                                                    */
    int numCh;
    int numFrames;
} f9_dev_buffer;

/* batched findPeakPosition: d_out_pos[i] (device int) for each buffer */
F9_API int f9_dev_find_peak_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, float threshold, int* d_out_pos);
/* The latency measurement reads each capture twice in the reference -- findPeakPosition, then calculateNoiseFloorDb over the
 * same buffer (Source/MainComponent.cpp:270-279).  This form serves both from one read: d_out_pos[i] as above, d_sumsq[i] the sum
 * of squares (double; rms = sqrt(sumsq / (numCh * numFrames)), noise floor = 20 log10f(max(rms, 1e-6f))), d_peak[i] (optional) max |x|. */
F9_API int f9_dev_latency_stats_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, float threshold, int* d_out_pos,
                                      double* d_sumsq, float* d_peak);
/* batched sum of squares (double) and peak per buffer: d_sumsq[i], d_peak[i] (device) */
F9_API int f9_dev_stats_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, double* d_sumsq, float* d_peak);

/* batched bounded-lag cross-correlation argmax against one stimulus (device float array):
 * d_out[i] = { max |r|, channel (-1: all zero), lag } per buffer; the threshold test
 * (value > threshold*||x||_2) is the caller's scalar. */
typedef struct f9_xcorr_result { double value; int ch; int lag; int reserved; } f9_xcorr_result;
F9_API int f9_dev_xcorr_peak_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, const float* d_stim,
                                   int stim_len, int lag_min, int lag_max, f9_xcorr_result* d_out);

/* One channel-segment of a sample-rate conversion.  Output samples [n0, n0+num_out) of the conversion
 * of a channel from reset state are written to out[0..num_out).  in[0] is sample `in_offset` of the
 * channel and in_avail samples are present; samples outside [in_offset, in_offset+in_avail) read as 0
 * (before 0: interpolator reset state; after the end: the 6-argument process() pushes zeros).
 * Time-segmenting a long file = several segments with different n0 whose `in` windows overlap by the
 * interpolator memory (halo 199 for WindowedSinc, 4 for Lagrange). */
typedef struct f9_resample_seg {
    const float* in;   long long in_offset; long long in_avail;
    float* out;        long long n0;        long long num_out;
} f9_resample_seg;

/* Build a reusable plan: segment table + per-ratio polyphase tables uploaded once. */
F9_API int  f9_resample_plan_create(f9_context* ctx, int kind, double speed_ratio,
                                    const f9_resample_seg* segs, int n_segs, f9_plan** out);
F9_API int  f9_resample_plan_run(f9_plan* plan);          /* enqueue; no sync */
F9_API void f9_plan_destroy(f9_plan* plan);
/* Input halo a segment starting at output n0 needs before its first fresh input, and the index of
 * the first input sample it touches (for host-side segmentation across GPUs). */
F9_API int  f9_resample_segment_input_range(int kind, double speed_ratio, long long n0, long long num_out,
                                            long long* first_in, long long* last_in_plus1);

/* Tail-scan plan over device buffers: d_stop_frame[i] (device long long), d_flags optional
 * (n * max_polls ints). */
typedef struct f9_tail_params {
    long long start_frame; int window, hop, required, mode; int has_nf; float nf_db, margin_pct;
} f9_tail_params;
F9_API int f9_dev_tail_scan_batch(f9_context* ctx, const f9_dev_buffer* bufs, const f9_tail_params* params, int n,
                                  long long* d_stop_frame, int* d_flags, int max_polls);

/* fused trim (+ optional DC removal) on device: out buffer i = trimLatency(bufs[i], latency_samples[i],
 * original_length[i]); out numFrames must equal original_length[i].  remove_dc: 0 no, 1 exactly rounded mean (parallel),
 * 2 the reference's sequential float accumulator (bit-exact, see f9_remove_dc_offset). */
F9_API int f9_dev_trim_batch(f9_context* ctx, const f9_dev_buffer* captured, const int* latency_samples,
                             const f9_dev_buffer* out, int n, int remove_dc);

/* stimuli on device buffers (bufs[i].base is written): generateImpulse for a batch, generateSineWave for one buffer */
F9_API int f9_dev_generate_impulse(f9_context* ctx, const f9_dev_buffer* bufs, int n);
F9_API int f9_dev_generate_sine_wave(f9_context* ctx, const f9_dev_buffer* buf, float frequency, float sample_rate, float amplitude,
                                     float phase0);

/* ======================= G. deinterleave / format convert =================== */
/* interleaved PCM (host) -> planar float (host): JUCE reader semantics (left-justify to int32, scale by
 * 1/0x7fffffff); destination channel c reads source channel min(c, src_ch-1). */
F9_API int f9_pcm_to_planar(f9_context* ctx, const void* src, int fmt, int src_ch, long long num_frames,
                            float* const* dst, int dst_ch);
/* planar float (host) -> interleaved 24-bit LE PCM (host): JUCE writer semantics (clip, round half even
 * of INT_MAX*x in double, keep top 24 bits). */
F9_API int f9_planar_to_pcm24(f9_context* ctx, const float* const* src, int numCh, long long num_frames,
                              unsigned char* dst);
/* planar <-> interleaved float (AudioProcessingService.swift:361-365, :524-531). */
F9_API int f9_interleave(f9_context* ctx, const float* const* src, int numCh, long long num_frames, float* dst);
F9_API int f9_deinterleave(f9_context* ctx, const float* src, int numCh, long long num_frames, float* const* dst);
/* device-resident forms */
F9_API int f9_dev_pcm_to_planar(f9_context* ctx, const void* d_src, int fmt, int src_ch, long long num_frames,
                                float* d_dst, long long dst_ch_stride, int dst_ch);
F9_API int f9_dev_planar_to_pcm24(f9_context* ctx, const float* d_src, long long src_ch_stride, int numCh,
                                  long long num_frames, unsigned char* d_dst);
/* batched device forms: one launch for a whole batch of files (the per-file calls above are launch-bound on short files).
 * Replaces the per-file reader->read / 24-bit writer passes of the save loop (Source/MainComponent.cpp:734-739, :785-801).
 * pcm_to_planar_batch: file i reads d_srcs[i] (src_ch interleaved channels, dst[i].numFrames frames of format fmt) into the
 * planar buffer dst[i] (dst[i].base is written).  planar_to_pcm24_batch: file i writes src[i] to d_dsts[i]
 * (src[i].numCh * src[i].numFrames * 3 bytes).  d_srcs / d_dsts are host arrays of device pointers. */
F9_API int f9_dev_pcm_to_planar_batch(f9_context* ctx, const void* const* d_srcs, int fmt, int src_ch,
                                      const f9_dev_buffer* dst, int n);
F9_API int f9_dev_planar_to_pcm24_batch(f9_context* ctx, const f9_dev_buffer* src, unsigned char* const* d_dsts, int n);

/* ===================== H. the batch job flow over several GPUs ===================== */
/* The reference's batch loop is one process over AppState.files (Source/MainComponent.cpp:581-621, :705-805).  The path shards
 * with no exchange step: unit of work = a file, or -- for a file that is large against one GPU's share -- a group of its channels
 * and / or a time segment of its conversion (input window with its own halo, f9_resample_segment_input_range); units are packed
 * greedily by output-sample count; one host thread and one f9_context per GPU; results are gathered on the host (no NCCL). */
typedef struct f9_multi f9_multi;
typedef struct f9_unit {
    int job;                        /* index into the job array                                                          */
    int device;                     /* 0 .. n_devices-1: position in the device list, set by f9_multi_partition          */
    int ch0, num_ch;                /* channels [ch0, ch0 + num_ch) of the job                                           */
    long long n0, num_out;          /* outputs [n0, n0 + num_out) of the conversion; num_out == 0: all of them           */
    int tail_only;                  /* the job's reverb-tail scan alone (split jobs)                                     */
    int reserved;
    long long cost;                 /* output samples (packing weight)                                                   */
} f9_unit;
/* devices: CUDA ordinals, one context (and, inside f9_multi_process_batch, one host thread) each.  The same ordinal may appear
 * more than once (two contexts sharing a GPU). */
F9_API int  f9_multi_create(const int* devices, int n_devices, f9_multi** out);
F9_API void f9_multi_destroy(f9_multi* m);
F9_API int  f9_multi_device_count(const f9_multi* m);
F9_API f9_context* f9_multi_context(f9_multi* m, int i);
F9_API const char* f9_multi_last_error(const f9_multi* m);
/* f9_process_batch over all the GPUs of m: partition, one thread per GPU, merge.  out_device (optional, n_jobs ints): the device
 * list position the job (or its first unit) ran on. */
F9_API int  f9_multi_process_batch(f9_multi* m, const f9_job* jobs, int n_jobs, f9_result* results, int* out_device);
/* The pieces, for hosts that run one process per GPU (bench.py under torchrun): every process computes the same partition, runs
 * the units of its own device on its own context and the per-job results are merged wherever they are gathered.
 * f9_shard_units: greedy packing, largest first onto the least loaded bin (ties: lower index). */
F9_API int  f9_shard_units(const long long* costs, int n, int world, int* out_bin);
F9_API int  f9_multi_partition(const f9_job* jobs, int n_jobs, int n_devices, long long seg_out, f9_unit* units, int max_units, int* n_units);
F9_API int  f9_process_units(f9_context* ctx, const f9_job* jobs, int n_jobs, const f9_unit* units, int n_units, int device,
                             f9_result* unit_results);
F9_API int  f9_merge_unit_results(const f9_job* jobs, int n_jobs, const f9_unit* units, const f9_result* unit_results, int n_units,
                                  f9_result* results);

#ifdef __cplusplus
}
#endif
#endif /* F9DSP_H */
