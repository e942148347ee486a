// extern "C" entry points of libf9dsp.so (declared in include/f9dsp.h): context, host-buffer helpers that mirror
// the reference's MainComponent signatures, juce::Interpolators-shaped objects, the batch job flow and the
// device-resident plans.  Host code here only moves data and finishes scalars; every sample is touched by a
// CUDA kernel.  There is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

#include "f9_internal.cuh"

using namespace f9;

static thread_local std::string g_create_error;

// ------------------------------------------------------------------------------------------------ context
// Blocking entry points own the whole arena: they start from offset 0 and finish with F9_FINISH.  Entry points
// that only enqueue (async_call) keep bumping while earlier enqueued work may still read the arena, and fall
// back to a stream sync when it is full.
int f9_context::arena_reserve(size_t d_bytes, size_t h_bytes, bool async_call) {
    d_bytes += 4096; h_bytes += 4096;
    if (!quiescent) {
        const bool fits = d_used + d_bytes + 512 <= d_cap && h_used + h_bytes + 512 <= h_cap;
        if (async_call && fits) return F9_OK;
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize");
        quiescent = true;
    }
    quiescent = false;          // work is about to be enqueued; F9_FINISH sets it back
    if (d_bytes > d_cap) {
        if (d_arena) { cudaStreamSynchronize(stream); cudaFree(d_arena); d_arena = nullptr; d_cap = 0; }
        size_t want = std::max(std::max(d_bytes, d_cap + d_cap / 2), size_t(8) << 20);
        cudaError_t e = cudaMalloc((void**) &d_arena, want);
        if (e != cudaSuccess) { err = std::string("cudaMalloc(arena): ") + cudaGetErrorString(e); cudaGetLastError(); return F9_ERR_NOMEM; }
        d_cap = want;
    }
    if (h_bytes > h_cap) {
        if (h_arena) { cudaStreamSynchronize(stream); cudaFreeHost(h_arena); h_arena = nullptr; h_cap = 0; }
        size_t want = std::max(std::max(h_bytes, h_cap + h_cap / 2), size_t(4) << 20);
        cudaError_t e = cudaHostAlloc((void**) &h_arena, want, cudaHostAllocDefault);
        if (e != cudaSuccess) { err = std::string("cudaHostAlloc(arena): ") + cudaGetErrorString(e); cudaGetLastError(); return F9_ERR_NOMEM; }
        h_cap = want;
    }
    arena_reset();
    return F9_OK;
}

int f9_context::get_poly(int kind, long long p, long long q, PolyDev* out) {
    PolyKey key{kind, p, q, sinc_epoch};
    auto it = poly_cache.find(key);
    if (it != poly_cache.end()) { *out = it->second; return F9_OK; }
    PolyHost H;
    build_poly(kind, sinc_table.data(), p, q, &H);
    PolyDev D; D.p = H.p; D.q = H.q; D.taps = H.taps; D.qpad = H.qpad;
    F9_TRY_CUDA(this, cudaMalloc((void**) &D.B, sizeof(int) * H.B.size()));
    F9_TRY_CUDA(this, cudaMalloc((void**) &D.W, sizeof(float) * H.W.size()));
    F9_TRY_CUDA(this, cudaMemcpyAsync(D.B, H.B.data(), sizeof(int) * H.B.size(), cudaMemcpyHostToDevice, stream));
    F9_TRY_CUDA(this, cudaMemcpyAsync(D.W, H.W.data(), sizeof(float) * H.W.size(), cudaMemcpyHostToDevice, stream));
    F9_TRY_CUDA(this, cudaStreamSynchronize(stream));
    poly_cache[key] = D;
    *out = D;
    return F9_OK;
}

extern "C" {

int f9_version(void) { return F9_VERSION_MAJOR * 100 + F9_VERSION_MINOR; }

int f9_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int f9_context_create(int device, f9_context** out) {
    if (!out) return F9_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        g_create_error = (e != cudaSuccess) ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)
                                            : std::string("no such CUDA device");
        cudaGetLastError();
        return F9_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return F9_ERR_CUDA; }
    if (prop.major != 10 || prop.minor != 0) {
        // sm_100a SASS is architecture-specific: it loads on compute capability 10.0 only (not on 10.3 or 12.x)
        g_create_error = "device is not compute capability 10.0 (B200): this library carries sm_100a code only";
        return F9_ERR_NO_DEVICE;
    }
    f9_context* ctx = new (std::nothrow) f9_context();
    if (!ctx) return F9_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); delete ctx; return F9_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    ctx->sinc_table.assign(kSincTableSize + 1, 0.0f);
    make_default_sinc_table(ctx->sinc_table.data());
    if ((e = cudaMalloc((void**) &ctx->d_sinc_table, sizeof(float) * (kSincTableSize + 1))) != cudaSuccess ||
        (e = cudaMemcpy(ctx->d_sinc_table, ctx->sinc_table.data(), sizeof(float) * (kSincTableSize + 1), cudaMemcpyHostToDevice)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); f9_context_destroy(ctx); return F9_ERR_CUDA;
    }
    *out = ctx;
    return F9_OK;
}

void f9_context_destroy(f9_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->poly_cache) { cudaFree(kv.second.B); cudaFree(kv.second.W); }
    for (auto& kv : ctx->band_cache) { cudaFree(kv.second.C); cudaFree(kv.second.wmin); }
    for (auto& kv : ctx->umma_cache) cudaFree((void*) kv.second.W);
    for (auto& kv : ctx->hankel_cache) cudaFree((void*) kv.second.W);
    if (ctx->d_sinc_table) cudaFree(ctx->d_sinc_table);
    if (ctx->cur_slot) ctx->swap_slot();
    if (ctx->alt_stream) { cudaStreamSynchronize(ctx->alt_stream); cudaStreamDestroy(ctx->alt_stream); }
    if (ctx->down_stream) { cudaStreamSynchronize(ctx->down_stream); cudaStreamDestroy(ctx->down_stream); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    ctx->arena_reset(); ctx->swap_slot(); ctx->arena_reset(); ctx->swap_slot();      // frees both slots' spill allocations
    if (ctx->parked.d_arena) cudaFree(ctx->parked.d_arena);
    if (ctx->parked.h_arena) cudaFreeHost(ctx->parked.h_arena);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* f9_last_error(const f9_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int f9_set_stream(f9_context* ctx, void* cuda_stream) {
    if (!ctx) return F9_ERR_INVALID;
    cudaStream_t next = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
    if (next == ctx->stream) return F9_OK;
    // The arenas (descriptors, tile prefixes, partial sums) are ordered against the stream that was current when they were
    // carved: work still in flight there must finish before a call on the new stream resets and overwrites them.
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->quiescent) { F9_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ctx->quiescent = true; }
    ctx->stream = next;
    return F9_OK;
}
int f9_context_set_option(f9_context* ctx, const char* name, int value) {
    if (!ctx || !name) return F9_ERR_INVALID;
    ctx->diag.m[name] = value;
    return F9_OK;
}
int f9_context_clear_options(f9_context* ctx) {
    if (!ctx) return F9_ERR_INVALID;
    ctx->diag.m.clear();
    return F9_OK;
}
int f9_synchronize(f9_context* ctx) {
    if (!ctx) return F9_ERR_INVALID;
    F9_FINISH(ctx);
    return F9_OK;
}
long long f9_launch_count(const f9_context* ctx) { return ctx ? ctx->launches : 0; }

int f9_host_alloc(f9_context* ctx, void** out, size_t bytes) {
    if (!ctx || !out) return F9_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return ctx->fail(F9_ERR_NOMEM, "cudaHostAlloc failed"); }
    return F9_OK;
}
int f9_host_free(f9_context* ctx, void* p) {
    if (!ctx) return F9_ERR_INVALID;
    if (p) F9_TRY_CUDA(ctx, cudaFreeHost(p));
    return F9_OK;
}

// ------------------------------------------------------------------------------------------------ settings math
int f9_recording_length(int source_frames, int latency_frames) { return source_frames + latency_frames + (latency_frames * 4); }
float f9_noise_floor_threshold_db(int has_nf, float nf_db, float margin_pct) { return nf_threshold_db(has_nf, nf_db, margin_pct); }
float f9_threshold_linear(float threshold_db) { return threshold_linear(threshold_db); }
double f9_latency_ms(int measured_latency_samples, double sample_rate) {
    if (measured_latency_samples < 0) return 0.0;
    return ((double) measured_latency_samples / sample_rate) * 1000.0;
}
int f9_needs_latency_remeasurement(int measured_latency_samples, int last_buffer_size, int buffer_size) {
    if (measured_latency_samples < 0) return 1;
    return last_buffer_size != buffer_size;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ helpers
namespace {

inline long long pad_stride(long long frames) { return (frames + 63) / 64 * 64; }

// Upload a host planar buffer into the device arena (one cudaMemcpyAsync per channel).
int upload_planar(f9_context* ctx, const float* const* ch, int numCh, long long numFrames, DevBuf* out) {
    const long long stride = pad_stride(std::max<long long>(numFrames, 1));
    float* d = (float*) ctx->d_alloc(sizeof(float) * (size_t) stride * std::max(numCh, 1));
    for (int c = 0; c < numCh; ++c)
        if (numFrames > 0)
            F9_TRY_CUDA(ctx, cudaMemcpyAsync(d + c * stride, ch[c], sizeof(float) * (size_t) numFrames, cudaMemcpyHostToDevice, ctx->stream));
    out->base = d; out->chStride = stride; out->numCh = numCh; out->numFrames = (int) numFrames;
    return F9_OK;
}
inline size_t planar_bytes(int numCh, long long numFrames) {
    return sizeof(float) * (size_t) pad_stride(std::max<long long>(numFrames, 1)) * std::max(numCh, 1) + 256;
}

template <typename T>
int upload_array(f9_context* ctx, const T* h, size_t n, T** d_out) {
    T* d = (T*) ctx->d_alloc(sizeof(T) * std::max<size_t>(n, 1));
    // stage through the pinned arena so the async copy does not read a dying stack/vector buffer
    T* hp = (T*) ctx->h_alloc(sizeof(T) * std::max<size_t>(n, 1));
    if (n) std::memcpy(hp, h, sizeof(T) * n);
    if (n) F9_TRY_CUDA(ctx, cudaMemcpyAsync(d, hp, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
    *d_out = d;
    return F9_OK;
}

int check_planar(f9_context* ctx, const float* const* ch, int numCh, long long numFrames) {
    if (!ctx) return F9_ERR_INVALID;
    if (numCh < 0 || numFrames < 0 || numFrames > 0x7fffffffLL) return ctx->fail(F9_ERR_INVALID, "negative or oversized buffer");
    if (numCh > 0 && !ch) return ctx->fail(F9_ERR_INVALID, "null channel array");
    for (int c = 0; c < numCh; ++c) if (numFrames > 0 && !ch[c]) return ctx->fail(F9_ERR_INVALID, "null channel pointer");
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    return F9_OK;
}

TailParams make_tail_params(long long start, int window, int hop, int required, int mode, int has_nf, float nf_db, float margin) {
    TailParams P{};
    P.startFrame = start; P.window = window; P.hop = hop; P.required = required; P.mode = mode;
    P.noNf = 0; P.below0 = 0;
    if (mode == F9_TAIL_PEAK) {
        if (!has_nf) { P.noNf = 1; P.rstar = -1.0f; }
        else P.rstar = largest_peak_below(nf_db + (nf_db * margin / 100.0f), &P.below0);
    } else {
        P.rstar = largest_rms_below(nf_threshold_db(has_nf, nf_db, margin), 1e-10f);
    }
    return P;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ C. host helpers
int f9_find_peak_position(f9_context* ctx, const float* const* ch, int numCh, int numFrames, float threshold, int* out_pos) {
    int rc = check_planar(ctx, ch, numCh, numFrames); if (rc) return rc;
    if (!out_pos) return ctx->fail(F9_ERR_INVALID, "null out_pos");
    if (numCh == 0 || numFrames == 0) { *out_pos = -1; return F9_OK; }
    DevBuf hb{}; std::vector<int> prefix;
    hb.numCh = numCh; hb.numFrames = numFrames;
    const int total = peak_prefix(&hb, 1, &prefix);
    rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + sizeof(PeakPartial) * (size_t) total + 4096, 4096); if (rc) return rc;
    rc = upload_planar(ctx, ch, numCh, numFrames, &hb); if (rc) return rc;
    DevBuf* d_bufs; int* d_prefix;
    rc = upload_array(ctx, &hb, 1, &d_bufs); if (rc) return rc;
    rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
    PeakPartial* d_part = (PeakPartial*) ctx->d_alloc(sizeof(PeakPartial) * (size_t) total);
    int* d_out = (int*) ctx->d_alloc(sizeof(int));
    F9_TRY_CUDA(ctx, launch_find_peak(d_bufs, 1, total, d_prefix, threshold, d_part, d_out, ctx->stream, &ctx->launches));
    int* h_out = (int*) ctx->h_alloc(sizeof(int));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    *out_pos = *h_out;
    return F9_OK;
}

// The latency measurement of timerCallback (Source/MainComponent.cpp:266-284) in one call: one upload and one read of the capture
// give findPeakPosition(buffer, threshold) and calculateNoiseFloorDb(buffer).
int f9_measure_latency(f9_context* ctx, const float* const* ch, int numCh, int numFrames, float threshold, int* out_pos, float* out_noise_floor_db) {
    int rc = check_planar(ctx, ch, numCh, numFrames); if (rc) return rc;
    if (!out_pos || !out_noise_floor_db) return ctx->fail(F9_ERR_INVALID, "null output");
    if (numCh == 0 || numFrames == 0) { *out_pos = -1; *out_noise_floor_db = noise_floor_db_from_rms(0.0f); return F9_OK; }
    DevBuf hb{}; std::vector<int> prefix;
    hb.numCh = numCh; hb.numFrames = numFrames;
    const int total = peak_prefix(&hb, 1, &prefix);
    rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + (sizeof(PeakPartial) + sizeof(double)) * (size_t) total + 4096, 4096); if (rc) return rc;
    rc = upload_planar(ctx, ch, numCh, numFrames, &hb); if (rc) return rc;
    DevBuf* d_bufs; int* d_prefix;
    rc = upload_array(ctx, &hb, 1, &d_bufs); if (rc) return rc;
    rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
    PeakPartial* d_part = (PeakPartial*) ctx->d_alloc(sizeof(PeakPartial) * (size_t) total);
    double* d_psum = (double*) ctx->d_alloc(sizeof(double) * (size_t) total);
    double* d_res = (double*) ctx->d_alloc(2 * sizeof(double));              // [0] sum of squares, [1] (as int) position
    int* d_pos = reinterpret_cast<int*>(d_res + 1);
    F9_TRY_CUDA(ctx, launch_find_peak(d_bufs, 1, total, d_prefix, threshold, d_part, d_pos, ctx->stream, &ctx->launches, d_psum, d_res, nullptr, (ctx->diag.has("F9_RMS_TREE_SUM") ? -1 : ctx->diag.get("F9_RMS_FORCE_ORDER"))));
    double* h_res = (double*) ctx->h_alloc(2 * sizeof(double));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_res, d_res, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    *out_pos = *reinterpret_cast<const int*>(h_res + 1);
    *out_noise_floor_db = noise_floor_db_from_rms((float) std::sqrt(h_res[0] / (double) ((long long) numCh * numFrames)));
    return F9_OK;
}

int f9_find_peak_interleaved(f9_context* ctx, const float* audio, long long n, float threshold, long long* out_index, int* out_found) {
    if (!ctx) return F9_ERR_INVALID;
    if (n < 0 || n > 0x7fffffffLL || (n > 0 && !audio) || !out_index) return ctx->fail(F9_ERR_INVALID, "bad interleaved buffer");
    // An interleaved stream scanned in order is a one-channel buffer: same strict-> argmax.
    int pos = -1;
    const float* chans[1] = {audio};
    // threshold handled here: Swift keeps index of the max even when it fails the threshold (default 0)
    int rc = f9_find_peak_position(ctx, chans, n > 0 ? 1 : 0, (int) n, -1.0f, &pos);
    if (rc) return rc;
    *out_index = pos < 0 ? 0 : pos;
    if (out_found) {
        float v = 0.0f;
        if (pos >= 0) v = std::fabs(audio[pos]);
        *out_found = (v > threshold) ? 1 : 0;
    }
    return F9_OK;
}

static int stats_one(f9_context* ctx, const float* const* ch, int numCh, int numFrames, double* sumsq, float* peak) {
    int rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + 16384, 4096); if (rc) return rc;
    DevBuf hb{};
    rc = upload_planar(ctx, ch, numCh, numFrames, &hb); if (rc) return rc;
    DevBuf* d_bufs;
    rc = upload_array(ctx, &hb, 1, &d_bufs); if (rc) return rc;
    double* d_psum = (double*) ctx->d_alloc(sizeof(double) * kStatPartialsPerBuf);
    float* d_pmax = (float*) ctx->d_alloc(sizeof(float) * kStatPartialsPerBuf);
    double* d_sum = (double*) ctx->d_alloc(sizeof(double));
    float* d_peak = (float*) ctx->d_alloc(sizeof(float));
    F9_TRY_CUDA(ctx, launch_stats(d_bufs, 1, d_psum, d_pmax, d_sum, d_peak, ctx->stream, &ctx->launches, (ctx->diag.has("F9_RMS_TREE_SUM") ? -1 : ctx->diag.get("F9_RMS_FORCE_ORDER"))));
    double* h_sum = (double*) ctx->h_alloc(sizeof(double));
    float* h_peak = (float*) ctx->h_alloc(sizeof(float));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_peak, d_peak, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    *sumsq = *h_sum; *peak = *h_peak;
    return F9_OK;
}

int f9_calculate_rms(f9_context* ctx, const float* const* ch, int numCh, int numFrames, float* out_rms) {
    int rc = check_planar(ctx, ch, numCh, numFrames); if (rc) return rc;
    if (!out_rms) return ctx->fail(F9_ERR_INVALID, "null out_rms");
    const long long total = (long long) numCh * numFrames;
    if (total == 0) { *out_rms = 0.0f; return F9_OK; }          // MainComponent.cpp:1000-1001
    double s; float pk;
    rc = stats_one(ctx, ch, numCh, numFrames, &s, &pk); if (rc) return rc;
    *out_rms = (float) std::sqrt(s / (double) total);
    return F9_OK;
}

int f9_calculate_noise_floor_db(f9_context* ctx, const float* const* ch, int numCh, int numFrames, float* out_db) {
    float rms = 0.0f;
    int rc = f9_calculate_rms(ctx, ch, numCh, numFrames, &rms); if (rc) return rc;
    if (!out_db) return ctx->fail(F9_ERR_INVALID, "null out_db");
    *out_db = noise_floor_db_from_rms(rms);
    return F9_OK;
}

int f9_tail_scan(f9_context* ctx, const float* const* ch, int numCh, long long numFrames, long long start_frame,
                 int window, int hop, int required, int mode, int has_nf, float nf_db, float margin_pct,
                 long long* out_stop_frame, int* flags, int max_flags, int* out_polls) {
    int rc = check_planar(ctx, ch, numCh, numFrames); if (rc) return rc;
    if (window <= 0 || hop <= 0 || required <= 0 || start_frame < 0 || (mode != F9_TAIL_RMS && mode != F9_TAIL_PEAK) || !out_stop_frame)
        return ctx->fail(F9_ERR_INVALID, "bad tail-scan parameters");
    long long polls = (numFrames - start_frame) / hop;
    if (polls < 0) polls = 0;
    if (out_polls) *out_polls = (int) polls;
    if (polls == 0 || numCh == 0) { *out_stop_frame = -1; return F9_OK; }
    rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + sizeof(int) * (size_t) polls + 16384, sizeof(int) * (size_t) polls + 4096);
    if (rc) return rc;
    DevBuf hb{};
    rc = upload_planar(ctx, ch, numCh, numFrames, &hb); if (rc) return rc;
    TailParams P = make_tail_params(start_frame, window, hop, required, mode, has_nf, nf_db, margin_pct);
    DevBuf* d_bufs; TailParams* d_params;
    rc = upload_array(ctx, &hb, 1, &d_bufs); if (rc) return rc;
    rc = upload_array(ctx, &P, 1, &d_params); if (rc) return rc;
    int* d_flags = (int*) ctx->d_alloc(sizeof(int) * (size_t) polls);
    long long* d_stop = (long long*) ctx->d_alloc(sizeof(long long));
    F9_TRY_CUDA(ctx, launch_tail_scan(d_bufs, d_params, 1, (int) polls, d_stop, d_flags, ctx->stream, &ctx->launches));
    long long* h_stop = (long long*) ctx->h_alloc(sizeof(long long));
    int* h_flags = (int*) ctx->h_alloc(sizeof(int) * (size_t) polls);
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_stop, d_stop, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * (size_t) polls, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    *out_stop_frame = *h_stop;
    if (flags) for (long long i = 0; i < polls && i < max_flags; ++i) flags[i] = h_flags[i];
    return F9_OK;
}

int f9_is_reverb_tail_below_noise_floor(f9_context* ctx, const float* const* ch, int numCh, int numFrames,
                                        int has_nf, float nf_db, float margin_pct, int* out_below) {
    if (!ctx) return F9_ERR_INVALID;
    if (!out_below) return ctx->fail(F9_ERR_INVALID, "null out_below");
    if (numCh <= 0 || numFrames <= 0) {
        // calculateRMS returns 0 for an empty buffer (:1000): windowDb = 20*log10f(1e-10f) = -200
        *out_below = (-200.0f < nf_threshold_db(has_nf, nf_db, margin_pct)) ? 1 : 0;
        return F9_OK;
    }
    long long stop = -1; int flag = -2, polls = 0;
    int rc = f9_tail_scan(ctx, ch, numCh, numFrames, 0, numFrames, numFrames, 1, F9_TAIL_RMS, has_nf, nf_db, margin_pct, &stop, &flag, 1, &polls);
    if (rc) return rc;
    *out_below = (flag == 1) ? 1 : 0;
    return F9_OK;
}

int f9_is_reverb_tail_below_noise_floor_swift(f9_context* ctx, const float* window, long long n,
                                              int has_nf, float nf_db, float margin_pct, int* out_below) {
    if (!ctx) return F9_ERR_INVALID;
    if (!out_below || n < 0 || n > 0x7fffffffLL || (n > 0 && !window)) return ctx->fail(F9_ERR_INVALID, "bad window");
    if (n == 0) {   // Swift: max() of empty = nil -> 0
        if (!has_nf) { *out_below = 1; return F9_OK; }
        *out_below = (-160.0f < nf_db + (nf_db * margin_pct / 100.0f)) ? 1 : 0;
        return F9_OK;
    }
    const float* chans[1] = {window};
    long long stop = -1; int flag = -2, polls = 0;
    int rc = f9_tail_scan(ctx, chans, 1, n, 0, (int) n, (int) n, 1, F9_TAIL_PEAK, has_nf, nf_db, margin_pct, &stop, &flag, 1, &polls);
    if (rc) return rc;
    *out_below = (flag == 1) ? 1 : 0;
    return F9_OK;
}

int f9_trim_latency(f9_context* ctx, const float* const* captured, int numCh, int captured_frames,
                    int latency_samples, int original_length, float* const* out, int* out_copied) {
    int rc = check_planar(ctx, captured, numCh, captured_frames); if (rc) return rc;
    if (numCh <= 0 || original_length < 0 || (original_length > 0 && !out)) return ctx->fail(F9_ERR_INVALID, "bad trim arguments");
    const int start = latency_samples / numCh;
    int n = original_length;
    if (start + n > captured_frames) n = std::max(0, captured_frames - start);
    if (start < 0) n = 0;
    if (out_copied) *out_copied = n;
    if (original_length == 0) return F9_OK;
    rc = ctx->arena_reserve(planar_bytes(numCh, captured_frames) + planar_bytes(numCh, original_length) + 16384, 4096); if (rc) return rc;
    DevBuf hc{}, ho{};
    rc = upload_planar(ctx, captured, numCh, captured_frames, &hc); if (rc) return rc;
    const long long ostride = pad_stride(original_length);
    float* d_out = (float*) ctx->d_alloc(sizeof(float) * (size_t) ostride * numCh);
    ho.base = d_out; ho.chStride = ostride; ho.numCh = numCh; ho.numFrames = original_length;
    DevBuf *d_c, *d_o; int* d_lat;
    rc = upload_array(ctx, &hc, 1, &d_c); if (rc) return rc;
    rc = upload_array(ctx, &ho, 1, &d_o); if (rc) return rc;
    rc = upload_array(ctx, &latency_samples, 1, &d_lat); if (rc) return rc;
    F9_TRY_CUDA(ctx, launch_trim(d_c, d_lat, d_o, 1, original_length, numCh, ctx->stream, &ctx->launches));
    for (int c = 0; c < numCh; ++c)
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(out[c], d_out + c * ostride, sizeof(float) * (size_t) original_length, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}

int f9_trim_latency_swift(f9_context* ctx, const float* captured, long long count, long long latency_samples,
                          long long source_frames, int channels, float* out, long long* out_count) {
    if (!ctx) return F9_ERR_INVALID;
    if (count < 0 || count > 0x7fffffffLL || (count > 0 && !captured) || channels <= 0 || source_frames < 0 || !out_count)
        return ctx->fail(F9_ERR_INVALID, "bad trim arguments");
    // AudioProcessingService.swift:681-703: slice [start, min(start+want, count)), or prefix(want) when start >= count.
    const long long want = source_frames * channels;
    long long start = latency_samples, n;
    if (!(start < count)) { start = 0; n = std::max<long long>(0, std::min(want, count)); }
    else n = std::min(start + want, count) - start;
    *out_count = n;
    if (n <= 0) return F9_OK;
    if (start < 0 || n > 0x7fffffffLL) return ctx->fail(F9_ERR_INVALID, "negative latency");
    // a slice of an interleaved stream is a one-channel trim with latency `start` and length n
    const float* chans[1] = {captured};
    float* outs[1] = {out};
    return f9_trim_latency(ctx, chans, 1, (int) count, (int) start, (int) n, outs, nullptr);
}

int f9_remove_dc_offset(f9_context* ctx, float* const* ch, int numCh, int numFrames) {
    return f9_remove_dc_offset_ex(ctx, ch, numCh, numFrames, 1);      // the drop-in for MainComponent::removeDCOffset: its own arithmetic
}
int f9_remove_dc_offset_ex(f9_context* ctx, float* const* ch, int numCh, int numFrames, int reference_order) {
    int rc = check_planar(ctx, ch, numCh, numFrames); if (rc) return rc;
    if (numCh == 0 || numFrames == 0) return F9_OK;
    rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + sizeof(double) * (size_t) numCh * kDcPartials + 16384, 4096); if (rc) return rc;
    DevBuf hb{};
    rc = upload_planar(ctx, ch, numCh, numFrames, &hb); if (rc) return rc;
    DevBuf* d_b;
    rc = upload_array(ctx, &hb, 1, &d_b); if (rc) return rc;
    double* d_sums = (double*) ctx->d_alloc(sizeof(double) * (size_t) numCh * kDcPartials);
    F9_TRY_CUDA(ctx, launch_remove_dc(d_b, 1, numCh, numFrames, d_sums, ctx->stream, &ctx->launches, reference_order ? 2 : 1));
    for (int c = 0; c < numCh; ++c)
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(ch[c], hb.base + c * hb.chStride, sizeof(float) * (size_t) numFrames, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}

// ------------------------------------------------------------------------------------------------ stimuli
// MainComponent::generateImpulse, Source/MainComponent.cpp:934-945: buffer.clear(), 0.9 on sample 0 of every channel.
int f9_generate_impulse(f9_context* ctx, float* const* ch, int numCh, int numFrames) {
    int rc = check_planar(ctx, ch, numCh, numFrames); if (rc) return rc;
    if (numCh == 0 || numFrames == 0) return F9_OK;
    rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + 16384, 4096); if (rc) return rc;
    const long long stride = pad_stride(numFrames);
    DevBuf hb{(float*) ctx->d_alloc(sizeof(float) * (size_t) stride * numCh), stride, numCh, numFrames};
    DevBuf* d_b;
    rc = upload_array(ctx, &hb, 1, &d_b); if (rc) return rc;
    F9_TRY_CUDA(ctx, launch_impulse(d_b, 1, numCh, numFrames, 0.9f, ctx->stream, &ctx->launches));
    for (int c = 0; c < numCh; ++c)
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(ch[c], hb.base + c * stride, sizeof(float) * (size_t) numFrames, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}
int f9_dev_generate_impulse(f9_context* ctx, const f9_dev_buffer* bufs, int n) {
    if (!ctx || n < 0 || (n > 0 && !bufs)) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(bufs);
    int maxCh = 0, maxFrames = 0;
    for (int i = 0; i < n; ++i) {
        if (hb[i].numCh < 0 || hb[i].numFrames < 0 || (hb[i].numCh > 0 && hb[i].numFrames > 0 && !hb[i].base)) return ctx->fail(F9_ERR_INVALID, "bad buffer");
        maxCh = std::max(maxCh, hb[i].numCh); maxFrames = std::max(maxFrames, hb[i].numFrames);
    }
    int rc = ctx->arena_reserve(sizeof(DevBuf) * (size_t) n + 8192, sizeof(DevBuf) * (size_t) n + 8192, true); if (rc) return rc;
    DevBuf* d_b;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    F9_TRY_CUDA(ctx, launch_impulse(d_b, n, maxCh, maxFrames, 0.9f, ctx->stream, &ctx->launches));
    return F9_OK;
}

// MainComponent::generateSineWave, Source/MainComponent.cpp:907-932 (callback_form: the audio callback's variant :141-167, whose
// sinePhase is the chain's own end value instead of the closed-form update of :929-931).
int f9_generate_sine_wave(f9_context* ctx, float* const* ch, int numCh, int numSamples, float frequency, float sample_rate,
                          float amplitude, float* phase_io, int callback_form) {
    int rc = check_planar(ctx, ch, numCh, numSamples); if (rc) return rc;
    if (!phase_io || !(sample_rate > 0.0f)) return ctx->fail(F9_ERR_INVALID, "bad sine arguments");
    const float inc = sine_phase_increment(frequency, sample_rate);
    if (numSamples == 0 || (numCh == 0 && !callback_form)) { if (!callback_form) *phase_io = sine_phase_after_block(*phase_io, inc, numSamples); return F9_OK; }
    rc = ctx->arena_reserve(planar_bytes(numCh, numSamples) + sizeof(float) * ((size_t) numSamples + 1) + 16384, 4096); if (rc) return rc;
    const long long stride = pad_stride(numSamples);
    DevBuf hb{(float*) ctx->d_alloc(sizeof(float) * (size_t) stride * std::max(numCh, 1)), stride, numCh, numSamples};
    float* d_ph = (float*) ctx->d_alloc(sizeof(float) * ((size_t) numSamples + 1));
    F9_TRY_CUDA(ctx, launch_sine(hb, *phase_io, inc, amplitude, numSamples, d_ph, ctx->stream, &ctx->launches));
    for (int c = 0; c < numCh; ++c)
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(ch[c], hb.base + c * stride, sizeof(float) * (size_t) numSamples, cudaMemcpyDeviceToHost, ctx->stream));
    float* h_end = (float*) ctx->h_alloc(sizeof(float));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_end, d_ph + numSamples, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    *phase_io = callback_form ? *h_end : sine_phase_after_block(*phase_io, inc, numSamples);
    return F9_OK;
}
int f9_dev_generate_sine_wave(f9_context* ctx, const f9_dev_buffer* buf, float frequency, float sample_rate, float amplitude, float phase0) {
    if (!ctx || !buf || !(sample_rate > 0.0f) || buf->numCh < 0 || buf->numFrames < 0) return F9_ERR_INVALID;
    if (buf->numCh == 0 || buf->numFrames == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ctx->arena_reserve(sizeof(float) * ((size_t) buf->numFrames + 1) + 8192, 4096, true); if (rc) return rc;
    float* d_ph = (float*) ctx->d_alloc(sizeof(float) * ((size_t) buf->numFrames + 1));
    const DevBuf hb = *reinterpret_cast<const DevBuf*>(buf);
    F9_TRY_CUDA(ctx, launch_sine(hb, phase0, sine_phase_increment(frequency, sample_rate), amplitude, buf->numFrames, d_ph, ctx->stream, &ctx->launches));
    return F9_OK;
}
// Swift SineWaveGenerator.generateSineWave, SineWaveGenerator.swift:35-59: double phase, interleaved, one sample per frame on every channel.
int f9_generate_sine_wave_swift(f9_context* ctx, float* buffer, int frame_count, int channel_count, double frequency, double sample_rate,
                                float amplitude, double* phase_io) {
    if (!ctx) return F9_ERR_INVALID;
    if (frame_count < 0 || channel_count < 0 || !phase_io || !(sample_rate > 0.0) || (frame_count > 0 && channel_count > 0 && !buffer))
        return ctx->fail(F9_ERR_INVALID, "bad sine arguments");
    if (frame_count == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t outBytes = sizeof(float) * (size_t) frame_count * std::max(channel_count, 1);
    int rc = ctx->arena_reserve(outBytes + sizeof(double) * ((size_t) frame_count + 1) + 16384, 4096); if (rc) return rc;
    float* d_out = (float*) ctx->d_alloc(outBytes);
    double* d_ph = (double*) ctx->d_alloc(sizeof(double) * ((size_t) frame_count + 1));
    const double inc = 2.0 * 3.14159265358979323846 * frequency / sample_rate;
    F9_TRY_CUDA(ctx, launch_sine_swift(d_out, channel_count, *phase_io, inc, amplitude, frame_count, d_ph, ctx->stream, &ctx->launches));
    if (channel_count > 0) F9_TRY_CUDA(ctx, cudaMemcpyAsync(buffer, d_out, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
    double* h_end = (double*) ctx->h_alloc(sizeof(double));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_end, d_ph + frame_count, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    *phase_io = *h_end;
    return F9_OK;
}

int f9_xcorr_peak(f9_context* ctx, const float* const* y, int numCh, int numFrames, const float* x, int stim_len,
                  int lag_min, int lag_max, float threshold, int* out_found, int* out_lag, int* out_ch, double* out_value) {
    int rc = check_planar(ctx, y, numCh, numFrames); if (rc) return rc;
    if (stim_len < 0 || (stim_len > 0 && !x) || lag_max < lag_min || !out_found || !out_lag)
        return ctx->fail(F9_ERR_INVALID, "bad xcorr arguments");
    *out_found = 0; *out_lag = 0; if (out_ch) *out_ch = -1; if (out_value) *out_value = 0.0;
    if (numCh == 0 || numFrames == 0 || stim_len == 0) return F9_OK;
    DevBuf hb{}; hb.numCh = numCh; hb.numFrames = numFrames;
    std::vector<int> prefix;
    const int total = xcorr_prefix(&hb, 1, lag_min, lag_max, &prefix);
    const bool fast = !ctx->diag.has("F9_XCORR_EXACT_ALL");
    const size_t fastBytes = fast ? xcorr_fast_scratch_bytes(1, numCh, lag_min, lag_max) : 0;
    rc = ctx->arena_reserve(planar_bytes(numCh, numFrames) + sizeof(float) * (size_t) stim_len + sizeof(XcPartial) * ((size_t) total + 2) + 16384 + fastBytes,
                            sizeof(float) * (size_t) stim_len + 8192);
    if (rc) return rc;
    rc = upload_planar(ctx, y, numCh, numFrames, &hb); if (rc) return rc;
    float* d_stim = (float*) ctx->d_alloc(sizeof(float) * (size_t) stim_len);
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_stim, x, sizeof(float) * (size_t) stim_len, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf* d_b; int* d_prefix;
    rc = upload_array(ctx, &hb, 1, &d_b); if (rc) return rc;
    rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
    XcPartial* d_part = (XcPartial*) ctx->d_alloc(sizeof(XcPartial) * (size_t) total);
    XcPartial* d_best = (XcPartial*) ctx->d_alloc(sizeof(XcPartial));
    if (fast) {
        void* d_scr = ctx->d_alloc(fastBytes);
        if (!d_scr) return ctx->fail(F9_ERR_NOMEM, "xcorr scratch");
        F9_TRY_CUDA(ctx, launch_xcorr_fast(&hb, d_b, 1, total, d_prefix, d_stim, stim_len, lag_min, lag_max, d_part, d_best, d_scr, ctx->stream, &ctx->launches));
    } else
    F9_TRY_CUDA(ctx, launch_xcorr(d_b, 1, total, d_prefix, d_stim, stim_len, lag_min, lag_max, d_part, d_best, ctx->stream, &ctx->launches));
    XcPartial* h_best = (XcPartial*) ctx->h_alloc(sizeof(XcPartial));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_best, d_best, sizeof(XcPartial), cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    double energy = 0.0;                    // ||x||_2 of the (short) stimulus: host scalar
    for (int i = 0; i < stim_len; ++i) energy += (double) x[i] * (double) x[i];
    const double norm = std::sqrt(energy);
    *out_lag = h_best->lag;
    if (out_ch) *out_ch = h_best->ch;
    if (out_value) *out_value = h_best->v;
    *out_found = (h_best->ch >= 0 && h_best->v > (double) threshold * norm) ? 1 : 0;
    return F9_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ resample plumbing
struct f9_interp {
    f9_context* ctx = nullptr;
    int kind = 0;
    int memory = 0;
    std::vector<float> hist;        // last `memory` inputs, oldest first (GenericInterpolator::lastInputSamples unrolled)
    double pos = 1.0;               // subSamplePos
};

struct f9_plan {
    f9_context* ctx = nullptr;
    std::vector<f9_plan*> parts;       // segments of mixed 16-byte alignment: one sub-plan per alignment class (this plan then launches nothing itself)
    ResampleLaunch L;
    Seg* d_segs = nullptr;
    int* d_prefix = nullptr;           // n_segs + 1 tile counts, then the plan's own overflow flag
    void* d_scratch = nullptr;
};

namespace {

// Core of process()/processAdding(): `lin` holds the inputs this call consumes, in order (n_used of them).
int interp_run(f9_interp* h, double ratio, const float* lin, int n_used, float* out, int num_out, bool adding, float gain, double new_pos) {
    f9_context* ctx = h->ctx;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const int M = h->memory;
    const size_t n_in = (size_t) M + (size_t) n_used + 1;       // history + fresh inputs + one zero of slack
    int rc = ctx->arena_reserve(sizeof(float) * (n_in + (size_t) num_out) + 65536, sizeof(float) * n_in + 65536);
    if (rc) return rc;
    float* h_in = (float*) ctx->h_alloc(sizeof(float) * n_in);
    std::memcpy(h_in, h->hist.data(), sizeof(float) * (size_t) M);
    if (n_used) std::memcpy(h_in + M, lin, sizeof(float) * (size_t) n_used);
    h_in[n_in - 1] = 0.0f;
    float* d_in = (float*) ctx->d_alloc(sizeof(float) * n_in);
    float* d_out = (float*) ctx->d_alloc(sizeof(float) * (size_t) num_out);
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, sizeof(float) * n_in, cudaMemcpyHostToDevice, ctx->stream));
    if (adding) F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_out, out, sizeof(float) * (size_t) num_out, cudaMemcpyHostToDevice, ctx->stream));

    ResampleLaunch L;
    rc = ctx->prepare_resample(h->kind, ratio, h->pos, /*allow_rational=*/false, &L); if (rc) return rc;
    Seg seg{};
    seg.in = d_in; seg.inOffset = -(long long) M; seg.inAvail = (long long) n_in;
    seg.out = d_out; seg.n0 = 0; seg.numOut = num_out;
    std::vector<int> prefix;
    const int tiles = resample_build_tiles(L, &seg, 1, &prefix);
    Seg* d_seg; int* d_prefix;
    rc = upload_array(ctx, &seg, 1, &d_seg); if (rc) return rc;
    rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
    L.d_segs = d_seg; L.d_tile_prefix = d_prefix; L.n_segs = 1; L.n_tiles = tiles;
    if (resample_needs_ovf(L)) L.d_ovf = (unsigned*) ctx->d_alloc(sizeof(unsigned));
    L.adding = adding ? 1 : 0; L.gain = gain;
    F9_TRY_CUDA(ctx, launch_resample(L, ctx->stream, &ctx->launches));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t) num_out, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);

    // new history = last M of (history ++ consumed inputs); new position from the exact recurrence
    if (n_used >= M) std::memcpy(h->hist.data(), lin + (n_used - M), sizeof(float) * (size_t) M);
    else if (n_used > 0) {
        std::memmove(h->hist.data(), h->hist.data() + n_used, sizeof(float) * (size_t) (M - n_used));
        std::memcpy(h->hist.data() + (M - n_used), lin, sizeof(float) * (size_t) n_used);
    }
    h->pos = new_pos;
    return F9_OK;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ D. interpolators
int f9_interp_create(f9_context* ctx, int kind, f9_interp** out) {
    if (!ctx || !out) return F9_ERR_INVALID;
    const int M = interp_memory(kind);
    if (M == 0) return ctx->fail(F9_ERR_INVALID, "unknown interpolator kind");
    f9_interp* h = new (std::nothrow) f9_interp();
    if (!h) return F9_ERR_NOMEM;
    h->ctx = ctx; h->kind = kind; h->memory = M;
    h->hist.assign((size_t) M, 0.0f);
    h->pos = 1.0;
    *out = h;
    return F9_OK;
}
void f9_interp_destroy(f9_interp* h) { delete h; }
int f9_interp_reset(f9_interp* h) {
    if (!h) return F9_ERR_INVALID;
    std::fill(h->hist.begin(), h->hist.end(), 0.0f);
    h->pos = 1.0;
    return F9_OK;
}
float f9_interp_base_latency(const f9_interp* h) { return h ? interp_latency(h->kind) : 0.0f; }

static int interp_process_impl(f9_interp* h, double ratio, const float* in, float* out, int num_out, bool adding, float gain) {
    if (!h) return F9_ERR_INVALID;
    if (num_out < 0 || (num_out > 0 && (!in || !out)) || !(ratio > 0.0) || !std::isfinite(ratio))
        return h->ctx->fail(F9_ERR_INVALID, "bad process() arguments");
    if (num_out == 0) return 0;
    double pos = h->pos;
    const int used = run_position_chain(&pos, ratio, num_out);
    int rc = interp_run(h, ratio, in, used, out, num_out, adding, gain, pos);
    return rc ? rc : used;
}
int f9_interp_process(f9_interp* h, double speed_ratio, const float* in, float* out, int num_out) {
    return interp_process_impl(h, speed_ratio, in, out, num_out, false, 1.0f);
}
int f9_interp_process_adding(f9_interp* h, double speed_ratio, const float* in, float* out, int num_out, float gain) {
    return interp_process_impl(h, speed_ratio, in, out, num_out, true, gain);
}
int f9_interp_process_wrap(f9_interp* h, double speed_ratio, const float* in, float* out, int num_out,
                           int num_in_available, int wrap_around) {
    if (!h) return F9_ERR_INVALID;
    if (num_out < 0 || (num_out > 0 && (!in || !out)) || !(speed_ratio > 0.0) || !std::isfinite(speed_ratio) || wrap_around < 0)
        return h->ctx->fail(F9_ERR_INVALID, "bad process() arguments");
    if (num_out == 0) return 0;
    // Linearise the input exactly as the 6-argument interpolate() walks it: wrap the read pointer back by
    // `wrap_around` when the available count runs out, or push zeros once exceeded.
    double pos = h->pos;
    const int total = run_position_chain(&pos, speed_ratio, num_out);
    std::vector<float> lin((size_t) total);
    long long rd = 0; int avail = num_in_available; bool exceeded = false;
    for (int i = 0; i < total; ++i) {
        if (exceeded) lin[(size_t) i] = 0.0f;
        else {
            lin[(size_t) i] = in[rd++];
            if (--avail <= 0) {
                if (wrap_around > 0) { rd -= wrap_around; avail += wrap_around; }
                else exceeded = true;
            }
        }
    }
    int rc = interp_run(h, speed_ratio, lin.data(), total, out, num_out, false, 1.0f, pos);
    if (rc) return rc;
    if (wrap_around == 0) return (int) rd;
    return (int) ((rd + wrap_around) % wrap_around);
}

int f9_sinc_table_set(f9_context* ctx, const float* table10001) {
    if (!ctx || !table10001) return F9_ERR_INVALID;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    std::memcpy(ctx->sinc_table.data(), table10001, sizeof(float) * kSincTableSize);
    ctx->sinc_table[kSincTableSize] = 0.0f;
    F9_FINISH(ctx);
    F9_TRY_CUDA(ctx, cudaMemcpy(ctx->d_sinc_table, ctx->sinc_table.data(), sizeof(float) * (kSincTableSize + 1), cudaMemcpyHostToDevice));
    ++ctx->sinc_epoch;          // polyphase tables built from the old table are no longer looked up
    return F9_OK;
}
int f9_sinc_table_get(const f9_context* ctx, float* table10001) {
    if (!table10001) return F9_ERR_INVALID;
    if (ctx) std::memcpy(table10001, ctx->sinc_table.data(), sizeof(float) * kSincTableSize);
    else make_default_sinc_table(table10001);          // no context (no GPU): the built-in default
    return F9_OK;
}

// ------------------------------------------------------------------------------------------------ F. device-resident
int f9_dev_find_peak_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, float threshold, int* d_out_pos) {
    if (!ctx || n < 0 || (n > 0 && (!bufs || !d_out_pos))) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(f9_dev_buffer) == sizeof(DevBuf), "layout");
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(bufs);
    std::vector<int> prefix;
    const int total = peak_prefix(hb, n, &prefix);
    int rc = ctx->arena_reserve(sizeof(DevBuf) * (size_t) n + sizeof(int) * (size_t) (n + 1) + sizeof(PeakPartial) * (size_t) total + 16384,
                                sizeof(DevBuf) * (size_t) n + sizeof(int) * (size_t) (n + 1) + 8192, true);
    if (rc) return rc;
    DevBuf* d_b; int* d_prefix;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
    PeakPartial* d_part = (PeakPartial*) ctx->d_alloc(sizeof(PeakPartial) * (size_t) std::max(total, 1));
    F9_TRY_CUDA(ctx, launch_find_peak(d_b, n, total, d_prefix, threshold, d_part, d_out_pos, ctx->stream, &ctx->launches));
    return F9_OK;
}

int f9_dev_latency_stats_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, float threshold, int* d_out_pos, double* d_sumsq, float* d_peak) {
    if (!ctx || n < 0 || (n > 0 && (!bufs || !d_out_pos || !d_sumsq))) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(bufs);
    std::vector<int> prefix;
    const int total = peak_prefix(hb, n, &prefix);
    int rc = ctx->arena_reserve(sizeof(DevBuf) * (size_t) n + sizeof(int) * (size_t) (n + 1) + (sizeof(PeakPartial) + sizeof(double)) * (size_t) total + 16384,
                                sizeof(DevBuf) * (size_t) n + sizeof(int) * (size_t) (n + 1) + 8192, true);
    if (rc) return rc;
    DevBuf* d_b; int* d_prefix;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
    PeakPartial* d_part = (PeakPartial*) ctx->d_alloc(sizeof(PeakPartial) * (size_t) std::max(total, 1));
    double* d_psum = (double*) ctx->d_alloc(sizeof(double) * (size_t) std::max(total, 1));
    F9_TRY_CUDA(ctx, launch_find_peak(d_b, n, total, d_prefix, threshold, d_part, d_out_pos, ctx->stream, &ctx->launches, d_psum, d_sumsq, d_peak, (ctx->diag.has("F9_RMS_TREE_SUM") ? -1 : ctx->diag.get("F9_RMS_FORCE_ORDER"))));
    return F9_OK;
}

int f9_dev_stats_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, double* d_sumsq, float* d_peak) {
    if (!ctx || n < 0 || (n > 0 && (!bufs || !d_sumsq || !d_peak))) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(bufs);
    int rc = ctx->arena_reserve((sizeof(DevBuf) + 12 * kStatPartialsPerBuf) * (size_t) n + 16384, sizeof(DevBuf) * (size_t) n + 8192, true);
    if (rc) return rc;
    DevBuf* d_b;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    double* d_psum = (double*) ctx->d_alloc(sizeof(double) * kStatPartialsPerBuf * (size_t) n);
    float* d_pmax = (float*) ctx->d_alloc(sizeof(float) * kStatPartialsPerBuf * (size_t) n);
    F9_TRY_CUDA(ctx, launch_stats(d_b, n, d_psum, d_pmax, d_sumsq, d_peak, ctx->stream, &ctx->launches, (ctx->diag.has("F9_RMS_TREE_SUM") ? -1 : ctx->diag.get("F9_RMS_FORCE_ORDER"))));
    return F9_OK;
}

int f9_dev_xcorr_peak_batch(f9_context* ctx, const f9_dev_buffer* bufs, int n, const float* d_stim,
                            int stim_len, int lag_min, int lag_max, f9_xcorr_result* d_out) {
    if (!ctx || n < 0 || (n > 0 && (!bufs || !d_out || !d_stim)) || stim_len <= 0 || lag_max < lag_min) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(f9_xcorr_result) == sizeof(XcPartial), "layout");
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(bufs);
    std::vector<int> prefix;
    const int total = xcorr_prefix(hb, n, lag_min, lag_max, &prefix);
    const size_t small = sizeof(DevBuf) * (size_t) n + sizeof(int) * (size_t) (n + 1) + 8192;
    if (ctx->diag.has("F9_XCORR_EXACT_ALL")) {                  // every lag in exact double sums (the round-1 path; tests compare the two)
        int rc = ctx->arena_reserve(small + sizeof(XcPartial) * (size_t) std::max(total, 1) + 16384, small, true);
        if (rc) return rc;
        DevBuf* d_b; int* d_prefix;
        rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
        rc = upload_array(ctx, prefix.data(), prefix.size(), &d_prefix); if (rc) return rc;
        XcPartial* d_part = (XcPartial*) ctx->d_alloc(sizeof(XcPartial) * (size_t) std::max(total, 1));
        F9_TRY_CUDA(ctx, launch_xcorr(d_b, n, total, d_prefix, d_stim, stim_len, lag_min, lag_max, d_part,
                                      reinterpret_cast<XcPartial*>(d_out), ctx->stream, &ctx->launches));
        return F9_OK;
    }
    // candidates on the tensor cores + exact verification, in groups of buffers whose approximate correlations fit ~1 GB of scratch
    int maxCh = 1;
    for (int i = 0; i < n; ++i) maxCh = std::max(maxCh, hb[i].numCh);
    const size_t perBuf = xcorr_fast_scratch_bytes(1, maxCh, lag_min, lag_max);
    const int group = (int) std::max<size_t>(1, std::min<size_t>((size_t) n, (size_t(1) << 30) / perBuf));
    for (int b0 = 0; b0 < n; b0 += group) {
        const int nb = std::min(group, n - b0);
        std::vector<int> pre((size_t) nb + 1);
        for (int i = 0; i <= nb; ++i) pre[(size_t) i] = prefix[(size_t) (b0 + i)] - prefix[(size_t) b0];
        const int tot = pre[(size_t) nb];
        const size_t scr = xcorr_fast_scratch_bytes(nb, maxCh, lag_min, lag_max);
        int rc = ctx->arena_reserve(small + sizeof(XcPartial) * (size_t) std::max(tot, 1) + scr + 16384, small, true);
        if (rc) return rc;
        DevBuf* d_b; int* d_prefix;
        rc = upload_array(ctx, hb + b0, (size_t) nb, &d_b); if (rc) return rc;
        rc = upload_array(ctx, pre.data(), pre.size(), &d_prefix); if (rc) return rc;
        XcPartial* d_part = (XcPartial*) ctx->d_alloc(sizeof(XcPartial) * (size_t) std::max(tot, 1));
        void* d_scr = ctx->d_alloc(scr);
        if (!d_part || !d_scr) return ctx->fail(F9_ERR_NOMEM, "xcorr scratch");
        F9_TRY_CUDA(ctx, launch_xcorr_fast(hb + b0, d_b, nb, tot, d_prefix, d_stim, stim_len, lag_min, lag_max, d_part,
                                           reinterpret_cast<XcPartial*>(d_out) + b0, d_scr, ctx->stream, &ctx->launches));
    }
    return F9_OK;
}

int f9_dev_tail_scan_batch(f9_context* ctx, const f9_dev_buffer* bufs, const f9_tail_params* params, int n,
                           long long* d_stop_frame, int* d_flags, int max_polls) {
    if (!ctx || n < 0 || (n > 0 && (!bufs || !params || !d_stop_frame)) || max_polls < 0 || (max_polls > 0 && !d_flags)) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(bufs);
    std::vector<TailParams> tp((size_t) n);
    for (int i = 0; i < n; ++i) {
        const f9_tail_params& p = params[i];
        if (p.window <= 0 || p.hop <= 0 || p.required <= 0 || p.start_frame < 0) return ctx->fail(F9_ERR_INVALID, "bad tail-scan parameters");
        tp[(size_t) i] = make_tail_params(p.start_frame, p.window, p.hop, p.required, p.mode, p.has_nf, p.nf_db, p.margin_pct);
    }
    int rc = ctx->arena_reserve((sizeof(DevBuf) + sizeof(TailParams)) * (size_t) n + 16384, (sizeof(DevBuf) + sizeof(TailParams)) * (size_t) n + 8192, true);
    if (rc) return rc;
    DevBuf* d_b; TailParams* d_p;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    rc = upload_array(ctx, tp.data(), tp.size(), &d_p); if (rc) return rc;
    F9_TRY_CUDA(ctx, launch_tail_scan(d_b, d_p, n, max_polls, d_stop_frame, d_flags, ctx->stream, &ctx->launches));
    return F9_OK;
}

int f9_dev_trim_batch(f9_context* ctx, const f9_dev_buffer* captured, const int* latency_samples,
                      const f9_dev_buffer* out, int n, int remove_dc) {
    if (!ctx || n < 0 || (n > 0 && (!captured || !latency_samples || !out))) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hc = reinterpret_cast<const DevBuf*>(captured);
    const DevBuf* ho = reinterpret_cast<const DevBuf*>(out);
    int maxCh = 0, maxFrames = 0;
    for (int i = 0; i < n; ++i) {
        if (hc[i].numCh <= 0 || ho[i].numCh != hc[i].numCh) return ctx->fail(F9_ERR_INVALID, "trim: channel counts differ");
        maxCh = std::max(maxCh, ho[i].numCh); maxFrames = std::max(maxFrames, ho[i].numFrames);
    }
    const size_t bytes = (2 * sizeof(DevBuf) + sizeof(int)) * (size_t) n + sizeof(double) * (size_t) n * maxCh * kDcPartials + 16384;
    int rc = ctx->arena_reserve(bytes, bytes, true); if (rc) return rc;
    DevBuf *d_c, *d_o; int* d_lat;
    rc = upload_array(ctx, hc, (size_t) n, &d_c); if (rc) return rc;
    rc = upload_array(ctx, ho, (size_t) n, &d_o); if (rc) return rc;
    rc = upload_array(ctx, latency_samples, (size_t) n, &d_lat); if (rc) return rc;
    double* d_part = remove_dc ? (double*) ctx->d_alloc(sizeof(double) * (size_t) n * maxCh * kDcPartials) : nullptr;      // fused removeDCOffset
    F9_TRY_CUDA(ctx, launch_trim(d_c, d_lat, d_o, n, maxFrames, maxCh, ctx->stream, &ctx->launches, d_part, nullptr, remove_dc == 2 ? 2 : 1));
    return F9_OK;
}

int f9_resample_plan_create(f9_context* ctx, int kind, double speed_ratio, const f9_resample_seg* segs, int n_segs, f9_plan** out) {
    if (!ctx || !out || n_segs < 0 || (n_segs > 0 && !segs)) return F9_ERR_INVALID;
    *out = nullptr;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(f9_resample_seg) == sizeof(Seg), "layout");
    f9_plan* P = new (std::nothrow) f9_plan();
    if (!P) return F9_ERR_NOMEM;
    P->ctx = ctx;
    int rc = ctx->prepare_resample(kind, speed_ratio, 1.0, true, &P->L);
    if (rc) { delete P; return rc; }
    const Seg* hs = reinterpret_cast<const Seg*>(segs);
    for (int i = 0; i < n_segs; ++i)
        if (hs[i].numOut < 0 || hs[i].n0 < 0 || hs[i].inAvail < 0 || (hs[i].numOut > 0 && (!hs[i].out || (hs[i].inAvail > 0 && !hs[i].in)))) {
            delete P; return ctx->fail(F9_ERR_INVALID, "bad resample segment");
        }
    // Tensor-core plans feed through TMA when every segment's first sample sits on a 16-byte boundary or all of them sit the same
    // 1-3 floats past one (resample_build_tiles: shifted tables).  Segments of MIXED alignment are planned per alignment class,
    // each class on the fast feed, instead of sending the whole plan to the register loader.
    if (P->L.umma && (P->L.um.p & 3) == 0 && !ctx->diag.has("F9_UMMA_NOSHIFT")) {
        std::vector<Seg> cls[4]; bool all4 = true;
        for (int i = 0; i < n_segs; ++i) {
            if (hs[i].numOut <= 0) continue;
            if (reinterpret_cast<uintptr_t>(hs[i].in) & 3) { all4 = false; break; }
            cls[(int) (((long long) (reinterpret_cast<uintptr_t>(hs[i].in) >> 2) - hs[i].inOffset) & 3)].push_back(hs[i]);
        }
        int used = 0;
        for (int c = 0; c < 4; ++c) used += cls[c].empty() ? 0 : 1;
        if (all4 && used > 1) {
            for (int c = 0; c < 4; ++c) {
                if (cls[c].empty()) continue;
                f9_plan* sub = nullptr;
                rc = f9_resample_plan_create(ctx, kind, speed_ratio, reinterpret_cast<const f9_resample_seg*>(cls[c].data()), (int) cls[c].size(), &sub);
                if (rc) { f9_plan_destroy(P); return rc; }
                P->parts.push_back(sub);
            }
            *out = P;
            return F9_OK;
        }
    }
    std::vector<int> prefix;
    const int tiles = resample_build_tiles(P->L, hs, n_segs, &prefix);
    if (tiles < 0) { delete P; return ctx->fail(F9_ERR_INVALID, "too many tiles"); }
    cudaError_t e;
    if ((e = cudaMalloc((void**) &P->d_segs, sizeof(Seg) * (size_t) std::max(n_segs, 1))) != cudaSuccess ||
        (e = cudaMalloc((void**) &P->d_prefix, sizeof(int) * (size_t) (n_segs + 2))) != cudaSuccess ||
        (n_segs > 0 && (e = cudaMemcpy(P->d_segs, hs, sizeof(Seg) * (size_t) n_segs, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (e = cudaMemcpy(P->d_prefix, prefix.data(), sizeof(int) * (size_t) (n_segs + 1), cudaMemcpyHostToDevice)) != cudaSuccess) {
        f9_plan_destroy(P);
        return ctx->fail_cuda(e, "plan upload");
    }
    P->L.d_segs = P->d_segs; P->L.d_tile_prefix = P->d_prefix; P->L.n_segs = n_segs; P->L.n_tiles = tiles;
    P->L.d_ovf = reinterpret_cast<unsigned*>(P->d_prefix + (n_segs + 1));
    if (const size_t sb = resample_scratch_bytes(P->L, tiles)) {
        if ((e = cudaMalloc(&P->d_scratch, sb)) != cudaSuccess) { f9_plan_destroy(P); return ctx->fail_cuda(e, "plan scratch"); }
        P->L.d_tile_recs = (UmmaTileRec*) P->d_scratch;
    }
    *out = P;
    return F9_OK;
}
int f9_resample_plan_run(f9_plan* plan) {
    if (!plan) return F9_ERR_INVALID;
    f9_context* ctx = plan->ctx;
    if (!plan->parts.empty()) {
        for (f9_plan* sub : plan->parts) { const int rc = f9_resample_plan_run(sub); if (rc) return rc; }
        return F9_OK;
    }
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    if (plan->L.recs_stream != ctx->stream) plan->L.recs_ready = false;
    F9_TRY_CUDA(ctx, launch_resample(plan->L, ctx->stream, &ctx->launches));
    plan->L.recs_ready = plan->L.d_tile_recs != nullptr;       // the segments of a plan never change: its tile records are built once
    plan->L.recs_stream = ctx->stream;
    return F9_OK;
}
void f9_plan_destroy(f9_plan* plan) {
    if (!plan) return;
    for (f9_plan* sub : plan->parts) f9_plan_destroy(sub);
    plan->parts.clear();
    cudaSetDevice(plan->ctx->device);
    cudaStreamSynchronize(plan->ctx->stream);
    if (plan->d_segs) cudaFree(plan->d_segs);
    if (plan->d_prefix) cudaFree(plan->d_prefix);
    if (plan->d_scratch) cudaFree(plan->d_scratch);
    delete plan;
}

int f9_resample_segment_input_range(int kind, double speed_ratio, long long n0, long long num_out, long long* first_in, long long* last_in_plus1) {
    const int taps = interp_memory(kind);
    if (taps == 0 || !(speed_ratio > 0.0) || n0 < 0 || num_out <= 0 || !first_in || !last_in_plus1) return F9_ERR_INVALID;
    long long p, q, c0, c1;
    if (find_rational(speed_ratio, 4096, &p, &q)) {
        // exact: newest input of output n is floor(n*p/q)
        c0 = (long long) (((__int128) n0 * p) / q) + 1;
        c1 = (long long) (((__int128) (n0 + num_out - 1) * p) / q) + 1;
    } else {
        position_closed_form(1.0, speed_ratio, n0, &c0, nullptr);
        position_closed_form(1.0, speed_ratio, n0 + num_out - 1, &c1, nullptr);
        c0 -= 1; c1 += 1;                       // the generic kernel stages one sample of slack
    }
    *first_in = (c0 - 1) - (taps - 1);
    *last_in_plus1 = c1;
    return F9_OK;
}

long long f9_resampled_length(long long n_in, double fs_in, double fs_out) {
    if (n_in <= 0 || !(fs_in > 0.0) || !(fs_out > 0.0)) return 0;
    if (fs_in == fs_out) return n_in;
    long long p, q;
    if (find_rational(fs_in / fs_out, 4096, &p, &q)) return (long long) (((__int128) n_in * q + p - 1) / p);
    return (long long) std::ceil((double) n_in * fs_out / fs_in);
}

// ------------------------------------------------------------------------------------------------ G. format convert
int f9_dev_pcm_to_planar(f9_context* ctx, const void* d_src, int fmt, int src_ch, long long num_frames,
                         float* d_dst, long long dst_ch_stride, int dst_ch) {
    if (!ctx || fmt < F9_PCM_U8 || fmt > F9_PCM_F32LE || src_ch <= 0 || dst_ch <= 0 || num_frames < 0) return F9_ERR_INVALID;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    F9_TRY_CUDA(ctx, launch_pcm_to_planar(d_src, fmt, src_ch, num_frames, d_dst, dst_ch_stride, dst_ch, ctx->stream, &ctx->launches));
    return F9_OK;
}
int f9_dev_planar_to_pcm24(f9_context* ctx, const float* d_src, long long src_ch_stride, int numCh, long long num_frames, unsigned char* d_dst) {
    if (!ctx || numCh <= 0 || num_frames < 0) return F9_ERR_INVALID;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    F9_TRY_CUDA(ctx, launch_planar_to_pcm24(d_src, src_ch_stride, numCh, num_frames, d_dst, ctx->stream, &ctx->launches));
    return F9_OK;
}

int f9_dev_pcm_to_planar_batch(f9_context* ctx, const void* const* d_srcs, int fmt, int src_ch, const f9_dev_buffer* dst, int n) {
    if (!ctx || fmt < F9_PCM_U8 || fmt > F9_PCM_F32LE || src_ch <= 0 || n < 0 || (n > 0 && (!d_srcs || !dst))) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    for (int i = 0; i < n; ++i)
        if (dst[i].numCh <= 0 || dst[i].numFrames < 0 || (dst[i].numFrames > 0 && (!d_srcs[i] || !dst[i].base))) return ctx->fail(F9_ERR_INVALID, "bad file descriptor in the batch");
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(dst);
    int rc = ctx->arena_reserve((sizeof(DevBuf) + sizeof(void*)) * (size_t) n + 16384, (sizeof(DevBuf) + sizeof(void*)) * (size_t) n + 8192, true);
    if (rc) return rc;
    DevBuf* d_b; const unsigned char** d_p;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    rc = upload_array(ctx, reinterpret_cast<const unsigned char* const*>(d_srcs), (size_t) n, &d_p); if (rc) return rc;
    F9_TRY_CUDA(ctx, launch_pcm_to_planar_batch(d_p, fmt, src_ch, hb, d_b, n, ctx->stream, &ctx->launches,
                                                ctx->diag.has("F9_PCM_BYTEWISE") ? nullptr : reinterpret_cast<const unsigned char* const*>(d_srcs)));
    return F9_OK;
}
int f9_dev_planar_to_pcm24_batch(f9_context* ctx, const f9_dev_buffer* src, unsigned char* const* d_dsts, int n) {
    if (!ctx || n < 0 || (n > 0 && (!src || !d_dsts))) return F9_ERR_INVALID;
    if (n == 0) return F9_OK;
    for (int i = 0; i < n; ++i)
        if (src[i].numCh <= 0 || src[i].numFrames < 0 || (src[i].numFrames > 0 && (!d_dsts[i] || !src[i].base))) return ctx->fail(F9_ERR_INVALID, "bad file descriptor in the batch");
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const DevBuf* hb = reinterpret_cast<const DevBuf*>(src);
    int rc = ctx->arena_reserve((sizeof(DevBuf) + sizeof(void*)) * (size_t) n + 16384, (sizeof(DevBuf) + sizeof(void*)) * (size_t) n + 8192, true);
    if (rc) return rc;
    DevBuf* d_b; unsigned char** d_p;
    rc = upload_array(ctx, hb, (size_t) n, &d_b); if (rc) return rc;
    rc = upload_array(ctx, const_cast<unsigned char**>(d_dsts), (size_t) n, &d_p); if (rc) return rc;
    F9_TRY_CUDA(ctx, launch_planar_to_pcm24_batch(hb, d_b, d_p, n, ctx->stream, &ctx->launches, ctx->diag.has("F9_PCM_BYTEWISE") ? nullptr : d_dsts));
    return F9_OK;
}

int f9_pcm_to_planar(f9_context* ctx, const void* src, int fmt, int src_ch, long long num_frames, float* const* dst, int dst_ch) {
    if (!ctx) return F9_ERR_INVALID;
    if (fmt < F9_PCM_U8 || fmt > F9_PCM_F32LE || src_ch <= 0 || dst_ch <= 0 || num_frames < 0 || (num_frames > 0 && (!src || !dst)))
        return ctx->fail(F9_ERR_INVALID, "bad pcm arguments");
    if (num_frames == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const int bps = (fmt == F9_PCM_U8) ? 1 : (fmt == F9_PCM_S16LE) ? 2 : (fmt == F9_PCM_S24LE) ? 3 : 4;
    const size_t src_bytes = (size_t) num_frames * src_ch * bps;
    const long long stride = pad_stride(num_frames);
    int rc = ctx->arena_reserve(src_bytes + sizeof(float) * (size_t) stride * dst_ch + 16384, 4096); if (rc) return rc;
    unsigned char* d_src = (unsigned char*) ctx->d_alloc(src_bytes);
    float* d_dst = (float*) ctx->d_alloc(sizeof(float) * (size_t) stride * dst_ch);
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, ctx->stream));
    F9_TRY_CUDA(ctx, launch_pcm_to_planar(d_src, fmt, src_ch, num_frames, d_dst, stride, dst_ch, ctx->stream, &ctx->launches));
    for (int c = 0; c < dst_ch; ++c)
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(dst[c], d_dst + c * stride, sizeof(float) * (size_t) num_frames, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}

int f9_planar_to_pcm24(f9_context* ctx, const float* const* src, int numCh, long long num_frames, unsigned char* dst) {
    int rc = check_planar(ctx, src, numCh, num_frames); if (rc) return rc;
    if (numCh <= 0 || (num_frames > 0 && !dst)) return ctx->fail(F9_ERR_INVALID, "bad pcm24 arguments");
    if (num_frames == 0) return F9_OK;
    const size_t out_bytes = (size_t) num_frames * numCh * 3;
    rc = ctx->arena_reserve(planar_bytes(numCh, num_frames) + out_bytes + 16384, 4096); if (rc) return rc;
    DevBuf hb{};
    rc = upload_planar(ctx, src, numCh, num_frames, &hb); if (rc) return rc;
    unsigned char* d_dst = (unsigned char*) ctx->d_alloc(out_bytes);
    F9_TRY_CUDA(ctx, launch_planar_to_pcm24(hb.base, hb.chStride, numCh, num_frames, d_dst, ctx->stream, &ctx->launches));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(dst, d_dst, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}

int f9_interleave(f9_context* ctx, const float* const* src, int numCh, long long num_frames, float* dst) {
    int rc = check_planar(ctx, src, numCh, num_frames); if (rc) return rc;
    if (numCh <= 0 || (num_frames > 0 && !dst)) return ctx->fail(F9_ERR_INVALID, "bad interleave arguments");
    if (num_frames == 0) return F9_OK;
    const size_t out_bytes = sizeof(float) * (size_t) num_frames * numCh;
    rc = ctx->arena_reserve(planar_bytes(numCh, num_frames) + out_bytes + 16384, 4096); if (rc) return rc;
    DevBuf hb{};
    rc = upload_planar(ctx, src, numCh, num_frames, &hb); if (rc) return rc;
    float* d_dst = (float*) ctx->d_alloc(out_bytes);
    F9_TRY_CUDA(ctx, launch_interleave(hb.base, hb.chStride, numCh, num_frames, d_dst, ctx->stream, &ctx->launches));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(dst, d_dst, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}

int f9_deinterleave(f9_context* ctx, const float* src, int numCh, long long num_frames, float* const* dst) {
    if (!ctx) return F9_ERR_INVALID;
    if (numCh <= 0 || num_frames < 0 || (num_frames > 0 && (!src || !dst))) return ctx->fail(F9_ERR_INVALID, "bad deinterleave arguments");
    if (num_frames == 0) return F9_OK;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t in_bytes = sizeof(float) * (size_t) num_frames * numCh;
    const long long stride = pad_stride(num_frames);
    int rc = ctx->arena_reserve(in_bytes + sizeof(float) * (size_t) stride * numCh + 16384, 4096); if (rc) return rc;
    float* d_src = (float*) ctx->d_alloc(in_bytes);
    float* d_dst = (float*) ctx->d_alloc(sizeof(float) * (size_t) stride * numCh);
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_src, src, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    F9_TRY_CUDA(ctx, launch_deinterleave(d_src, numCh, num_frames, d_dst, stride, ctx->stream, &ctx->launches));
    for (int c = 0; c < numCh; ++c)
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(dst[c], d_dst + c * stride, sizeof(float) * (size_t) num_frames, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    return F9_OK;
}

}  // extern "C"
