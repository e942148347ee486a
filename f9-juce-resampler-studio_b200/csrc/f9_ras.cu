// juce::ResamplingAudioSource on the GPU (JUCE 8.0.10 juce_audio_basics/sources/juce_ResamplingAudioSource.cpp,
// named by north_star next to the Interpolators; the reference has no call site): linear interpolation with a double
// sub-sample position, a 2nd-order Butterworth low-pass with double state on the input when ratio > 1.0001 and on the output
// when ratio < 0.9999.
//
//   * f9_ras_* : the stateful object, getNextAudioBlock for getNextAudioBlock.  The control flow (how many samples are
//     pulled, the position recurrence, the filter "stoking" near ratio 1) runs on the host exactly as JUCE's; the sample
//     arithmetic (filters, lerp) runs in kernels that use JUCE's operation order with round-to-nearest intrinsics, so a
//     block is bit-identical to the scalar code.
//   * f9_ras_convert / f9_dev_ras_convert : whole channels from reset state, batched.  The IIR is run chunk-parallel: a
//     thread starts a warm-up of W samples before its chunk from zero state, W chosen from the pole radius so that the
//     missing history is below 1e-17 (a stable biquad forgets its state geometrically); the position is the closed form
//     m * ratio.  Both differ from the sequential code by double rounding noise only: parity within the 2^-20 tolerance.
//     Ratios whose poles sit too close to the unit circle run one thread per channel.
// Product code: nothing here includes or links oracle/.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "f9_internal.cuh"

namespace f9 {
namespace {

struct RasCoef { double c0, c1, c2, c4, c5; int intel; };

// createLowPass + setFilterCoefficients
RasCoef ras_coefficients(double frequencyRatio, bool intel) {
    const double pi = 3.14159265358979323846;
    const double proportionalRate = frequencyRatio > 1.0 ? 0.5 / frequencyRatio : 0.5 * frequencyRatio;
    const double n = 1.0 / std::tan(pi * std::max(0.001, proportionalRate));
    const double nSquared = n * n;
    const double c1 = 1.0 / (1.0 + std::sqrt(2.0) * n + nSquared);
    const double c[6] = {c1, c1 * 2.0, c1, 1.0, c1 * 2.0 * (1.0 - nSquared), c1 * (1.0 - std::sqrt(2.0) * n + nSquared)};
    const double a = 1.0 / c[3];
    RasCoef k; k.c0 = c[0] * a; k.c1 = c[1] * a; k.c2 = c[2] * a; k.c4 = c[4] * a; k.c5 = c[5] * a; k.intel = intel ? 1 : 0;
    return k;
}
// samples after which a unit of filter state has decayed below 1e-17 (0: do not chunk)
int ras_warmup(const RasCoef& k) {
    const double r2 = std::fabs(k.c5);                        // |pole|^2 for a complex pair; an upper bound is enough
    const double disc = k.c4 * k.c4 - 4.0 * k.c5;
    double radius = disc < 0.0 ? std::sqrt(r2) : 0.5 * (std::fabs(k.c4) + std::sqrt(disc));
    if (!(radius < 0.995)) return 0;
    if (radius < 1e-3) radius = 1e-3;
    const int w = (int) std::ceil(std::log(1e-17) / std::log(radius)) + 8;
    return std::min(std::max(w, 64), 8192);
}

// applyFilter, JUCE's order: ((((c0 in + c1 x1) + c2 x2) - c4 y1) - c5 y2), products and sums rounded separately
__device__ __forceinline__ double ras_step(const RasCoef& k, double in, double& x1, double& x2, double& y1, double& y2) {
    double out = __dadd_rn(__dadd_rn(__dmul_rn(k.c0, in), __dmul_rn(k.c1, x1)), __dmul_rn(k.c2, x2));
    out = __dadd_rn(out, -__dmul_rn(k.c4, y1));
    out = __dadd_rn(out, -__dmul_rn(k.c5, y2));
    if (k.intel && !(out < -1.0e-8 || out > 1.0e-8)) out = 0.0;
    x2 = x1; x1 = in; y2 = y1; y1 = out;
    return out;
}

// One thread per (stream, chunk).  src/dst: stream s at base + s * stride (may alias only when chunk >= n: one thread per
// stream).  Samples past nValid read as zero (the input source past the end of the file).  state: 4 doubles per stream,
// read when there is a single chunk, written back by the thread of the last chunk.
__global__ void __launch_bounds__(128)
ras_biquad_kernel(const float* src, long long srcStride, float* dst, long long dstStride, int nStreams,
                  long long n, long long nValid, long long chunk, int warm, RasCoef k, double* __restrict__ state) {
    const long long chunksPer = (n + chunk - 1) / chunk;
    const long long id = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= chunksPer * nStreams) return;
    const int s = (int) (id / chunksPer);
    const long long c0 = (id % chunksPer) * chunk, c1 = min(n, c0 + chunk);
    const float* in = src + (long long) s * srcStride;
    float* out = dst + (long long) s * dstStride;
    double x1 = 0, x2 = 0, y1 = 0, y2 = 0;
    if (chunksPer == 1 && state) { x1 = state[4 * s]; x2 = state[4 * s + 1]; y1 = state[4 * s + 2]; y2 = state[4 * s + 3]; }
    for (long long i = max(0LL, c0 - warm); i < c0; ++i) ras_step(k, i < nValid ? (double) in[i] : 0.0, x1, x2, y1, y2);
    for (long long i = c0; i < c1; ++i) out[i] = (float) ras_step(k, i < nValid ? (double) in[i] : 0.0, x1, x2, y1, y2);
    if (state && c1 == n) { state[4 * s] = x1; state[4 * s + 1] = x2; state[4 * s + 2] = y1; state[4 * s + 3] = y2; }
}

// Chunk-parallel form with coalesced traffic: a CTA of 128 threads owns 128 consecutive chunks of one stream.  Thread j walks
// chunk j (after its warm-up) sequentially, but the samples travel through shared memory 32 per chunk at a time: a warp reads
// or writes 32 consecutive floats of one chunk (one 128-byte line) per instruction, the threads then read their own row (pitch
// 33: conflict-free).  With one thread streaming its own chunk from global memory every load touched 32 different lines and
// every store wrote 4 bytes of a sector (8 ms for 512 channels of 10 s; this form is bound by the FP64 recurrence).
constexpr int kRasT = 32;                        // samples per chunk and step
__global__ void __launch_bounds__(128, 6)
ras_biquad_tiled_kernel(const float* __restrict__ src, long long srcStride, float* __restrict__ dst, long long dstStride,
                        long long n, long long nValid, int chunk, int warm, RasCoef k) {
    __shared__ float tile[128][kRasT + 1];
    const long long chunksPer = (n + chunk - 1) / chunk;
    const long long blocksPer = (chunksPer + 127) / 128;
    const int s = (int) (blockIdx.x / blocksPer);
    const long long cb = (blockIdx.x % blocksPer) * 128;          // first chunk of this CTA
    const float* in = src + (long long) s * srcStride;
    float* out = dst + (long long) s * dstStride;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long c0 = (cb + t) * (long long) chunk;             // this thread's chunk [c0, c0 + chunk)
    double x1 = 0, x2 = 0, y1 = 0, y2 = 0;
    // relative sample index r runs over [-warm, chunk) in steps of kRasT (warm and chunk are multiples of kRasT).  The 32 row pieces a
    // warp brings in per step are loaded a step ahead into registers, all 32 loads back to back: with a load and its shared-memory
    // store per loop iteration the stores waited for one memory latency each (ncu: 48 % of all stalls on the first STS), and the
    // recurrence of the step now runs while the next step's loads are in flight.
    float nx[32];
    const long long g00 = (cb + w) * (long long) chunk + lane;   // row w of the CTA, this lane's sample of a step at r0 = 0
    auto fetch = [&](int r0) {
        const long long g0 = g00 + r0;
        const float* p0 = in + g0;
        const bool inside = g0 >= 0 && g0 + 124LL * chunk < nValid;            // every row's sample of this lane exists
        if (inside) {
            #pragma unroll
            for (int k = 0; k < 32; ++k) nx[k] = __ldg(p0 + (size_t) (4 * k) * (size_t) chunk);
        } else {
            #pragma unroll
            for (int k = 0; k < 32; ++k) {
                const long long g = g0 + (long long) (4 * k) * chunk;
                nx[k] = (g >= 0 && g < nValid) ? __ldg(in + g) : 0.0f;
            }
        }
    };
    fetch(-warm);
    for (int r0 = -warm; r0 < chunk; r0 += kRasT) {
        #pragma unroll
        for (int k = 0; k < 32; ++k) tile[w + 4 * k][lane] = nx[k];
        __syncthreads();
        if (r0 + kRasT < chunk) fetch(r0 + kRasT);
        const bool live = c0 + r0 >= 0 && c0 < n;                   // before sample 0 the state is the reset state: nothing to run
        if (live) {
            #pragma unroll 4
            for (int i = 0; i < kRasT; ++i) tile[t][i] = (float) ras_step(k, (double) tile[t][i], x1, x2, y1, y2);
        }
        __syncthreads();
        if (r0 >= 0) {
            #pragma unroll 8
            for (int j = w; j < 128; j += 4) {
                const long long g = (cb + j) * (long long) chunk + r0 + lane;
                if (g < n && (cb + j) < chunksPer) out[g] = tile[j][lane];
            }
        }
        __syncthreads();
    }
}

// out[s][m] = src[pos] + alpha * (src[pos + 1] - src[pos]) in float, products and sums rounded separately.
// idx / alpha: per-output position (host recurrence, the stateful object) or nullptr: closed form m * ratio from offset 0
// (exact integers when the ratio is p / q).  Samples past nValid read as zero.
__global__ void __launch_bounds__(256)
ras_lerp_kernel(const float* __restrict__ src, long long srcStride, long long nValid, float* __restrict__ dst, long long dstStride,
                int nStreams, long long numOut, const int* __restrict__ idx, const float* __restrict__ alphas,
                double ratio, long long p, long long q) {
    const long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= numOut) return;
    long long pos; float alpha;
    if (idx) { pos = idx[m]; alpha = alphas[m]; }
    else if (q > 0) { const long long t = m * p; pos = t / q; alpha = (float) ((double) (t - pos * q) / (double) q); }
    else { const double t = (double) m * ratio; const double f = floor(t); pos = (long long) f; alpha = (float) (t - f); }
    for (int s = blockIdx.y; s < nStreams; s += gridDim.y) {
        const float* in = src + (long long) s * srcStride;
        const float a = pos < nValid ? in[pos] : 0.0f, b = pos + 1 < nValid ? in[pos + 1] : 0.0f;
        dst[(long long) s * dstStride + m] = __fadd_rn(a, __fmul_rn(alpha, __fadd_rn(b, -a)));
    }
}

inline long long pad64(long long n) { return (std::max<long long>(n, 1) + 63) / 64 * 64; }

// Whole streams from reset state on device buffers: src (nStreams x nIn, stride) -> dst (nStreams x numOut, stride).
// scratch: nStreams * pad64(max(filtered length, numOut)) floats.
int ras_convert_device(f9_context* ctx, const float* d_src, long long srcStride, int nStreams, long long nIn, double ratio,
                       float* d_dst, long long dstStride, long long numOut, float* d_scratch, long long scratchStride, bool intel) {
    if (numOut <= 0 || nStreams <= 0) return F9_OK;
    cudaStream_t s = ctx->stream;
    const RasCoef k = ras_coefficients(ratio, intel);
    long long p = 0, q = 0;
    if (!find_rational(ratio, 1 << 20, &p, &q) || p > (1LL << 40) / std::max<long long>(numOut, 1)) { p = 0; q = 0; }
    const long long need = (long long) std::floor((double) (numOut - 1) * ratio) + 2;      // source samples the last output reads
    auto filter = [&](const float* src, long long sStride, long long n, long long nValid, float* dst, long long dStride) {
        int warm = ras_warmup(k);
        if (warm > 0 && src != dst) {                          // chunk-parallel, staged through shared memory
            warm = (warm + kRasT - 1) / kRasT * kRasT;
            const int chunk = std::max(4 * warm, 2048);
            const long long chunksPer = (n + chunk - 1) / chunk, blocksPer = (chunksPer + 127) / 128;
            ras_biquad_tiled_kernel<<<(unsigned) (blocksPer * nStreams), 128, 0, s>>>(src, sStride, dst, dStride, n, nValid, chunk, warm, k);
        } else {                                               // poles too close to the unit circle: one thread per stream
            ras_biquad_kernel<<<(unsigned) ((nStreams + 127) / 128), 128, 0, s>>>(src, sStride, dst, dStride, nStreams, n, nValid, n, 0, k, nullptr);
        }
        ++ctx->launches;
        return cudaGetLastError();
    };
    const dim3 lgrid((unsigned) ((numOut + 255) / 256), (unsigned) std::min(nStreams, 64));
    if (ratio > 1.0001) {                                      // down-sampling: filter the input, then interpolate
        F9_TRY_CUDA(ctx, filter(d_src, srcStride, need, nIn, d_scratch, scratchStride));
        ras_lerp_kernel<<<lgrid, 256, 0, s>>>(d_scratch, scratchStride, need, d_dst, dstStride, nStreams, numOut, nullptr, nullptr, ratio, p, q);
        ++ctx->launches;
        F9_TRY_CUDA(ctx, cudaGetLastError());
    } else if (ratio < 0.9999) {                               // up-sampling: interpolate, then filter the output
        ras_lerp_kernel<<<lgrid, 256, 0, s>>>(d_src, srcStride, nIn, d_scratch, scratchStride, nStreams, numOut, nullptr, nullptr, ratio, p, q);
        ++ctx->launches;
        F9_TRY_CUDA(ctx, cudaGetLastError());
        F9_TRY_CUDA(ctx, filter(d_scratch, scratchStride, numOut, numOut, d_dst, dstStride));
    } else {
        ras_lerp_kernel<<<lgrid, 256, 0, s>>>(d_src, srcStride, nIn, d_dst, dstStride, nStreams, numOut, nullptr, nullptr, ratio, p, q);
        ++ctx->launches;
        F9_TRY_CUDA(ctx, cudaGetLastError());
    }
    return F9_OK;
}

}  // namespace
}  // namespace f9

using namespace f9;

// ------------------------------------------------------------------------------------------------ stateful object
struct f9_resampling_source {
    f9_context* ctx = nullptr;
    int numChannels = 0;
    double ratio = 1.0, lastRatio = 1.0;
    RasCoef coef{};
    bool intel = true, prepared = false;
    double subSampleOffset = 0.0;
    std::vector<std::vector<float>> pending;       // the samples JUCE's ring buffer holds: [bufferPos, bufferPos + sampsInBuffer)
    std::vector<double> state;                     // x1, x2, y1, y2 per channel
};

extern "C" {

int f9_ras_create(f9_context* ctx, int num_channels, f9_resampling_source** out) {
    if (!ctx || !out || num_channels <= 0) return F9_ERR_INVALID;
    f9_resampling_source* h = new (std::nothrow) f9_resampling_source();
    if (!h) return F9_ERR_NOMEM;
    h->ctx = ctx; h->numChannels = num_channels;
    h->pending.assign((size_t) num_channels, std::vector<float>());
    h->state.assign((size_t) num_channels * 4, 0.0);
    h->coef = ras_coefficients(1.0, true);
    *out = h;
    return F9_OK;
}
void f9_ras_destroy(f9_resampling_source* h) { delete h; }
int f9_ras_set_intel_denormal_flush(f9_resampling_source* h, int on) {
    if (!h) return F9_ERR_INVALID;
    h->intel = on != 0; h->coef.intel = on != 0;
    return F9_OK;
}
int f9_ras_set_resampling_ratio(f9_resampling_source* h, double samples_in_per_output_sample) {
    if (!h || !std::isfinite(samples_in_per_output_sample)) return F9_ERR_INVALID;
    h->ratio = std::max(0.0, samples_in_per_output_sample);
    return F9_OK;
}
double f9_ras_get_resampling_ratio(const f9_resampling_source* h) { return h ? h->ratio : 0.0; }
int f9_ras_flush_buffers(f9_resampling_source* h) {
    if (!h) return F9_ERR_INVALID;
    for (auto& p : h->pending) p.clear();
    h->subSampleOffset = 0.0;
    std::fill(h->state.begin(), h->state.end(), 0.0);
    return F9_OK;
}
int f9_ras_prepare_to_play(f9_resampling_source* h, int samples_per_block_expected, double sample_rate) {
    if (!h || samples_per_block_expected < 0) return F9_ERR_INVALID;
    (void) sample_rate;                                         // forwarded to the input source by JUCE; no arithmetic depends on it
    h->coef = ras_coefficients(h->ratio, h->intel);             // createLowPass(ratio); lastRatio keeps its value, as in JUCE
    h->prepared = true;
    return f9_ras_flush_buffers(h);
}
int f9_ras_release_resources(f9_resampling_source* h) {
    if (!h) return F9_ERR_INVALID;
    for (auto& p : h->pending) { p.clear(); p.shrink_to_fit(); }
    return F9_OK;
}

int f9_ras_num_samples_to_pull(const f9_resampling_source* h, int num_samples) {
    if (!h || num_samples < 0) return F9_ERR_INVALID;
    const double needD = (double) num_samples * h->ratio;
    if (!(needD < 1.0e9)) return F9_ERR_INVALID;
    return std::max(0, (int) std::lrint(needD) + 3 - (int) h->pending[0].size());
}

int f9_ras_get_next_audio_block(f9_resampling_source* h, const float* const* in, int num_in_available, float* const* out, int num_samples) {
    if (!h || num_samples < 0 || num_in_available < 0 || (num_samples > 0 && !out) || (num_in_available > 0 && !in)) return F9_ERR_INVALID;
    f9_context* ctx = h->ctx;
    const int nCh = h->numChannels;
    const double localRatio = h->ratio;
    if (h->lastRatio != localRatio) { h->coef = ras_coefficients(localRatio, h->intel); h->lastRatio = localRatio; }
    const double needD = (double) num_samples * localRatio;
    if (!(needD < 1.0e9)) return ctx->fail(F9_ERR_INVALID, "block too long for this ratio");
    const int sampsNeeded = (int) std::lrint(needD) + 3;
    const int have = (int) h->pending[0].size();
    const int pull = std::max(0, sampsNeeded - have);
    const int total = have + pull;
    if (num_samples == 0 && pull == 0) return 0;
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));

    // host: position recurrence (JUCE's, so bufferPos / subSampleOffset carry over bit for bit)
    std::vector<int> idx((size_t) num_samples); std::vector<float> alpha((size_t) num_samples);
    double sub = h->subSampleOffset; int pos = 0;
    for (int m = 0; m < num_samples; ++m) {
        idx[(size_t) m] = pos; alpha[(size_t) m] = (float) sub;
        sub += localRatio;
        while (sub >= 1.0) { ++pos; sub -= 1.0; }
    }
    if (pos > total) return ctx->fail(F9_ERR_INVALID, "ratio consumed more input than the block pulled");

    const long long sStride = pad64(total + 1), oStride = pad64(num_samples);
    const size_t dBytes = sizeof(float) * (size_t) nCh * (size_t) (sStride + oStride) + sizeof(double) * 4 * (size_t) nCh +
                          (sizeof(int) + sizeof(float)) * (size_t) std::max(num_samples, 1) + 65536;
    int rc = ctx->arena_reserve(dBytes, dBytes); if (rc) return rc;
    float* h_src = (float*) ctx->h_alloc(sizeof(float) * (size_t) nCh * (size_t) sStride);
    float* d_src = (float*) ctx->d_alloc(sizeof(float) * (size_t) nCh * (size_t) sStride);
    float* d_out = (float*) ctx->d_alloc(sizeof(float) * (size_t) nCh * (size_t) oStride);
    float* h_out = (float*) ctx->h_alloc(sizeof(float) * (size_t) nCh * (size_t) oStride);
    double* h_state = (double*) ctx->h_alloc(sizeof(double) * 4 * (size_t) nCh);
    double* d_state = (double*) ctx->d_alloc(sizeof(double) * 4 * (size_t) nCh);
    int* h_idx = (int*) ctx->h_alloc(sizeof(int) * (size_t) std::max(num_samples, 1));
    int* d_idx = (int*) ctx->d_alloc(sizeof(int) * (size_t) std::max(num_samples, 1));
    float* h_alpha = (float*) ctx->h_alloc(sizeof(float) * (size_t) std::max(num_samples, 1));
    float* d_alpha = (float*) ctx->d_alloc(sizeof(float) * (size_t) std::max(num_samples, 1));
    for (int c = 0; c < nCh; ++c) {
        float* row = h_src + (size_t) c * (size_t) sStride;
        if (have) std::memcpy(row, h->pending[(size_t) c].data(), sizeof(float) * (size_t) have);
        const int real = std::min(pull, num_in_available);      // the input source delivers zeros past its end
        if (real) std::memcpy(row + have, in[c], sizeof(float) * (size_t) real);
        std::fill(row + have + real, row + sStride, 0.0f);
    }
    std::memcpy(h_state, h->state.data(), sizeof(double) * 4 * (size_t) nCh);
    if (num_samples) { std::memcpy(h_idx, idx.data(), sizeof(int) * (size_t) num_samples); std::memcpy(h_alpha, alpha.data(), sizeof(float) * (size_t) num_samples); }
    cudaStream_t s = ctx->stream;
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_src, h_src, sizeof(float) * (size_t) nCh * (size_t) sStride, cudaMemcpyHostToDevice, s));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_state, h_state, sizeof(double) * 4 * (size_t) nCh, cudaMemcpyHostToDevice, s));
    if (num_samples) {
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_idx, h_idx, sizeof(int) * (size_t) num_samples, cudaMemcpyHostToDevice, s));
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_alpha, h_alpha, sizeof(float) * (size_t) num_samples, cudaMemcpyHostToDevice, s));
    }
    const unsigned fgrid = (unsigned) ((nCh + 127) / 128);
    if (localRatio > 1.0001 && pull > 0) {                      // pre-filter the newly pulled samples, in place, one thread per channel
        ras_biquad_kernel<<<fgrid, 128, 0, s>>>(d_src + have, sStride, d_src + have, sStride, nCh, pull, pull, pull, 0, h->coef, d_state);
        ++ctx->launches; F9_TRY_CUDA(ctx, cudaGetLastError());
    }
    if (num_samples) {
        const dim3 lgrid((unsigned) ((num_samples + 255) / 256), (unsigned) std::min(nCh, 64));
        ras_lerp_kernel<<<lgrid, 256, 0, s>>>(d_src, sStride, total, d_out, oStride, nCh, num_samples, d_idx, d_alpha, localRatio, 0, 0);
        ++ctx->launches; F9_TRY_CUDA(ctx, cudaGetLastError());
        if (localRatio < 0.9999) {                              // post-filter the block
            ras_biquad_kernel<<<fgrid, 128, 0, s>>>(d_out, oStride, d_out, oStride, nCh, num_samples, num_samples, num_samples, 0, h->coef, d_state);
            ++ctx->launches; F9_TRY_CUDA(ctx, cudaGetLastError());
        }
        F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(float) * (size_t) nCh * (size_t) oStride, cudaMemcpyDeviceToHost, s));
    }
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_src, d_src, sizeof(float) * (size_t) nCh * (size_t) sStride, cudaMemcpyDeviceToHost, s));
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_state, d_state, sizeof(double) * 4 * (size_t) nCh, cudaMemcpyDeviceToHost, s));
    F9_FINISH(ctx);

    std::memcpy(h->state.data(), h_state, sizeof(double) * 4 * (size_t) nCh);
    for (int c = 0; c < nCh; ++c) {
        const float* row = h_src + (size_t) c * (size_t) sStride;
        h->pending[(size_t) c].assign(row + pos, row + total);  // what stays in the ring: [bufferPos, endOfBufferPos)
        if (num_samples) std::memcpy(out[c], h_out + (size_t) c * (size_t) oStride, sizeof(float) * (size_t) num_samples);
    }
    h->subSampleOffset = sub;
    if (localRatio >= 0.9999 && localRatio <= 1.0001 && num_samples > 0) {       // keep the idle filter stoked with the last outputs
        for (int c = 0; c < nCh; ++c) {
            double* fs = h->state.data() + 4 * (size_t) c;       // x1, x2, y1, y2
            const float* end = out[c] + num_samples - 1;
            if (num_samples > 1) fs[3] = fs[1] = (double) *(end - 1);
            else { fs[3] = fs[2]; fs[1] = fs[0]; }
            fs[2] = fs[0] = (double) *end;
        }
    }
    return pull;
}

// ------------------------------------------------------------------------------------------------ whole channels, batched
int f9_dev_ras_convert(f9_context* ctx, const float* d_in, long long in_stride, int num_streams, long long num_in, double ratio,
                       float* d_out, long long out_stride, long long num_out, float* d_scratch, long long scratch_stride) {
    if (!ctx || num_streams < 0 || num_in < 0 || num_out < 0 || !(ratio > 0.0) || !std::isfinite(ratio)) return F9_ERR_INVALID;
    if (num_streams == 0 || num_out == 0) return F9_OK;
    if (!d_in || !d_out || !d_scratch || scratch_stride < f9_ras_scratch_frames(ratio, num_out)) return ctx->fail(F9_ERR_INVALID, "bad ras buffers");
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    return ras_convert_device(ctx, d_in, in_stride, num_streams, num_in, ratio, d_out, out_stride, num_out, d_scratch, scratch_stride, true);
}
long long f9_ras_scratch_frames(double ratio, long long num_out) {
    if (!(ratio > 0.0) || num_out <= 0) return 64;
    return pad64(std::max<long long>(num_out, (long long) std::floor((double) (num_out - 1) * ratio) + 2));
}
int f9_ras_convert(f9_context* ctx, const float* const* in, int numCh, long long num_in, double ratio, float* const* out, long long num_out) {
    if (!ctx || numCh <= 0 || num_in < 0 || num_out < 0 || !(ratio > 0.0) || !std::isfinite(ratio)) return F9_ERR_INVALID;
    if (num_out == 0) return F9_OK;
    if (!out || (num_in > 0 && !in)) return ctx->fail(F9_ERR_INVALID, "bad ras buffers");
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long iStride = pad64(num_in), oStride = pad64(num_out), sStride = f9_ras_scratch_frames(ratio, num_out);
    const size_t dBytes = sizeof(float) * (size_t) numCh * (size_t) (iStride + oStride + sStride) + 65536;
    const size_t hBytes = sizeof(float) * (size_t) numCh * (size_t) (iStride + oStride) + 65536;
    int rc = ctx->arena_reserve(dBytes, hBytes); if (rc) return rc;
    float* h_in = (float*) ctx->h_alloc(sizeof(float) * (size_t) numCh * (size_t) iStride);
    float* d_in = (float*) ctx->d_alloc(sizeof(float) * (size_t) numCh * (size_t) iStride);
    float* d_out = (float*) ctx->d_alloc(sizeof(float) * (size_t) numCh * (size_t) oStride);
    float* d_scr = (float*) ctx->d_alloc(sizeof(float) * (size_t) numCh * (size_t) sStride);
    float* h_out = (float*) ctx->h_alloc(sizeof(float) * (size_t) numCh * (size_t) oStride);
    for (int c = 0; c < numCh; ++c) if (num_in) std::memcpy(h_in + (size_t) c * (size_t) iStride, in[c], sizeof(float) * (size_t) num_in);
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_in, h_in, sizeof(float) * (size_t) numCh * (size_t) iStride, cudaMemcpyHostToDevice, ctx->stream));
    rc = ras_convert_device(ctx, d_in, iStride, numCh, num_in, ratio, d_out, oStride, num_out, d_scr, sStride, true); if (rc) return rc;
    F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(float) * (size_t) numCh * (size_t) oStride, cudaMemcpyDeviceToHost, ctx->stream));
    F9_FINISH(ctx);
    for (int c = 0; c < numCh; ++c) std::memcpy(out[c], h_out + (size_t) c * (size_t) oStride, sizeof(float) * (size_t) num_out);
    return F9_OK;
}

}  // extern "C"
