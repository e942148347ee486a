// =============================================================================
// f9_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// This file is a scalar CPU restatement of the reference's DSP hot path.  It is
// the *checker* for the CUDA product under f9-juce-resampler-studio_b200/.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.  The product never links, imports or calls anything here.
//
// Build: g++ -O2 -ffp-contract=off (see oracle/Makefile) so every float result
// is reproducible (no FMA contraction).
//
// Parity status
//   * trim / peak / RMS / noise-floor / thresholds / recording length:
//     restated from the reference's own C++ and Swift (file:line on each
//     function, paths relative to /root/reference).  PINNED by the worked
//     examples in the reference's docs (tests/golden/doc_vectors.json).
//   * juce::Interpolators (WindowedSinc, Lagrange, ...), AudioFormatReader /
//     Writer int<->float conversion: the arithmetic lives in JUCE 8.0.10
//     (module juce_audio_basics / juce_audio_formats), which is NOT vendored in
//     the reference and not on this machine.  These are restated from the
//     published algorithm ("[JUCE-recall]").  PARITY UNPINNED for resampled
//     sample values.  In particular WindowedSincTraits::lookupTable[10001] is a
//     literal table in JUCE whose generating window is unknown; sinc_table()
//     below is a documented stand-in and every consumer takes the table as a
//     parameter so the real one can be swapped in.
//   * bounded-lag cross-correlation: the reference has none (it peak-picks an
//     impulse); defined here so that it degenerates to findPeakPosition for an
//     impulse stimulus, with the same scan order and strict-> tie-breaking.
// =============================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

// Minimal stand-in for juce::AudioBuffer<float> read access (planar float32).
struct Planar {
    const float* const* ch;
    int numCh;
    int numFrames;
};

// ---- Source/MainComponent.cpp:983-1004  calculateRMS ------------------------
// float product (rounded to float), widened, sequential double sum, channel-major.
float rms_of(const Planar& b) {
    double sumOfSquares = 0.0;
    int total = 0;
    for (int c = 0; c < b.numCh; ++c) {
        const float* d = b.ch[c];
        for (int i = 0; i < b.numFrames; ++i) {
            float sq = d[i] * d[i];
            sumOfSquares += sq;
            ++total;
        }
    }
    if (total == 0) return 0.0f;
    return (float) std::sqrt(sumOfSquares / total);
}

// ---- Source/AppState.h:252-258  getNoiseFloorThresholdDb --------------------
float nf_threshold_db(int hasNf, float nfDb, float marginPct) {
    if (!hasNf) return -80.0f;
    return nfDb + (nfDb * marginPct / 100.0f);
}

}  // namespace

// =============================================================================
// (2) latency detection
// =============================================================================

// Source/MainComponent.cpp:950-975  findPeakPosition (planar).
// Strict '>' so ties keep the earliest sample of the lowest channel; all-zero
// gives -1; NaN never wins; returns the frame index only.
ORC_API int orc_find_peak_position(const float* const* ch, int numCh, int numFrames, float threshold) {
    float maxValue = 0.0f;
    int maxPosition = -1;
    for (int c = 0; c < numCh; ++c) {
        const float* d = ch[c];
        for (int i = 0; i < numFrames; ++i) {
            float a = std::fabs(d[i]);
            if (a > maxValue) { maxValue = a; maxPosition = i; }
        }
    }
    return (maxValue > threshold) ? maxPosition : -1;
}

// _Swift Code/.../Services/LatencyMeasurementService.swift:147-171
// analyzeCapturedAudio on interleaved data: index is the interleaved index,
// default index 0, *found = 0 when maxValue <= threshold (Swift throws).
ORC_API long long orc_find_peak_interleaved(const float* audio, long long n, float threshold, int* found) {
    float maxValue = 0.0f;
    long long maxIndex = 0;
    for (long long i = 0; i < n; ++i) {
        float a = std::fabs(audio[i]);
        if (a > maxValue) { maxValue = a; maxIndex = i; }
    }
    if (found) *found = (maxValue > threshold) ? 1 : 0;
    return maxIndex;
}

// Source/MainComponent.cpp:983-1004
ORC_API float orc_calculate_rms(const float* const* ch, int numCh, int numFrames) {
    return rms_of(Planar{ch, numCh, numFrames});
}

// Source/MainComponent.cpp:977-981  20*log10f(max(rms,1e-6f))
ORC_API float orc_noise_floor_db(const float* const* ch, int numCh, int numFrames) {
    float rms = rms_of(Planar{ch, numCh, numFrames});
    return 20.0f * std::log10(std::max(rms, 1e-6f));
}

// LatencyMeasurementService.swift:173-181  calculateNoiseFloor: Float reduce, Float sqrt.
ORC_API float orc_noise_floor_db_swift(const float* audio, long long n) {
    float sum = 0.0f;
    for (long long i = 0; i < n; ++i) sum = sum + audio[i] * audio[i];
    float rms = std::sqrt(sum / (float) n);
    return 20.0f * std::log10(std::max(rms, 1e-6f));
}

// =============================================================================
// settings math  (Source/AppState.h:221-258, Models/ProcessingSettings.swift:59-88)
// =============================================================================
ORC_API int orc_recording_length(int sourceFileSamples, int latencySamples) {          // AppState.h:240-243
    return sourceFileSamples + latencySamples + (latencySamples * 4);
}
ORC_API float orc_threshold_linear(float thresholdDb) {                               // AppState.h:246-249
    return std::pow(10.0f, thresholdDb / 20.0f);
}
ORC_API float orc_noise_floor_threshold_db(int hasNf, float nfDb, float marginPct) {   // AppState.h:252-258
    return nf_threshold_db(hasNf, nfDb, marginPct);
}
ORC_API double orc_latency_ms(int measuredLatencySamples, double sampleRate) {         // AppState.h:231-237
    if (measuredLatencySamples < 0) return 0.0;
    return ((double) measuredLatencySamples / sampleRate) * 1000.0;
}
ORC_API int orc_needs_latency_remeasurement(int measuredLatencySamples, int lastBuf, int curBuf) {  // AppState.h:221-228
    if (measuredLatencySamples < 0) return 1;
    return lastBuf != curBuf;
}

// =============================================================================
// (3) trimming and tail silence
// =============================================================================

// Source/MainComponent.cpp:824-861  trimLatency (planar, zero padded to originalLength).
// out: numCh channel pointers with originalLength frames each.  Returns framesToCopy.
ORC_API int orc_trim_latency(const float* const* captured, int numCh, int capturedFrames,
                             int latencySamples, int originalLength, float* const* out) {
    const int latencyFrames = latencySamples / numCh;      // truncating int division (:835)
    const int startFrame = latencyFrames;
    int framesToCopy = originalLength;
    if (startFrame + framesToCopy > capturedFrames)
        framesToCopy = std::max(0, capturedFrames - startFrame);
    for (int c = 0; c < numCh; ++c)
        for (int i = 0; i < originalLength; ++i) out[c][i] = 0.0f;
    if (framesToCopy > 0 && startFrame >= 0) {
        for (int c = 0; c < numCh; ++c)
            std::memcpy(out[c], captured[c] + startFrame, sizeof(float) * (size_t) framesToCopy);
        return framesToCopy;
    }
    return 0;      // nothing copied (the reference returns only the buffer; the count is this wrapper's)
}

// AudioProcessingService.swift:681-703  trimLatency (interleaved, NO padding).
// Returns the number of samples written to out (out must hold sourceFrames*channelCount).
ORC_API long long orc_trim_latency_swift(const float* captured, long long count, long long latencySamples,
                                         long long sourceFrames, int channelCount, float* out) {
    const long long start = latencySamples;
    const long long want = sourceFrames * channelCount;
    if (!(start < count)) {                       // guard failed: prefix(desired)
        long long n = std::min(want, count);
        if (n < 0) n = 0;
        std::memcpy(out, captured, sizeof(float) * (size_t) n);
        return n;
    }
    const long long end = std::min(start + want, count);
    const long long n = end - start;
    std::memcpy(out, captured + start, sizeof(float) * (size_t) n);
    return n;
}

// Source/MainComponent.cpp:863-882  isReverbTailBelowNoiseFloor (RMS based).
ORC_API int orc_tail_below_floor(const float* const* ch, int numCh, int numFrames,
                                 int hasNf, float nfDb, float marginPct) {
    float rms = rms_of(Planar{ch, numCh, numFrames});
    float windowDb = 20.0f * std::log10(std::max(rms, 1e-10f));
    float thresholdDb = nf_threshold_db(hasNf, nfDb, marginPct);
    return windowDb < thresholdDb;
}

// AudioProcessingService.swift:710-737  isReverbTailBelowNoiseFloor (peak based).
ORC_API int orc_tail_below_floor_swift(const float* window, long long n, int hasNf, float nfDb, float marginPct) {
    float maxAbs = 0.0f;
    for (long long i = 0; i < n; ++i) { float a = std::fabs(window[i]); if (a > maxAbs) maxAbs = a; }
    if (!hasNf) return maxAbs < 0.0001f;
    float thresholdDb = nfDb + (nfDb * marginPct / 100.0f);
    float maxDb = maxAbs > 0 ? 20.0f * std::log10(maxAbs) : -160.0f;
    return maxDb < thresholdDb;
}

// Offline restatement of the reverb-mode stop loop.
//   Swift: AudioProcessingService.swift:423-453 (min length, then every 50 ms test the
//          last 100 ms, 3 consecutive, reset on sound);  C++ intent: claude.md:346-367
//          (last 2048 frames every buffer).
// Poll i (i >= 0) happens when e_i = startFrame + (i+1)*hop frames have been captured
// (e_i <= numFrames).  The window is the last `window` frames [e_i-window, e_i); a poll
// with e_i < window is skipped without touching the counter (Swift :441).  mode 0 = C++
// RMS predicate, mode 1 = Swift peak predicate (all channels of the frames, as the
// interleaved window holds them).  flags (optional) gets -1 skipped / 0 / 1 per poll.
// Returns the stop frame e_i of the poll that reaches `required`, or -1.
ORC_API long long orc_tail_scan(const float* const* ch, int numCh, long long numFrames,
                                long long startFrame, int window, int hop, int required, int mode,
                                int hasNf, float nfDb, float marginPct, int* flags, int* numPolls) {
    int consecutive = 0, polls = 0;
    long long stop = -1;
    std::vector<const float*> w((size_t) numCh);
    for (long long i = 0;; ++i) {
        long long e = startFrame + (i + 1) * (long long) hop;
        if (e > numFrames) break;
        int f = -1;
        if (e >= window) {
            for (int c = 0; c < numCh; ++c) w[(size_t) c] = ch[c] + (e - window);
            if (mode == 0) {
                f = orc_tail_below_floor(w.data(), numCh, window, hasNf, nfDb, marginPct);
            } else {
                float maxAbs = 0.0f;
                for (int k = 0; k < window; ++k)
                    for (int c = 0; c < numCh; ++c) { float a = std::fabs(w[(size_t) c][k]); if (a > maxAbs) maxAbs = a; }
                if (!hasNf) f = maxAbs < 0.0001f;
                else {
                    float thr = nfDb + (nfDb * marginPct / 100.0f);
                    float db = maxAbs > 0 ? 20.0f * std::log10(maxAbs) : -160.0f;
                    f = db < thr;
                }
            }
            consecutive = f ? consecutive + 1 : 0;
        }
        if (flags) flags[polls] = f;
        ++polls;
        if (stop < 0 && consecutive >= required) { stop = e; if (!flags) break; }
    }
    if (numPolls) *numPolls = polls;
    return stop;
}

// Source/MainComponent.cpp:884-902  removeDCOffset (float sequential accumulator), in place.
ORC_API void orc_remove_dc_offset(float* const* ch, int numCh, int numFrames) {
    for (int c = 0; c < numCh; ++c) {
        float* d = ch[c];
        float sum = 0.0f;
        for (int i = 0; i < numFrames; ++i) sum += d[i];
        float dc = sum / numFrames;
        for (int i = 0; i < numFrames; ++i) d[i] -= dc;
    }
}

// =============================================================================
// stimuli (synthetic-input shapes)
// =============================================================================
// Source/MainComponent.cpp:934-945  generateImpulse: 0.9 on sample 0 of every channel.
ORC_API void orc_generate_impulse(float* const* ch, int numCh, int numFrames) {
    for (int c = 0; c < numCh; ++c) {
        for (int i = 0; i < numFrames; ++i) ch[c][i] = 0.0f;
        if (numFrames > 0) ch[c][0] = 0.9f;
    }
}
// Source/MainComponent.cpp:907-932  generateSineWave: float phase, wrap at 2*pi. Returns new phase.
ORC_API float orc_generate_sine(float* const* ch, int numCh, int numSamples, float freq, float sampleRate, float phase0) {
    const float amplitude = 0.5f;
    const float twoPi = 2.0f * 3.14159265358979323846f;
    const float inc = (freq * 2.0f * 3.14159265358979323846f) / sampleRate;
    for (int c = 0; c < numCh; ++c) {
        float phase = phase0;
        for (int i = 0; i < numSamples; ++i) {
            ch[c][i] = amplitude * std::sin(phase);
            phase += inc;
            if (phase >= twoPi) phase -= twoPi;
        }
    }
    float p = phase0 + inc * numSamples;
    if (p >= twoPi) p -= twoPi;
    return p;
}

// Source/MainComponent.cpp:141-167  the audio callback's sine: same samples, sinePhase is the chain itself. Returns new phase.
ORC_API float orc_generate_sine_callback(float* const* ch, int numCh, int numSamples, float freq, float sampleRate, float phase0) {
    const float amplitude = 0.5f;
    const float twoPi = 2.0f * 3.14159265358979323846f;
    const float inc = (freq * 2.0f * 3.14159265358979323846f) / sampleRate;
    float phase = phase0;
    for (int i = 0; i < numSamples; ++i) {
        const float sample = amplitude * std::sin(phase);
        for (int c = 0; c < numCh; ++c) ch[c][i] = sample;
        phase += inc;
        if (phase >= twoPi) phase -= twoPi;
    }
    return phase;
}
// Swift SineWaveGenerator.swift:35-59: double phase, interleaved, Float(sin(phase)) * amplitude. Returns new phase.
ORC_API double orc_generate_sine_swift(float* buffer, int frameCount, int channelCount, double freq, double sampleRate,
                                       float amplitude, double phase) {
    const double pi = 3.14159265358979323846;
    const double inc = 2.0 * pi * freq / sampleRate;
    for (int f = 0; f < frameCount; ++f) {
        const float sample = (float) std::sin(phase) * amplitude;
        for (int c = 0; c < channelCount; ++c) buffer[(size_t) f * channelCount + c] = sample;
        phase += inc;
        if (phase >= 2.0 * pi) phase -= 2.0 * pi;
    }
    return phase;
}

// =============================================================================
// (1) sample-rate conversion -- juce::Interpolators  [JUCE-recall, JUCE 8.0.10
// juce_audio_basics/utilities/juce_GenericInterpolator.h, juce_Interpolators.h,
// juce_LagrangeInterpolator.cpp, juce_WindowedSincInterpolator.cpp].  The
// reference links the module (JuceLibraryCode/JuceHeader.h:16) but has no call
// site; SURVEY.md Appendix A records the algorithm.
// =============================================================================
namespace {

constexpr int kSincTableSize = 10001;   // 100 zero crossings x 100 points + 1

// Stand-in for WindowedSincTraits::lookupTable: sinc(x) * Hann(x/100), x = i/100,
// evaluated in double and rounded to float, exact zeros at integer crossings.
void make_sinc_table(float* t) {
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < kSincTableSize; ++i) {
        if (i == 0) { t[i] = 1.0f; continue; }
        if (i % 100 == 0) { t[i] = 0.0f; continue; }
        double x = (double) i / 100.0;
        double s = std::sin(pi * x) / (pi * x);
        double w = 0.5 * (1.0 + std::cos(pi * x / 100.0));
        t[i] = (float) (s * w);
    }
}

struct SincTraits {
    static constexpr int memory = 200;
    static constexpr float latency = 100.0f;
    const float* table;
    float value(const float* inputs, float offset, int indexBuffer) const {
        const int numCrossings = 100;
        const float floatCrossings = (float) numCrossings;
        float result = 0.0f;
        int samplePosition = indexBuffer;
        float firstFrac = 0.0f;
        float lastSincPosition = -1.0f;
        int index = 0, sign = -1;
        for (int i = -numCrossings; i <= numCrossings; ++i) {
            float sincPosition = (1.0f - offset) + (float) i;
            if (i == -numCrossings || (sincPosition >= 0 && lastSincPosition < 0)) {
                float indexFloat = (sincPosition >= 0.f ? sincPosition : -sincPosition) * 100.0f;
                float indexFloored = std::floor(indexFloat);
                index = (int) indexFloored;
                firstFrac = indexFloat - indexFloored;
                sign = (sincPosition < 0 ? -1 : 1);
            }
            if (sincPosition == 0.0f) {
                result += inputs[samplePosition];
            } else if (sincPosition < floatCrossings && sincPosition > -floatCrossings) {
                float v1 = table[index], v2 = table[index + 1];
                float w = v1 + (firstFrac * (v2 - v1));
                result += inputs[samplePosition] * w;
            }
            if (++samplePosition == numCrossings * 2) samplePosition = 0;
            lastSincPosition = sincPosition;
            index += 100 * sign;
        }
        return result;
    }
};

template <int k> struct LagHelper { static void calc(float& a, float b) { a *= b * (1.0f / k); } };
template <> struct LagHelper<0> { static void calc(float&, float) {} };
template <int k> float lagCoef(float input, float offset) {
    LagHelper<0 - k>::calc(input, -2.0f - offset);
    LagHelper<1 - k>::calc(input, -1.0f - offset);
    LagHelper<2 - k>::calc(input, 0.0f - offset);
    LagHelper<3 - k>::calc(input, 1.0f - offset);
    LagHelper<4 - k>::calc(input, 2.0f - offset);
    return input;
}
struct LagrangeTraits {
    static constexpr int memory = 5;
    static constexpr float latency = 2.0f;
    float value(const float* inputs, float offset, int index) const {
        float result = 0.0f;
        result += lagCoef<0>(inputs[index], offset); if (++index == 5) index = 0;
        result += lagCoef<1>(inputs[index], offset); if (++index == 5) index = 0;
        result += lagCoef<2>(inputs[index], offset); if (++index == 5) index = 0;
        result += lagCoef<3>(inputs[index], offset); if (++index == 5) index = 0;
        result += lagCoef<4>(inputs[index], offset);
        return result;
    }
};
struct CatmullRomTraits {
    static constexpr int memory = 4;
    static constexpr float latency = 2.0f;
    float value(const float* inputs, float offset, int index) const {
        float y0 = inputs[index]; if (++index == 4) index = 0;
        float y1 = inputs[index]; if (++index == 4) index = 0;
        float y2 = inputs[index]; if (++index == 4) index = 0;
        float y3 = inputs[index];
        float halfY0 = 0.5f * y0, halfY3 = 0.5f * y3;
        return y1 + offset * ((0.5f * y2 - halfY0)
                 + (offset * (((y0 + 2.0f * y2) - (halfY3 + 2.5f * y1))
                 + (offset * ((halfY3 + 1.5f * y1) - (halfY0 + 1.5f * y2))))));
    }
};
struct LinearTraits {
    static constexpr int memory = 2;
    static constexpr float latency = 1.0f;
    float value(const float* inputs, float offset, int index) const {
        float y0 = inputs[index];
        float y1 = inputs[index == 0 ? 1 : 0];
        return y1 * offset + y0 * (1.0f - offset);
    }
};
struct ZohTraits {
    static constexpr int memory = 1;
    static constexpr float latency = 0.0f;
    float value(const float* inputs, float, int) const { return inputs[0]; }
};

struct InterpBase {
    virtual ~InterpBase() {}
    virtual void reset() = 0;
    virtual int process(double ratio, const float* in, float* out, int numOut, float gain, bool adding) = 0;
    virtual int processWrap(double ratio, const float* in, float* out, int numOut, int avail, int wrap, float gain, bool adding) = 0;
    virtual float latency() const = 0;
    virtual double pos() const = 0;
};

// juce_GenericInterpolator.h: ring of the last `memory` inputs, indexBuffer = next write
// (= oldest), double subSamplePos starting at 1.0.
template <class Traits>
struct Generic : InterpBase {
    Traits traits;
    float last[Traits::memory];
    int indexBuffer = 0;
    double subSamplePos = 1.0;
    explicit Generic(Traits t) : traits(t) { reset(); }
    void reset() override {
        indexBuffer = 0; subSamplePos = 1.0;
        for (int i = 0; i < Traits::memory; ++i) last[i] = 0.0f;
    }
    void push(float v) { last[indexBuffer] = v; if (++indexBuffer == Traits::memory) indexBuffer = 0; }
    int process(double ratio, const float* in, float* out, int numOut, float gain, bool adding) override {
        double pos = subSamplePos;
        int numUsed = 0;
        while (numOut > 0) {
            while (pos >= 1.0) { push(in[numUsed++]); pos -= 1.0; }
            float v = traits.value(last, (float) pos, indexBuffer);
            if (adding) *out++ += gain * v; else *out++ = v;
            pos += ratio;
            --numOut;
        }
        subSamplePos = pos;
        return numUsed;
    }
    int processWrap(double ratio, const float* input, float* out, int numOut, int avail, int wrap, float gain, bool adding) override {
        const float* originalIn = input;
        double pos = subSamplePos;
        bool exceeded = false;
        while (numOut > 0) {
            while (pos >= 1.0) {
                if (exceeded) push(0.0f);
                else {
                    push(*input++);
                    if (--avail <= 0) {
                        if (wrap > 0) { input -= wrap; avail += wrap; }
                        else exceeded = true;
                    }
                }
                pos -= 1.0;
            }
            float v = traits.value(last, (float) pos, indexBuffer);
            if (adding) *out++ += gain * v; else *out++ = v;
            pos += ratio;
            --numOut;
        }
        subSamplePos = pos;
        if (wrap == 0) return (int) (input - originalIn);
        return ((int) (input - originalIn) + wrap) % wrap;
    }
    float latency() const override { return Traits::latency; }
    double pos() const override { return subSamplePos; }
};

struct Handle {
    InterpBase* impl = nullptr;
    std::vector<float> table;     // WindowedSinc lookup table owned by the handle
};

}  // namespace

ORC_API void orc_sinc_table(float* out10001) { make_sinc_table(out10001); }

// kind: 0 WindowedSinc, 1 Lagrange, 2 CatmullRom, 3 Linear, 4 ZeroOrderHold.
// table10001: optional replacement for the WindowedSinc lookup table (nullptr = stand-in).
ORC_API void* orc_interp_create(int kind, const float* table10001) {
    Handle* h = new Handle();
    switch (kind) {
        case 0: {
            h->table.resize(kSincTableSize + 1);
            if (table10001) std::memcpy(h->table.data(), table10001, sizeof(float) * kSincTableSize);
            else make_sinc_table(h->table.data());
            h->table[kSincTableSize] = 0.0f;
            h->impl = new Generic<SincTraits>(SincTraits{h->table.data()});
            break;
        }
        case 1: h->impl = new Generic<LagrangeTraits>(LagrangeTraits{}); break;
        case 2: h->impl = new Generic<CatmullRomTraits>(CatmullRomTraits{}); break;
        case 3: h->impl = new Generic<LinearTraits>(LinearTraits{}); break;
        case 4: h->impl = new Generic<ZohTraits>(ZohTraits{}); break;
        default: delete h; return nullptr;
    }
    return h;
}
ORC_API void orc_interp_destroy(void* p) { Handle* h = (Handle*) p; if (h) { delete h->impl; delete h; } }
ORC_API void orc_interp_reset(void* p) { ((Handle*) p)->impl->reset(); }
ORC_API float orc_interp_latency(void* p) { return ((Handle*) p)->impl->latency(); }
ORC_API double orc_interp_pos(void* p) { return ((Handle*) p)->impl->pos(); }
ORC_API int orc_interp_process(void* p, double ratio, const float* in, float* out, int numOut) {
    return ((Handle*) p)->impl->process(ratio, in, out, numOut, 1.0f, false);
}
ORC_API int orc_interp_process_adding(void* p, double ratio, const float* in, float* out, int numOut, float gain) {
    return ((Handle*) p)->impl->process(ratio, in, out, numOut, gain, true);
}
ORC_API int orc_interp_process_wrap(void* p, double ratio, const float* in, float* out, int numOut, int avail, int wrap) {
    return ((Handle*) p)->impl->processWrap(ratio, in, out, numOut, avail, wrap, 1.0f, false);
}

// Whole-channel conversion from reset state: numIn inputs available, zeros after that
// (the 6-argument process() with wrapAround = 0).  Returns inputs consumed.
ORC_API int orc_resample_channel(int kind, const float* table10001, double ratio,
                                 const float* in, int numIn, float* out, int numOut) {
    void* h = orc_interp_create(kind, table10001);
    if (!h) return -1;
    int used = 0;
    if (numOut > 0) {
        if (numIn > 0) used = orc_interp_process_wrap(h, ratio, in, out, numOut, numIn, 0);
        else { float z = 0.0f; used = orc_interp_process_wrap(h, ratio, &z, out, numOut, 1, 0) - 1; }
    }
    orc_interp_destroy(h);
    return used;
}

// The same conversion with every product and the sum in double (WindowedSinc only): the float weights, the float inputs and the
// position chain are the interpolator's own, only the 200-term accumulation is exact to ~1e-16.  This is NOT what the reference
// computes; it is the yardstick that shows how far the reference's own sequential float sum is from the exact value (at 0 dBFS
// noise: up to ~1.5 x 2^-20) and how far the tensor-core kernel is (tests/test_gpu_resample.py).
ORC_API int orc_resample_channel_exact(const float* table10001, double ratio, const float* in, int numIn, double* out, int numOut) {
    std::vector<float> table(kSincTableSize + 1, 0.0f);
    if (table10001) std::memcpy(table.data(), table10001, sizeof(float) * kSincTableSize); else make_sinc_table(table.data());
    const int M = 200;
    float last[M]; for (int i = 0; i < M; ++i) last[i] = 0.0f;
    int indexBuffer = 0, used = 0;
    double pos = 1.0;
    for (int n = 0; n < numOut; ++n) {
        while (pos >= 1.0) { last[indexBuffer] = used < numIn ? in[used] : 0.0f; ++used; if (++indexBuffer == M) indexBuffer = 0; pos -= 1.0; }
        const float offset = (float) pos;
        // SincTraits::value with a double accumulator
        double result = 0.0;
        int samplePosition = indexBuffer; float firstFrac = 0.0f, lastSincPosition = -1.0f; int index = 0, sign = -1;
        for (int i = -100; i <= 100; ++i) {
            const float sincPosition = (1.0f - offset) + (float) i;
            if (i == -100 || (sincPosition >= 0 && lastSincPosition < 0)) {
                const float indexFloat = (sincPosition >= 0.f ? sincPosition : -sincPosition) * 100.0f;
                const float indexFloored = std::floor(indexFloat);
                index = (int) indexFloored; firstFrac = indexFloat - indexFloored; sign = (sincPosition < 0 ? -1 : 1);
            }
            if (sincPosition == 0.0f) result += (double) last[samplePosition];
            else if (sincPosition < 100.0f && sincPosition > -100.0f) {
                const float v1 = table[index], v2 = table[index + 1];
                const float w = v1 + (firstFrac * (v2 - v1));
                result += (double) last[samplePosition] * (double) w;
            }
            if (++samplePosition == M) samplePosition = 0;
            lastSincPosition = sincPosition;
            index += 100 * sign;
        }
        out[n] = result;
        pos += ratio;
    }
    return used;
}

// Multi-threaded whole-file conversion used as the CPU baseline: one interpolator
// object per channel, channels statically partitioned over `threads` host threads
// (bench.py launches the threads; this entry converts channels [c0, c1)).
ORC_API void orc_resample_channels(int kind, const float* table10001, double ratio,
                                   const float* const* in, int numIn, float* const* out, int numOut,
                                   int c0, int c1) {
    for (int c = c0; c < c1; ++c) orc_resample_channel(kind, table10001, ratio, in[c], numIn, out[c], numOut);
}

// =============================================================================
// bounded-lag cross-correlation with argmax (defined by this project; the reference
// only peak-picks, LatencyMeasurementService.swift:164 says so).
//   r_c[lag] = sum_i (double) x[i] * (double) y_c[i + lag]      (i ascending, y = 0 outside)
//   scan: channel 0 lags lagMin..lagMax ascending, then channel 1, ...; strict '>' on
//   |r| (double) starting from 0  => earliest lag of the lowest channel wins ties, like
//   findPeakPosition (Source/MainComponent.cpp:950-975).
//   found  <=>  max|r| > threshold * sqrt(sum x^2)   (so an impulse a*delta reduces to
//   |y| > threshold exactly).
// Returns found; writes lag / channel / value.
// =============================================================================
ORC_API int orc_xcorr_peak(const float* const* y, int numCh, int numFrames,
                           const float* x, int stimLen, int lagMin, int lagMax, float threshold,
                           int* outLag, int* outCh, double* outValue) {
    double best = 0.0; int bestLag = 0, bestCh = -1;
    for (int c = 0; c < numCh; ++c) {
        const float* yc = y[c];
        for (int lag = lagMin; lag <= lagMax; ++lag) {
            int i0 = std::max(0, -lag);
            int i1 = std::min(stimLen, numFrames - lag);
            double acc = 0.0;
            for (int i = i0; i < i1; ++i) acc += (double) x[i] * (double) yc[i + lag];
            double a = std::fabs(acc);
            if (a > best) { best = a; bestLag = lag; bestCh = c; }
        }
    }
    double energy = 0.0;
    for (int i = 0; i < stimLen; ++i) energy += (double) x[i] * (double) x[i];
    double norm = std::sqrt(energy);
    if (outLag) *outLag = bestLag;
    if (outCh) *outCh = bestCh;
    if (outValue) *outValue = best;
    return (bestCh >= 0 && best > (double) threshold * norm) ? 1 : 0;
}
// Exact value of one lag (used to check the guard-band path).
ORC_API double orc_xcorr_at(const float* y, int numFrames, const float* x, int stimLen, int lag) {
    int i0 = std::max(0, -lag);
    int i1 = std::min(stimLen, numFrames - lag);
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) acc += (double) x[i] * (double) y[i + lag];
    return acc;
}

// =============================================================================
// (d) deinterleave / format convert  [JUCE-recall: juce_audio_formats
// AudioFormatReader::read + ReadHelper (int PCM left-justified to int32, then
// convertFixedToFloat x 1/0x7fffffff), AudioFormatWriter::writeFromFloatArrays
// (convertFloatsToInts) + WavAudioFormatWriter (top 24 bits, little endian)].
// Reference call sites: Source/MainComponent.cpp:734-739 (read), :784-801 (write);
// Swift AudioProcessingService.swift:361-365, :524-531 (planar<->interleaved).
// =============================================================================
// fmt: 1 = u8 (WAV 8-bit, offset binary), 2 = s16le, 3 = s24le packed, 4 = s32le, 5 = f32le.
// Source interleaved with srcCh channels; destination planar with dstCh channels, channel c
// reads source channel min(c, srcCh-1) (mono -> stereo duplication,
// AudioProcessingService.swift:579-580).
ORC_API void orc_pcm_to_planar(const unsigned char* src, int fmt, int srcCh, long long numFrames,
                               float* const* dst, int dstCh) {
    const float scale = 1.0f / 0x7fffffff;
    for (int c = 0; c < dstCh; ++c) {
        int sc = std::min(c, srcCh - 1);
        for (long long f = 0; f < numFrames; ++f) {
            long long s = f * srcCh + sc;
            float v;
            switch (fmt) {
                case 1: { int32_t i = (int32_t) (((uint32_t) (src[s] - 128)) << 24); v = (float) i * scale; break; }
                case 2: { uint32_t u = (uint32_t) src[2 * s] | ((uint32_t) src[2 * s + 1] << 8);
                          int32_t i = (int32_t) (u << 16); v = (float) i * scale; break; }
                case 3: { uint32_t u = (uint32_t) src[3 * s] | ((uint32_t) src[3 * s + 1] << 8) | ((uint32_t) src[3 * s + 2] << 16);
                          int32_t i = (int32_t) (u << 8); v = (float) i * scale; break; }
                case 4: { int32_t i; std::memcpy(&i, src + 4 * s, 4); v = (float) i * scale; break; }
                default: { std::memcpy(&v, src + 4 * s, 4); break; }
            }
            dst[c][f] = v;
        }
    }
}
// planar float -> interleaved 24-bit little-endian PCM (3 bytes per sample).
ORC_API void orc_planar_to_pcm24(const float* const* src, int numCh, long long numFrames, unsigned char* dst) {
    for (long long f = 0; f < numFrames; ++f)
        for (int c = 0; c < numCh; ++c) {
            const double samp = src[c][f];
            int32_t i;
            if (samp <= -1.0) i = std::numeric_limits<int>::min();
            else if (samp >= 1.0) i = std::numeric_limits<int>::max();
            else i = (int32_t) std::nearbyint(std::numeric_limits<int>::max() * samp);   // roundToInt: half to even
            int32_t t = i >> 8;
            unsigned char* d = dst + 3 * (f * numCh + c);
            d[0] = (unsigned char) (t & 0xff); d[1] = (unsigned char) ((t >> 8) & 0xff); d[2] = (unsigned char) ((t >> 16) & 0xff);
        }
}
// planar <-> interleaved float (AudioProcessingService.swift:361-365, :524-531).
ORC_API void orc_interleave(const float* const* src, int numCh, long long numFrames, float* dst) {
    for (long long f = 0; f < numFrames; ++f) for (int c = 0; c < numCh; ++c) dst[f * numCh + c] = src[c][f];
}
ORC_API void orc_deinterleave(const float* src, int numCh, long long numFrames, float* const* dst) {
    for (long long f = 0; f < numFrames; ++f) for (int c = 0; c < numCh; ++c) dst[c][f] = src[f * numCh + c];
}

// =============================================================================
// juce::ResamplingAudioSource arithmetic [JUCE-recall, SURVEY.md Appendix A.3]:
// linear interpolation with a double position + 2nd-order Butterworth low-pass
// (double state) on the input when ratio > 1.0001, on the output when ratio < 0.9999.
// Single channel, whole buffer from reset state.  Listed in SURVEY 8(f) rank 4.
// =============================================================================
ORC_API void orc_resampling_source_coeffs(double ratio, double* c6) {
    const double pi = 3.14159265358979323846;
    double r = ratio > 1.0 ? 0.5 / ratio : 0.5 * ratio;
    double n = 1.0 / std::tan(pi * std::max(0.001, r));
    double nSquared = n * n;
    double c1 = 1.0 / (1.0 + std::sqrt(2.0) * n + nSquared);
    c6[0] = c1; c6[1] = c1 * 2.0; c6[2] = c1; c6[3] = 1.0;
    c6[4] = c1 * 2.0 * (1.0 - nSquared); c6[5] = c1 * (1.0 - std::sqrt(2.0) * n + nSquared);
}

// ---- the whole class [JUCE-recall: JUCE 8.0.10 juce_audio_basics/sources/juce_ResamplingAudioSource.cpp] ---------------
// State and control flow of getNextAudioBlock restated member for member: ring buffer of pulled input (pre-filtered when
// ratio > 1.0001), double subSampleOffset, float lerp `src[pos] + alpha * (src[next] - src[pos])`, post-filter when
// ratio < 0.9999, filter states "stoked" with the last outputs when no filter runs.  The input AudioSource is the caller's
// planar arrays with a read cursor; past their end it delivers zeros (what AudioFormatReaderSource does past the end of a
// file).  applyFilter's JUCE_INTEL branch (outputs within +-1e-8 are flushed to 0) is kept: north_star runs the
// reference on x86 Linux.  PARITY UNPINNED (no JUCE here, no call site in the reference).
namespace {
struct RasFilterState { double x1 = 0, x2 = 0, y1 = 0, y2 = 0; };
struct Ras {
    int numChannels = 0;
    double ratio = 1.0, lastRatio = 1.0;
    double coefficients[6] = {0, 0, 0, 0, 0, 0};
    double subSampleOffset = 0.0;
    int bufferPos = 0, sampsInBuffer = 0, bufferSize = 0;
    std::vector<std::vector<float>> buffer;
    std::vector<RasFilterState> filterStates;
    bool intel = true;

    void setSize(int n, bool keep) {
        for (auto& b : buffer) { if (!keep) b.assign((size_t) n, 0.0f); else b.resize((size_t) n, 0.0f); }
        bufferSize = n;
    }
    void createLowPass(double frequencyRatio) {
        double c[6]; orc_resampling_source_coeffs(frequencyRatio, c);
        const double a = 1.0 / c[3];                                    // setFilterCoefficients
        coefficients[0] = c[0] * a; coefficients[1] = c[1] * a; coefficients[2] = c[2] * a; coefficients[3] = c[3];
        coefficients[4] = c[4] * a; coefficients[5] = c[5] * a;
    }
    void resetFilters() { for (auto& f : filterStates) f = RasFilterState(); }
    void flushBuffers() { for (auto& b : buffer) std::fill(b.begin(), b.end(), 0.0f); bufferPos = 0; sampsInBuffer = 0; subSampleOffset = 0.0; resetFilters(); }
    void prepareToPlay(int samplesPerBlockExpected) {
        const int scaledBlockSize = (int) std::lrint(samplesPerBlockExpected * ratio);        // roundToInt
        buffer.assign((size_t) numChannels, std::vector<float>());
        setSize(scaledBlockSize + 32, false);
        filterStates.assign((size_t) numChannels, RasFilterState());
        createLowPass(ratio);
        flushBuffers();
    }
    void applyFilter(float* samples, int num, RasFilterState& fs) const {
        while (--num >= 0) {
            const double in = *samples;
            double out = coefficients[0] * in + coefficients[1] * fs.x1 + coefficients[2] * fs.x2
                         - coefficients[4] * fs.y1 - coefficients[5] * fs.y2;
            if (intel && !(out < -1.0e-8 || out > 1.0e-8)) out = 0;
            fs.x2 = fs.x1; fs.x1 = in; fs.y2 = fs.y1; fs.y1 = out;
            *samples++ = (float) out;
        }
    }
    // returns the number of samples pulled from the input source
    long long getNextAudioBlock(const float* const* in, long long inTotal, long long cursor, float* const* out, int numSamples) {
        const long long cursor0 = cursor;
        const double localRatio = ratio;
        if (lastRatio != localRatio) { createLowPass(localRatio); lastRatio = localRatio; }
        const int sampsNeeded = (int) std::lrint(numSamples * localRatio) + 3;
        if (bufferSize < sampsNeeded + 8) {
            bufferPos %= bufferSize;
            setSize(sampsNeeded + 32, true);
        }
        bufferPos %= bufferSize;
        int endOfBufferPos = bufferPos + sampsInBuffer;
        const int channelsToProcess = numChannels;
        while (sampsNeeded > sampsInBuffer) {
            endOfBufferPos %= bufferSize;
            const int numToDo = std::min(sampsNeeded - sampsInBuffer, bufferSize - endOfBufferPos);
            for (int c = 0; c < channelsToProcess; ++c)                                       // input->getNextAudioBlock(readInfo)
                for (int i = 0; i < numToDo; ++i) {
                    const long long s = cursor + i;
                    buffer[(size_t) c][(size_t) (endOfBufferPos + i)] = s < inTotal ? in[c][s] : 0.0f;
                }
            cursor += numToDo;
            if (localRatio > 1.0001)
                for (int i = channelsToProcess; --i >= 0;) applyFilter(buffer[(size_t) i].data() + endOfBufferPos, numToDo, filterStates[(size_t) i]);
            sampsInBuffer += numToDo;
            endOfBufferPos += numToDo;
        }
        int nextPos = (bufferPos + 1) % bufferSize;
        for (int m = 0; m < numSamples; ++m) {
            const float alpha = (float) subSampleOffset;
            for (int c = 0; c < channelsToProcess; ++c) {
                const float* src = buffer[(size_t) c].data();
                out[c][m] = src[bufferPos] + alpha * (src[nextPos] - src[bufferPos]);
            }
            subSampleOffset += localRatio;
            while (subSampleOffset >= 1.0) {
                if (++bufferPos >= bufferSize) bufferPos = 0;
                --sampsInBuffer;
                nextPos = (bufferPos + 1) % bufferSize;
                subSampleOffset -= 1.0;
            }
        }
        if (localRatio < 0.9999) {
            for (int i = channelsToProcess; --i >= 0;) applyFilter(out[i], numSamples, filterStates[(size_t) i]);
        } else if (localRatio <= 1.0001 && numSamples > 0) {
            for (int i = channelsToProcess; --i >= 0;) {
                const float* endOfBuffer = out[i] + numSamples - 1;
                RasFilterState& fs = filterStates[(size_t) i];
                if (numSamples > 1) fs.y2 = fs.x2 = *(endOfBuffer - 1);
                else { fs.y2 = fs.y1; fs.x2 = fs.x1; }
                fs.y1 = fs.x1 = *endOfBuffer;
            }
        }
        return cursor - cursor0;
    }
};
}  // namespace

ORC_API void* orc_ras_create(int numChannels, int intelFlush) { Ras* r = new Ras(); r->numChannels = numChannels; r->intel = intelFlush != 0; return r; }
ORC_API void orc_ras_destroy(void* p) { delete (Ras*) p; }
ORC_API void orc_ras_set_ratio(void* p, double samplesInPerOutputSample) { ((Ras*) p)->ratio = std::max(0.0, samplesInPerOutputSample); }
ORC_API void orc_ras_prepare(void* p, int samplesPerBlockExpected) { ((Ras*) p)->prepareToPlay(samplesPerBlockExpected); }
ORC_API void orc_ras_flush(void* p) { ((Ras*) p)->flushBuffers(); }
ORC_API long long orc_ras_get_next_block(void* p, const float* const* in, long long inTotal, long long cursor, float* const* out, int numSamples) {
    return ((Ras*) p)->getNextAudioBlock(in, inTotal, cursor, out, numSamples);
}
// Whole file from reset state in blocks of `block` output samples (the result does not depend on the block size: every input
// sample is pulled and filtered exactly once, in order).
ORC_API void orc_ras_convert(int numCh, const float* const* in, long long nIn, double ratio, float* const* out, long long numOut, int block, int intelFlush) {
    Ras r; r.numChannels = numCh; r.intel = intelFlush != 0; r.ratio = std::max(0.0, ratio); r.lastRatio = r.ratio;
    r.prepareToPlay(block);
    std::vector<float*> o((size_t) numCh);
    long long cursor = 0;
    for (long long done = 0; done < numOut; done += block) {
        const int n = (int) std::min<long long>(block, numOut - done);
        for (int c = 0; c < numCh; ++c) o[(size_t) c] = out[c] + done;
        cursor += r.getNextAudioBlock(in, nIn, cursor, o.data(), n);
    }
}
