// F9Dsp.hpp -- header-only C++17 host layer above the C ABI (include/f9dsp.h).
//
// It mirrors the reference's own interface for the DSP path so the MainComponent / AppState job flow can
// call it unchanged: same names, argument meaning and error behaviour (sentinels, no exceptions across the
// boundary).  Reference signatures: Source/MainComponent.h:186-237, Source/AppState.h:183-259,
// juce::Interpolators (JUCE 8.0.10 juce_audio_basics, linked at JuceLibraryCode/JuceHeader.h:16).
// No JUCE types are needed: AudioBufferView is the read/write-pointer subset of juce::AudioBuffer<float>.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "f9dsp.h"

namespace f9 {

// ---- the subset of juce::AudioBuffer<float> the helpers use -------------------------------------------
struct AudioBufferView {
    float* const* channels = nullptr;
    int numChannels = 0;
    int numSamples = 0;
    int getNumChannels() const noexcept { return numChannels; }
    int getNumSamples() const noexcept { return numSamples; }
    const float* getReadPointer(int ch) const noexcept { return channels[ch]; }
    float* getWritePointer(int ch) const noexcept { return channels[ch]; }
    const float* const* getArrayOfReadPointers() const noexcept { return channels; }
};

// Owning planar buffer (what trimLatency returns by value, Source/MainComponent.cpp:848-860).
class AudioBuffer {
public:
    AudioBuffer() = default;
    AudioBuffer(int numChannels, int numSamples) { setSize(numChannels, numSamples); }
    void setSize(int numChannels, int numSamples) {
        data_.assign((size_t) numChannels * (size_t) numSamples, 0.0f);
        ptrs_.resize((size_t) numChannels);
        for (int c = 0; c < numChannels; ++c) ptrs_[(size_t) c] = data_.data() + (size_t) c * (size_t) numSamples;
        ch_ = numChannels; n_ = numSamples;
    }
    void clear() { std::fill(data_.begin(), data_.end(), 0.0f); }
    int getNumChannels() const noexcept { return ch_; }
    int getNumSamples() const noexcept { return n_; }
    const float* getReadPointer(int c) const noexcept { return ptrs_[(size_t) c]; }
    float* getWritePointer(int c) noexcept { return ptrs_[(size_t) c]; }
    AudioBufferView view() noexcept { return AudioBufferView{ptrs_.data(), ch_, n_}; }
    operator AudioBufferView() noexcept { return view(); }
private:
    std::vector<float> data_;
    std::vector<float*> ptrs_;
    int ch_ = 0, n_ = 0;
};

// ---- BufferSize (Source/AppState.h:10-16) ---------------------------------------------------------------
enum class BufferSize : int { samples128 = 128, samples256 = 256, samples512 = 512, samples1024 = 1024 };

// ---- ProcessingSettings (Source/AppState.h:183-259): the fields the path reads, same types and defaults ---
struct ProcessingSettings {
    double sampleRate = 44100.0;
    BufferSize bufferSize = BufferSize::samples256;
    int    measuredLatencySamples = -1;        // -1 means not measured
    BufferSize lastBufferSizeWhenMeasured = BufferSize::samples256;
    float  measuredNoiseFloorDb = 0.0f;
    bool   hasNoiseFloorMeasurement = false;
    bool   useReverbMode = false;
    float  noiseFloorMarginPercent = 10.0f;
    int    silenceBetweenFilesMs = 150;
    float  thresholdDb = -40.0f;
    bool   trimEnabled = true;
    bool   dcRemovalEnabled = true;

    bool   needsLatencyRemeasurement() const { return f9_needs_latency_remeasurement(measuredLatencySamples, (int) lastBufferSizeWhenMeasured, (int) bufferSize) != 0; }
    double getLatencyInMs() const { return f9_latency_ms(measuredLatencySamples, sampleRate); }
    int    getRecordingLength(int sourceFileSamples, int latencySamples) const { return f9_recording_length(sourceFileSamples, latencySamples); }
    float  getThresholdLinear() const { return f9_threshold_linear(thresholdDb); }
    float  getNoiseFloorThresholdDb() const { return f9_noise_floor_threshold_db(hasNoiseFloorMeasurement ? 1 : 0, measuredNoiseFloorDb, noiseFloorMarginPercent); }
};

// ---- one GPU context per host thread --------------------------------------------------------------------
class Context {
public:
    explicit Context(int device = 0) { status_ = f9_context_create(device, &ctx_); if (status_) error_ = f9_last_error(nullptr); }
    ~Context() { f9_context_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    bool ok() const noexcept { return ctx_ != nullptr; }
    int status() const noexcept { return status_; }
    std::string lastError() const { return ctx_ ? f9_last_error(ctx_) : error_; }
    f9_context* get() const noexcept { return ctx_; }
private:
    f9_context* ctx_ = nullptr;
    int status_ = F9_OK;
    std::string error_;
};

// ---- juce::Interpolators-shaped classes -----------------------------------------------------------------
template <int Kind>
class GenericInterpolator {
public:
    explicit GenericInterpolator(Context& c) { f9_interp_create(c.get(), Kind, &h_); }
    ~GenericInterpolator() { f9_interp_destroy(h_); }
    GenericInterpolator(const GenericInterpolator&) = delete;
    GenericInterpolator& operator=(const GenericInterpolator&) = delete;
    static constexpr float getBaseLatency() noexcept {
        return Kind == F9_WINDOWED_SINC ? 100.0f : Kind == F9_LAGRANGE ? 2.0f : Kind == F9_CATMULL_ROM ? 2.0f : Kind == F9_LINEAR ? 1.0f : 0.0f;
    }
    void reset() noexcept { f9_interp_reset(h_); }
    /** Same contract as juce: returns the number of input samples consumed (a negative F9_ERR_* on failure). */
    int process(double speedRatio, const float* inputSamples, float* outputSamples, int numOutputSamplesToProduce) noexcept {
        return f9_interp_process(h_, speedRatio, inputSamples, outputSamples, numOutputSamplesToProduce);
    }
    int process(double speedRatio, const float* inputSamples, float* outputSamples, int numOutputSamplesToProduce,
                int numInputSamplesAvailable, int wrapAround) noexcept {
        return f9_interp_process_wrap(h_, speedRatio, inputSamples, outputSamples, numOutputSamplesToProduce, numInputSamplesAvailable, wrapAround);
    }
    int processAdding(double speedRatio, const float* inputSamples, float* outputSamples, int numOutputSamplesToProduce, float gain) noexcept {
        return f9_interp_process_adding(h_, speedRatio, inputSamples, outputSamples, numOutputSamplesToProduce, gain);
    }
private:
    f9_interp* h_ = nullptr;
};
struct Interpolators {
    using WindowedSinc  = GenericInterpolator<F9_WINDOWED_SINC>;
    using Lagrange      = GenericInterpolator<F9_LAGRANGE>;
    using CatmullRom    = GenericInterpolator<F9_CATMULL_ROM>;
    using Linear        = GenericInterpolator<F9_LINEAR>;
    using ZeroOrderHold = GenericInterpolator<F9_ZERO_ORDER_HOLD>;
};

// ---- juce::AudioSource / juce::ResamplingAudioSource (JUCE 8.0.10 juce_audio_basics/sources), GPU backed -----------------
struct AudioSourceChannelInfo {
    AudioBufferView* buffer = nullptr;
    int startSample = 0, numSamples = 0;
};
class AudioSource {
public:
    virtual ~AudioSource() = default;
    virtual void prepareToPlay(int samplesPerBlockExpected, double sampleRate) = 0;
    virtual void releaseResources() = 0;
    virtual void getNextAudioBlock(const AudioSourceChannelInfo& bufferToFill) = 0;
};
/** Same calls as juce::ResamplingAudioSource.  The input source is pulled on the host exactly as JUCE pulls it
    (round(numSamples * ratio) + 3 samples minus what is still buffered, in one request); filters and interpolation run on
    the GPU behind f9_ras_get_next_audio_block. */
class ResamplingAudioSource : public AudioSource {
public:
    ResamplingAudioSource(Context& c, AudioSource* inputSource, bool deleteInputWhenDeleted, int numChannels = 2)
        : input_(inputSource), own_(deleteInputWhenDeleted), numChannels_(numChannels) { f9_ras_create(c.get(), numChannels, &h_); }
    ~ResamplingAudioSource() override { f9_ras_destroy(h_); if (own_) delete input_; }
    ResamplingAudioSource(const ResamplingAudioSource&) = delete;
    ResamplingAudioSource& operator=(const ResamplingAudioSource&) = delete;
    void setResamplingRatio(double samplesInPerOutputSample) { f9_ras_set_resampling_ratio(h_, samplesInPerOutputSample); }
    double getResamplingRatio() const noexcept { return f9_ras_get_resampling_ratio(h_); }
    void prepareToPlay(int samplesPerBlockExpected, double sampleRate) override {
        const double ratio = getResamplingRatio();
        input_->prepareToPlay((int) std::lrint(samplesPerBlockExpected * ratio), sampleRate * ratio);
        f9_ras_prepare_to_play(h_, samplesPerBlockExpected, sampleRate);
    }
    void flushBuffers() { f9_ras_flush_buffers(h_); }
    void releaseResources() override { input_->releaseResources(); f9_ras_release_resources(h_); }
    void getNextAudioBlock(const AudioSourceChannelInfo& info) override {
        const int channels = std::min(numChannels_, info.buffer->numChannels);
        const int pull = std::max(0, f9_ras_num_samples_to_pull(h_, info.numSamples));
        pulled_.setSize(numChannels_, std::max(pull, 1));
        if (pull > 0) {
            AudioBufferView v = pulled_.view();
            AudioSourceChannelInfo readInfo; readInfo.buffer = &v; readInfo.startSample = 0; readInfo.numSamples = pull;
            input_->getNextAudioBlock(readInfo);
        }
        std::vector<float*> out((size_t) numChannels_, nullptr);
        scratch_.setSize(numChannels_, std::max(info.numSamples, 1));
        for (int c = 0; c < numChannels_; ++c)
            out[(size_t) c] = c < channels ? info.buffer->channels[c] + info.startSample : scratch_.view().channels[c];
        f9_ras_get_next_audio_block(h_, pulled_.view().channels, pull, out.data(), info.numSamples);
    }
private:
    AudioSource* input_ = nullptr;
    bool own_ = false;
    int numChannels_ = 2;
    f9_resampling_source* h_ = nullptr;
    AudioBuffer pulled_, scratch_;
};

// ---- the MainComponent helper set (Source/MainComponent.h:186-237), GPU backed ---------------------------
class BatchDsp {
public:
    BatchDsp(Context& c, ProcessingSettings& s) : settings(s), ctx_(c) {}

    /** MainComponent::trimLatency: latencySamples interleaved, originalLength frames; zero padded. */
    AudioBuffer trimLatency(const AudioBufferView& captured, int latencySamples, int originalLength) {
        AudioBuffer trimmed(captured.numChannels, originalLength);
        AudioBufferView v = trimmed.view();
        f9_trim_latency(ctx_.get(), captured.channels, captured.numChannels, captured.numSamples, latencySamples, originalLength, v.channels, nullptr);
        return trimmed;
    }
    bool isReverbTailBelowNoiseFloor(const AudioBufferView& audioWindow) {
        int below = 0;
        f9_is_reverb_tail_below_noise_floor(ctx_.get(), audioWindow.channels, audioWindow.numChannels, audioWindow.numSamples,
                                            settings.hasNoiseFloorMeasurement ? 1 : 0, settings.measuredNoiseFloorDb, settings.noiseFloorMarginPercent, &below);
        return below != 0;
    }
    void removeDCOffset(const AudioBufferView& buffer) { f9_remove_dc_offset(ctx_.get(), buffer.channels, buffer.numChannels, buffer.numSamples); }
    /** MainComponent::findPeakPosition: frame index, -1 when no peak above threshold (also on failure, as the
        reference's caller treats -1 as "could not detect impulse", Source/MainComponent.cpp:287-290). */
    int findPeakPosition(const AudioBufferView& buffer, float threshold) {
        int pos = -1;
        if (f9_find_peak_position(ctx_.get(), buffer.channels, buffer.numChannels, buffer.numSamples, threshold, &pos) != F9_OK) return -1;
        return pos;
    }
    float calculateNoiseFloorDb(const AudioBufferView& buffer) {
        float db = -120.0f;
        f9_calculate_noise_floor_db(ctx_.get(), buffer.channels, buffer.numChannels, buffer.numSamples, &db);
        return db;
    }
    float calculateRMS(const AudioBufferView& buffer) {
        float rms = 0.0f;
        f9_calculate_rms(ctx_.get(), buffer.channels, buffer.numChannels, buffer.numSamples, &rms);
        return rms;
    }
    /** MainComponent::generateSineWave (Source/MainComponent.cpp:907-932): amplitude 0.5, float phase chain, sinePhase updated as there. */
    void generateSineWave(const AudioBufferView& buffer, int numSamples) {
        f9_generate_sine_wave(ctx_.get(), buffer.channels, buffer.numChannels, numSamples, sineFrequency, (float) settings.sampleRate,
                              0.5f, &sinePhase, 0);
    }
    /** MainComponent::generateImpulse (Source/MainComponent.cpp:934-945): cleared buffer, 0.9 on sample 0 of every channel. */
    void generateImpulse(const AudioBufferView& buffer) { f9_generate_impulse(ctx_.get(), buffer.channels, buffer.numChannels, buffer.numSamples); }
    float sinePhase = 0.0f;                 // Source/MainComponent.h:153-154
    float sineFrequency = 1000.0f;

    /** The body of timerCallback's latency completion (Source/MainComponent.cpp:265-294). */
    bool completeLatencyMeasurement(const AudioBufferView& latencyCaptureBuffer) {
        // findPeakPosition(buffer, 0.1f) and calculateNoiseFloorDb(buffer) from one upload and one read of the capture
        int peak = -1; float noiseFloorDb = -120.0f;
        if (f9_measure_latency(ctx_.get(), latencyCaptureBuffer.channels, latencyCaptureBuffer.numChannels, latencyCaptureBuffer.numSamples,
                               0.1f, &peak, &noiseFloorDb) != F9_OK || peak < 0) return false;
        settings.measuredLatencySamples = peak * 2;              // "Assuming stereo" (:275)
        settings.lastBufferSizeWhenMeasured = settings.bufferSize;
        settings.measuredNoiseFloorDb = noiseFloorDb;
        settings.hasNoiseFloorMeasurement = true;
        return true;
    }
    /** saveCurrentRecording's DSP (Source/MainComponent.cpp:751-769): trim, then DC removal when enabled. */
    AudioBuffer processRecording(const AudioBufferView& recordingBuffer, int playbackFrames) {
        AudioBuffer trimmed = trimLatency(recordingBuffer, settings.measuredLatencySamples, playbackFrames);
        if (settings.dcRemovalEnabled) removeDCOffset(trimmed.view());
        return trimmed;
    }
    ProcessingSettings& settings;
private:
    Context& ctx_;
};

}  // namespace f9
