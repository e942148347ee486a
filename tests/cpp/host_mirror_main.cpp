// Exercises the C++ host mirror (f9-juce-resampler-studio_b200/host/F9Dsp.hpp) the way MainComponent would:
// reads a planar stereo capture from argv[1] (int32 frames, then 2*frames float32), runs the latency-measurement
// completion, the save path and one interpolator, and writes the results to argv[2] for the Python test to compare
// with the oracle.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "F9Dsp.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t frames = 0, playback = 0;
    if (std::fread(&frames, 4, 1, f) != 1 || std::fread(&playback, 4, 1, f) != 1) return 3;
    f9::AudioBuffer cap(2, frames);
    for (int c = 0; c < 2; ++c)
        if (std::fread(cap.getWritePointer(c), 4, (size_t) frames, f) != (size_t) frames) return 3;
    std::fclose(f);

    f9::Context ctx(0);
    if (!ctx.ok()) { std::fprintf(stderr, "no context: %s\n", ctx.lastError().c_str()); return 4; }
    f9::ProcessingSettings settings;
    f9::BatchDsp dsp(ctx, settings);

    const bool measured = dsp.completeLatencyMeasurement(cap.view());
    const float rms = dsp.calculateRMS(cap.view());
    f9::AudioBuffer trimmed = dsp.processRecording(cap.view(), playback);         // trim + DC removal (default on)
    settings.dcRemovalEnabled = false;
    f9::AudioBuffer trimmedOnly = dsp.processRecording(cap.view(), playback);
    const bool below = dsp.isReverbTailBelowNoiseFloor(trimmedOnly.view());

    f9::Interpolators::WindowedSinc sinc(ctx);
    f9::Interpolators::Lagrange lag(ctx);
    const int numOut = 4000;
    std::vector<float> o1((size_t) numOut), o2((size_t) numOut);
    const int used1 = sinc.process(0.91875, cap.getReadPointer(0), o1.data(), numOut);
    const int used2 = lag.process(0.91875, cap.getReadPointer(1), o2.data(), numOut);

    // juce::ResamplingAudioSource over an AudioSource that reads the capture (zeros past its end), three uneven blocks
    struct MemorySource : f9::AudioSource {
        const f9::AudioBuffer& src; int pos = 0;
        explicit MemorySource(const f9::AudioBuffer& s) : src(s) {}
        void prepareToPlay(int, double) override {}
        void releaseResources() override {}
        void getNextAudioBlock(const f9::AudioSourceChannelInfo& info) override {
            for (int c = 0; c < info.buffer->numChannels; ++c)
                for (int i = 0; i < info.numSamples; ++i)
                    info.buffer->channels[c][info.startSample + i] = pos + i < src.getNumSamples() ? src.getReadPointer(c)[pos + i] : 0.0f;
            pos += info.numSamples;
        }
    };
    const int rasBlocks[3] = {512, 333, 1155};
    f9::AudioBuffer rasOut(2, 2000);
    {
        f9::ResamplingAudioSource ras(ctx, new MemorySource(cap), true, 2);
        ras.setResamplingRatio(96000.0 / 44100.0);
        ras.prepareToPlay(512, 44100.0);
        f9::AudioBufferView v = rasOut.view();
        int at = 0;
        for (int b = 0; b < 3; ++b) {
            f9::AudioSourceChannelInfo info; info.buffer = &v; info.startSample = at; info.numSamples = rasBlocks[b];
            ras.getNextAudioBlock(info);
            at += rasBlocks[b];
        }
    }

    // the stimuli: generateImpulse into a device block, generateSineWave over two consecutive blocks (sinePhase carried)
    f9::AudioBuffer imp(2, 256), sine(2, 1024);
    {
        f9::AudioBufferView iv = imp.view();
        for (int c = 0; c < 2; ++c) for (int i = 0; i < 256; ++i) iv.channels[c][i] = 7.0f;      // generateImpulse clears first
        dsp.generateImpulse(iv);
        float* half[2] = {sine.getWritePointer(0), sine.getWritePointer(1)};
        f9::AudioBufferView s0{half, 2, 512};
        dsp.generateSineWave(s0, 512);
        float* half2[2] = {sine.getWritePointer(0) + 512, sine.getWritePointer(1) + 512};
        f9::AudioBufferView s1{half2, 2, 512};
        dsp.generateSineWave(s1, 512);
    }

    FILE* g = std::fopen(argv[2], "wb");
    if (!g) return 5;
    int32_t hdr[8] = {measured ? 1 : 0, settings.measuredLatencySamples, below ? 1 : 0, used1, used2, numOut,
                      trimmed.getNumSamples(), settings.getRecordingLength(playback, settings.measuredLatencySamples / 2)};
    std::fwrite(hdr, 4, 8, g);
    float fl[2] = {rms, settings.measuredNoiseFloorDb};
    std::fwrite(fl, 4, 2, g);
    for (int c = 0; c < 2; ++c) std::fwrite(trimmed.getReadPointer(c), 4, (size_t) playback, g);
    for (int c = 0; c < 2; ++c) std::fwrite(trimmedOnly.getReadPointer(c), 4, (size_t) playback, g);
    std::fwrite(o1.data(), 4, (size_t) numOut, g);
    std::fwrite(o2.data(), 4, (size_t) numOut, g);
    for (int c = 0; c < 2; ++c) std::fwrite(rasOut.getReadPointer(c), 4, 2000, g);
    for (int c = 0; c < 2; ++c) std::fwrite(imp.getReadPointer(c), 4, 256, g);
    for (int c = 0; c < 2; ++c) std::fwrite(sine.getReadPointer(c), 4, 1024, g);
    std::fwrite(&dsp.sinePhase, 4, 1, g);
    std::fclose(g);
    return 0;
}
