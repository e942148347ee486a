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
    std::fclose(g);
    return 0;
}
