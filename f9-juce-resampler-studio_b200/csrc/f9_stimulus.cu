// Test stimuli of the latency measurement and the hardware loop test, generated on the device:
//   generateImpulse   Source/MainComponent.cpp:934-945 (Swift sendImpulse, LatencyMeasurementService.swift:130-145)
//   generateSineWave  Source/MainComponent.cpp:907-932 (Swift SineWaveGenerator.swift:35-59)
// The sine's phase is a sequential chain (phase += inc; wrap at 2 pi) in float (C++) or double (Swift): rounding makes it
// non-associative, so one thread walks the chain and writes the phases, then the samples are evaluated in parallel.
#include "f9_internal.cuh"

namespace f9 {
namespace {

__global__ void __launch_bounds__(256)
impulse_kernel(const DevBuf* __restrict__ bufs, float amplitude) {
    const DevBuf B = bufs[blockIdx.z];
    const int ch = blockIdx.y;
    if (ch >= B.numCh) return;
    float* __restrict__ d = const_cast<float*>(B.base) + (long long) ch * B.chStride;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < B.numFrames; i += (long long) gridDim.x * blockDim.x)
        d[i] = (i == 0) ? amplitude : 0.0f;          // buffer.clear(); setSample(ch, 0, amplitude)
}

// phase[i] = phase before sample i; *phase_end = the chain's value after n samples (not what sinePhase becomes: the
// reference updates the member in closed form, see the host side).
template <typename T>
__global__ void sine_phase_kernel(T phase0, T inc, T twoPi, int n, T* __restrict__ phases) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    T phase = phase0;
    for (int i = 0; i < n; ++i) {
        phases[i] = phase;
        phase = phase + inc;                        // one rounding (no contraction possible: nothing to fuse with)
        if (phase >= twoPi) phase = phase - twoPi;
    }
    phases[n] = phase;
}

// C++ form: data[i] = amplitude * std::sin(phase) per planar channel, every channel the same chain.
__global__ void __launch_bounds__(256)
sine_fill_kernel(DevBuf B, const float* __restrict__ phases, float amplitude, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // sinf of the host libm is correctly rounded in all but rare cases; so is the double sine rounded to float
    const float s = (float) sin((double) phases[i]);
    const float v = __fmul_rn(amplitude, s);
    for (int c = 0; c < B.numCh; ++c) const_cast<float*>(B.base)[(long long) c * B.chStride + i] = v;
}
// Swift form: sample = Float(sin(phase)) * amplitude, interleaved, every channel of a frame the same sample.
__global__ void __launch_bounds__(256)
sine_fill_swift_kernel(float* __restrict__ out, int channels, const double* __restrict__ phases, float amplitude, int frames) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= frames) return;
    const float v = __fmul_rn((float) sin(phases[i]), amplitude);
    for (int c = 0; c < channels; ++c) out[(long long) i * channels + c] = v;
}

}  // namespace

cudaError_t launch_impulse(const DevBuf* d_bufs, int n, int maxCh, int maxFrames, float amplitude, cudaStream_t s, long long* launches) {
    if (n <= 0 || maxCh <= 0) return cudaSuccess;
    const int bx = std::max(1, std::min((maxFrames + 255) / 256, 1024));
    impulse_kernel<<<dim3((unsigned) bx, (unsigned) maxCh, (unsigned) n), 256, 0, s>>>(d_bufs, amplitude);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_sine(const DevBuf& buf, float phase0, float inc, float amplitude, int n, float* d_phases /* n + 1 */,
                        cudaStream_t s, long long* launches) {
    if (n <= 0) return cudaSuccess;
    sine_phase_kernel<float><<<1, 32, 0, s>>>(phase0, inc, 2.0f * 3.14159265358979323846f, n, d_phases);
    ++*launches;
    sine_fill_kernel<<<(n + 255) / 256, 256, 0, s>>>(buf, d_phases, amplitude, n);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_sine_swift(float* d_out, int channels, double phase0, double inc, float amplitude, int frames, double* d_phases /* frames + 1 */,
                              cudaStream_t s, long long* launches) {
    if (frames <= 0) return cudaSuccess;
    sine_phase_kernel<double><<<1, 32, 0, s>>>(phase0, inc, 2.0 * 3.14159265358979323846, frames, d_phases);
    ++*launches;
    sine_fill_swift_kernel<<<(frames + 255) / 256, 256, 0, s>>>(d_out, channels, d_phases, amplitude, frames);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace f9
