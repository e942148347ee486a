// Trim, DC removal and the deinterleave / format-convert stage.  Pure streaming copies: 4 B read + 4 B written
// per sample (3 + 4 for 24-bit PCM), coalesced, grid sized from the data.
#include "f9_internal.cuh"

namespace f9 {
namespace {

constexpr int kThreads = 256;

// ---- trimLatency (Source/MainComponent.cpp:824-861) and removeDCOffset (:884-902), batched ---------------
// grid: (tiles over frames, channel, buffer).  out = zeros, then captured[start .. start+n) when n > 0 && start >= 0; with SUB the
// channel's mean is subtracted on the way (removeDCOffset runs right after trimLatency in the save path, :766-769: fused, the
// trimmed buffer is written once instead of written, read and rewritten).  A thread moves 4 consecutive frames per step: 128-bit
// stores when the destination is aligned, 128-bit loads when the source is too (the latency decides), scalar loads otherwise.
struct TrimGeom { const float* src; float* dst; int n, frames; };
__device__ __forceinline__ TrimGeom trim_geom(const DevBuf& C, const DevBuf& O, int latency, int ch) {
    TrimGeom G;
    const int start = latency / C.numCh;                     // truncating division (:835)
    G.frames = O.numFrames;                                  // originalLength
    G.n = G.frames;
    if (start + G.n > C.numFrames) G.n = max(0, C.numFrames - start);
    if (start < 0) G.n = 0;
    G.src = C.base + (long long) ch * C.chStride + start;
    G.dst = const_cast<float*>(O.base) + (long long) ch * O.chStride;
    return G;
}
// Sum of a channel's partial sums in index order (deterministic); kDcPartials doubles per (buffer, channel).
__device__ __forceinline__ double dc_total(const double* __restrict__ partials, size_t slot) {
    double t = 0.0;
    #pragma unroll 4
    for (int c = 0; c < kDcPartials; ++c) t += partials[slot * kDcPartials + c];
    return t;
}
template <bool SUB>
__global__ void __launch_bounds__(kThreads)
trim_kernel(const DevBuf* __restrict__ cap, const int* __restrict__ latency, const DevBuf* __restrict__ out,
            const double* __restrict__ partials, const int* __restrict__ dcMask, int maxCh) {
    const int b = blockIdx.z, ch = blockIdx.y;
    const DevBuf C = cap[b];
    const DevBuf O = out[b];
    if (ch >= O.numCh) return;
    const TrimGeom G = trim_geom(C, O, latency[b], ch);
    float dc = 0.0f;
    if (SUB && G.frames > 0 && (dcMask == nullptr || dcMask[b] != 0)) dc = (float) dc_total(partials, (size_t) b * maxCh + ch) / (float) G.frames;
    const bool dstVec = (reinterpret_cast<uintptr_t>(G.dst) & 15) == 0, srcVec = (reinterpret_cast<uintptr_t>(G.src) & 15) == 0;
    #pragma unroll 2
    for (int i = (blockIdx.x * kThreads + threadIdx.x) * 4; i < G.frames; i += gridDim.x * kThreads * 4) {
        float4 v;
        if (srcVec && i + 3 < G.n) v = __ldg(reinterpret_cast<const float4*>(G.src + i));
        else {
            v.x = (i < G.n) ? __ldg(G.src + i) : 0.0f;         v.y = (i + 1 < G.n) ? __ldg(G.src + i + 1) : 0.0f;
            v.z = (i + 2 < G.n) ? __ldg(G.src + i + 2) : 0.0f; v.w = (i + 3 < G.n) ? __ldg(G.src + i + 3) : 0.0f;
        }
        if (SUB) { v.x = __fsub_rn(v.x, dc); v.y = __fsub_rn(v.y, dc); v.z = __fsub_rn(v.z, dc); v.w = __fsub_rn(v.w, dc); }
        if (dstVec && i + 3 < G.frames) __stcs(reinterpret_cast<float4*>(G.dst + i), v);
        else {
            G.dst[i] = v.x;
            if (i + 1 < G.frames) G.dst[i + 1] = v.y;
            if (i + 2 < G.frames) G.dst[i + 2] = v.z;
            if (i + 3 < G.frames) G.dst[i + 3] = v.w;
        }
    }
}

// ---- removeDCOffset (Source/MainComponent.cpp:884-902) -------------------------------------------------
// The reference sums sequentially in float; a parallel sum can only match to tolerance.
// Pass 1: kDcPartials CTAs per (buffer, channel), each a deterministic tree sum in double over its slice of the frames.
// Pass 2: fold the partials in index order, subtract (float)(sum) / numFrames.
__device__ __forceinline__ void dc_partial(const float* __restrict__ x, int frames, double* __restrict__ dstPartial) {
    // slice blockIdx.x of kDcPartials, boundaries on multiples of 4 frames
    const int per = ((frames + kDcPartials - 1) / kDcPartials + 3) & ~3;
    const int lo = min(frames, (int) blockIdx.x * per), hi = min(frames, lo + per);
    double s = 0.0;
    if ((reinterpret_cast<uintptr_t>(x + lo) & 15) == 0) {
        const int nv = (hi - lo) >> 2;
        const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + lo);
        #pragma unroll 4
        for (int i = threadIdx.x; i < nv; i += kThreads) { const float4 v = __ldg(xv + i); s += ((double) v.x + (double) v.y) + ((double) v.z + (double) v.w); }
        for (int i = lo + 4 * nv + threadIdx.x; i < hi; i += kThreads) s += (double) __ldg(x + i);
    } else {
        #pragma unroll 4
        for (int i = lo + threadIdx.x; i < hi; i += kThreads) s += (double) __ldg(x + i);
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __shared__ double sh[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
        *dstPartial = t;
    }
}
// Reference order (mode 2): the reference's own accumulator -- one float, the channel's samples added one after the other
// (Source/MainComponent.cpp:892-896).  That chain is sequential by definition, so one thread walks it (loads run ahead of the
// adds); the sum lands in partial 0, the other partials are zero, and (float) total / frames below is the reference's dcOffset
// bit for bit.  About 2 ms for a 10 s channel at 96 kHz, every channel of the batch at the same time.
__device__ __forceinline__ void dc_sequential(const float* __restrict__ x, int frames, double* __restrict__ dstPartials) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0) { dstPartials[blockIdx.x] = 0.0; return; }
    float sum = 0.0f;
    int i = 0;
    for (; i < frames && ((reinterpret_cast<uintptr_t>(x + i) & 15) != 0); ++i) sum = __fadd_rn(sum, __ldg(x + i));
    const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + i);
    const int nv = (frames - i) >> 2;
    int k = 0;
    for (; k + 4 <= nv; k += 4) {                              // four independent loads in flight, then sixteen ordered adds
        const float4 a = __ldg(xv + k), b = __ldg(xv + k + 1), c = __ldg(xv + k + 2), d = __ldg(xv + k + 3);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, a.x), a.y), a.z), a.w);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, b.x), b.y), b.z), b.w);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, c.x), c.y), c.z), c.w);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, d.x), d.y), d.z), d.w);
    }
    for (; k < nv; ++k) { const float4 a = __ldg(xv + k); sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, a.x), a.y), a.z), a.w); }
    for (i += 4 * nv; i < frames; ++i) sum = __fadd_rn(sum, __ldg(x + i));
    dstPartials[0] = (double) sum;
}
// grid: (kDcPartials, channel, buffer)
__global__ void __launch_bounds__(kThreads)
dc_sum_kernel(const DevBuf* __restrict__ bufs, int maxCh, double* __restrict__ partials, int mode) {
    const int b = blockIdx.z, ch = blockIdx.y;
    const DevBuf B = bufs[b];
    if (ch >= B.numCh) return;
    double* dst = partials + ((size_t) b * maxCh + ch) * kDcPartials;
    if (mode == 2) dc_sequential(B.base + (long long) ch * B.chStride, B.numFrames, dst);
    else dc_partial(B.base + (long long) ch * B.chStride, B.numFrames, dst + blockIdx.x);
}
// The same over the region trimLatency copies (the zero padding adds nothing to the sum): the mean of the trimmed buffer
// without materialising it first.
__global__ void __launch_bounds__(kThreads)
dc_sum_src_kernel(const DevBuf* __restrict__ cap, const int* __restrict__ latency, const DevBuf* __restrict__ out, int maxCh,
                  double* __restrict__ partials, const int* __restrict__ dcMask, int dcDefault) {
    const int b = blockIdx.z, ch = blockIdx.y;
    const DevBuf C = cap[b];
    const DevBuf O = out[b];
    if (ch >= O.numCh) return;
    const int mode = dcMask ? dcMask[b] : dcDefault;           // 0 no DC removal, 1 parallel double sum, 2 the reference's sequential float sum
    if (mode == 0) return;
    const TrimGeom G = trim_geom(C, O, latency[b], ch);
    double* dst = partials + ((size_t) b * maxCh + ch) * kDcPartials;
    if (mode == 2) dc_sequential(G.src, G.n, dst);             // the zero padding after the copied region adds nothing to the chain
    else dc_partial(G.src, G.n, dst + blockIdx.x);
}
__global__ void __launch_bounds__(kThreads)
dc_sub_kernel(const DevBuf* __restrict__ bufs, int maxCh, const double* __restrict__ partials) {
    const int b = blockIdx.z, ch = blockIdx.y;
    const DevBuf B = bufs[b];
    if (ch >= B.numCh || B.numFrames <= 0) return;
    const float dc = (float) dc_total(partials, (size_t) b * maxCh + ch) / (float) B.numFrames;
    float* __restrict__ x = const_cast<float*>(B.base) + (long long) ch * B.chStride;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int nv = B.numFrames >> 2;
        float4* __restrict__ xv = reinterpret_cast<float4*>(x);
        #pragma unroll 2
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < nv; i += gridDim.x * kThreads) {
            float4 v = xv[i];
            v.x = __fsub_rn(v.x, dc); v.y = __fsub_rn(v.y, dc); v.z = __fsub_rn(v.z, dc); v.w = __fsub_rn(v.w, dc);
            xv[i] = v;
        }
        for (int i = 4 * nv + blockIdx.x * kThreads + threadIdx.x; i < B.numFrames; i += gridDim.x * kThreads) x[i] = __fsub_rn(x[i], dc);
    } else {
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < B.numFrames; i += gridDim.x * kThreads) x[i] = __fsub_rn(x[i], dc);
    }
}

// ---- PCM -> planar float (JUCE reader: left-justify to int32, * 1/0x7fffffff) ---------------------------
__device__ __forceinline__ float pcm_load(const unsigned char* __restrict__ src, int fmt, long long s) {
    const float scale = 1.0f / 0x7fffffff;
    switch (fmt) {
        case F9_PCM_U8:  { const int v = (int) ((unsigned) (src[s] - 128) << 24); return __fmul_rn((float) v, scale); }
        case F9_PCM_S16LE: { const unsigned u = (unsigned) src[2 * s] | ((unsigned) src[2 * s + 1] << 8);
                             return __fmul_rn((float) (int) (u << 16), scale); }
        case F9_PCM_S24LE: { const unsigned u = (unsigned) src[3 * s] | ((unsigned) src[3 * s + 1] << 8) | ((unsigned) src[3 * s + 2] << 16);
                             return __fmul_rn((float) (int) (u << 8), scale); }
        case F9_PCM_S32LE: { const unsigned u = (unsigned) src[4 * s] | ((unsigned) src[4 * s + 1] << 8) | ((unsigned) src[4 * s + 2] << 16) | ((unsigned) src[4 * s + 3] << 24);
                             return __fmul_rn((float) (int) u, scale); }
        default: { const unsigned u = (unsigned) src[4 * s] | ((unsigned) src[4 * s + 1] << 8) | ((unsigned) src[4 * s + 2] << 16) | ((unsigned) src[4 * s + 3] << 24);
                   return __uint_as_float(u); }
    }
}
// One thread per frame; the CTA's frames are staged through shared memory so both the interleaved read and the
// planar writes are coalesced.
// frames per CTA: sized by the launcher so the staging tile stays under ~40 KB for any channel count
__host__ __device__ inline int cvt_frames(int bytesPerFrame) { int f = 40960 / (bytesPerFrame > 0 ? bytesPerFrame : 1); f = f > 1024 ? 1024 : f; f &= ~31; return f < 32 ? 32 : f; }
__device__ __forceinline__ void pcm_to_planar_tile(unsigned char* raw, const unsigned char* __restrict__ src, int fmt, int srcCh, long long frames,
                                                   float* __restrict__ dst, long long dstStride, int dstCh, int kCvtFrames) {
    const int bps = (fmt == F9_PCM_U8) ? 1 : (fmt == F9_PCM_S16LE) ? 2 : (fmt == F9_PCM_S24LE) ? 3 : 4;
    const long long f0 = (long long) blockIdx.x * kCvtFrames;
    if (f0 >= frames) return;
    const int nf = (int) min((long long) kCvtFrames, frames - f0);
    const long long byte0 = f0 * srcCh * bps;
    const int nbytes = nf * srcCh * bps;
    // coalesced byte copy (32-bit words where the tile start is aligned)
    if ((byte0 & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const unsigned* __restrict__ s4 = reinterpret_cast<const unsigned*>(src + byte0);
        unsigned* r4 = reinterpret_cast<unsigned*>(raw);
        const int nw = nbytes >> 2;
        for (int i = threadIdx.x; i < nw; i += kThreads) r4[i] = __ldg(s4 + i);
        for (int i = (nw << 2) + threadIdx.x; i < nbytes; i += kThreads) raw[i] = src[byte0 + i];
    } else {
        for (int i = threadIdx.x; i < nbytes; i += kThreads) raw[i] = src[byte0 + i];
    }
    __syncthreads();
    for (int c = 0; c < dstCh; ++c) {
        const int sc = min(c, srcCh - 1);       // mono -> stereo duplication (AudioProcessingService.swift:579-580)
        float* __restrict__ d = dst + (long long) c * dstStride + f0;
        for (int f = threadIdx.x; f < nf; f += kThreads) d[f] = pcm_load(raw, fmt, (long long) f * srcCh + sc);
    }
}
__global__ void __launch_bounds__(kThreads)
pcm_to_planar_kernel(const unsigned char* __restrict__ src, int fmt, int srcCh, long long frames,
                     float* __restrict__ dst, long long dstStride, int dstCh, int kCvtFrames) {
    extern __shared__ unsigned char raw[];
    pcm_to_planar_tile(raw, src, fmt, srcCh, frames, dst, dstStride, dstCh, kCvtFrames);
}
// Batched form: blockIdx.y = file; file i reads srcs[i] (srcCh interleaved channels of its destination's frame count) into
// the planar buffer dsts[i].  One launch for a whole batch instead of one per file (6 us each: the per-file form is launch-bound).
__global__ void __launch_bounds__(kThreads)
pcm_to_planar_batch_kernel(const unsigned char* const* __restrict__ srcs, int fmt, int srcCh, const DevBuf* __restrict__ dsts, int kCvtFrames) {
    extern __shared__ unsigned char raw[];
    const DevBuf D = dsts[blockIdx.y];
    pcm_to_planar_tile(raw, srcs[blockIdx.y], fmt, srcCh, D.numFrames, const_cast<float*>(D.base), D.chStride, D.numCh, kCvtFrames);
}

// ---- planar float -> interleaved 24-bit LE (JUCE writer: clip, roundToInt(INT_MAX * (double) x), top 24 bits) ---
__device__ __forceinline__ int float_to_i32(float x) {
    const double samp = (double) x;
    if (samp <= -1.0) return (int) 0x80000000;
    if (samp >= 1.0) return 0x7fffffff;
    return __double2int_rn(__dmul_rn(2147483647.0, samp));
}
__device__ __forceinline__ void planar_to_pcm24_tile(unsigned char* raw, const float* __restrict__ src, long long srcStride, int numCh, long long frames,
                                                     unsigned char* __restrict__ dst, int kCvtFrames) {
    const long long f0 = (long long) blockIdx.x * kCvtFrames;
    if (f0 >= frames) return;
    const int nf = (int) min((long long) kCvtFrames, frames - f0);
    for (int c = 0; c < numCh; ++c) {
        const float* __restrict__ s = src + (long long) c * srcStride + f0;
        for (int f = threadIdx.x; f < nf; f += kThreads) {
            const int t = float_to_i32(__ldg(s + f)) >> 8;
            unsigned char* r = raw + 3 * (f * numCh + c);
            r[0] = (unsigned char) (t & 0xff); r[1] = (unsigned char) ((t >> 8) & 0xff); r[2] = (unsigned char) ((t >> 16) & 0xff);
        }
    }
    __syncthreads();
    const long long byte0 = f0 * numCh * 3;
    const int nbytes = nf * numCh * 3;
    if ((byte0 & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        unsigned* __restrict__ d4 = reinterpret_cast<unsigned*>(dst + byte0);
        const unsigned* r4 = reinterpret_cast<const unsigned*>(raw);
        const int nw = nbytes >> 2;
        for (int i = threadIdx.x; i < nw; i += kThreads) d4[i] = r4[i];
        for (int i = (nw << 2) + threadIdx.x; i < nbytes; i += kThreads) dst[byte0 + i] = raw[i];
    } else {
        for (int i = threadIdx.x; i < nbytes; i += kThreads) dst[byte0 + i] = raw[i];
    }
}
__global__ void __launch_bounds__(kThreads)
planar_to_pcm24_kernel(const float* __restrict__ src, long long srcStride, int numCh, long long frames, unsigned char* __restrict__ dst, int kCvtFrames) {
    extern __shared__ unsigned char raw[];
    planar_to_pcm24_tile(raw, src, srcStride, numCh, frames, dst, kCvtFrames);
}
// Batched form: blockIdx.y = file (its own channel count and length); tiles of kCvtFrames frames, sized for the widest file.
__global__ void __launch_bounds__(kThreads)
planar_to_pcm24_batch_kernel(const DevBuf* __restrict__ srcs, unsigned char* const* __restrict__ dsts, int kCvtFrames) {
    extern __shared__ unsigned char raw[];
    const DevBuf B = srcs[blockIdx.y];
    planar_to_pcm24_tile(raw, B.base, B.chStride, B.numCh, B.numFrames, dsts[blockIdx.y], kCvtFrames);
}

// ---- fast paths of the 24-bit / 16-bit payload (1 or 2 channels, 16-byte aligned planes and payloads) -------------------
// The byte-staged kernels above reach 67 % of the HBM roofline: one-shot CTAs of 1024 frames behind a block barrier, scalar loads,
// three byte stores per sample.  Here a thread owns 16 interleaved samples -- 48 bytes of 24-bit payload, three 16-byte words --
// and everything moves as 128-bit accesses: planes with LDG.128 / STG.128 (a lane's 8 or 16 frames are contiguous), the payload
// through a warp-private 1536-byte slice of shared memory (a lane's three words sit 48 bytes apart: conflict-free; the warp then
// moves the slice as three fully coalesced 512-byte rows).  No block barrier, persistent grid-stride loop.
constexpr int kFastWarps = kThreads / 32;
__device__ __forceinline__ unsigned pk24(int v) { return (unsigned) (v >> 8) & 0xffffffu; }     // top 24 bits of the int32 sample

template <int NCH>
__global__ void __launch_bounds__(kThreads)
pcm24_pack_fast_kernel(const DevBuf* __restrict__ srcs, unsigned char* const* __restrict__ dsts) {
    __shared__ uint4 slice[kFastWarps][96];
    const DevBuf B = srcs[blockIdx.y];
    unsigned char* __restrict__ dst = dsts[blockIdx.y];
    constexpr int FR = 16 / NCH;                                   // frames per thread
    const long long groups = (long long) B.numFrames / FR;         // whole 16-sample groups
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long base = ((long long) blockIdx.x * kFastWarps + warp) * 32; base < groups; base += (long long) gridDim.x * kThreads) {
        const long long g = base + lane;
        unsigned w[12];
        if (g < groups) {
            int t[16];
            if (NCH == 2) {
                const float4* __restrict__ l4 = reinterpret_cast<const float4*>(B.base + g * FR);
                const float4* __restrict__ r4 = reinterpret_cast<const float4*>(B.base + B.chStride + g * FR);
                const float4 a0 = __ldg(l4), a1 = __ldg(l4 + 1), b0 = __ldg(r4), b1 = __ldg(r4 + 1);
                const float L[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, R[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                #pragma unroll
                for (int i = 0; i < 8; ++i) { t[2 * i] = float_to_i32(L[i]); t[2 * i + 1] = float_to_i32(R[i]); }
            } else {
                const float4* __restrict__ x4 = reinterpret_cast<const float4*>(B.base + g * FR);
                #pragma unroll
                for (int i = 0; i < 4; ++i) { const float4 v = __ldg(x4 + i); t[4 * i] = float_to_i32(v.x); t[4 * i + 1] = float_to_i32(v.y); t[4 * i + 2] = float_to_i32(v.z); t[4 * i + 3] = float_to_i32(v.w); }
            }
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned a = pk24(t[4 * q]), b = pk24(t[4 * q + 1]), c = pk24(t[4 * q + 2]), d = pk24(t[4 * q + 3]);
                w[3 * q] = a | (b << 24); w[3 * q + 1] = (b >> 8) | (c << 16); w[3 * q + 2] = (c >> 16) | (d << 8);
            }
            #pragma unroll
            for (int k = 0; k < 3; ++k) slice[warp][3 * lane + k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
        }
        __syncwarp();
        const int valid = (int) min((long long) 32, groups - base) * 3;           // 16-byte words of this warp's row
        uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst + base * 48);
        #pragma unroll
        for (int k = 0; k < 3; ++k) if (lane + 32 * k < valid) __stcs(d4 + lane + 32 * k, slice[warp][lane + 32 * k]);
        __syncwarp();
    }
    // the last frames (fewer than one group): bytes, by the first threads of the file's first CTA
    if (blockIdx.x == 0) {
        const long long f0 = groups * FR;
        const int rem = (int) (B.numFrames - f0) * NCH;
        if ((int) threadIdx.x < rem) {
            const int f = threadIdx.x / NCH, c = threadIdx.x - f * NCH;
            const int v = float_to_i32(__ldg(B.base + (long long) c * B.chStride + f0 + f)) >> 8;
            unsigned char* r = dst + (f0 * NCH + threadIdx.x) * 3;
            r[0] = (unsigned char) (v & 0xff); r[1] = (unsigned char) ((v >> 8) & 0xff); r[2] = (unsigned char) ((v >> 16) & 0xff);
        }
    }
}

// 24-bit (FMT24) or 16-bit interleaved payload of SRC channels -> DST planar channels (DST >= SRC: mono duplicated).
template <int SRC, int DST, bool FMT24>
__global__ void __launch_bounds__(kThreads)
pcm_unpack_fast_kernel(const unsigned char* const* __restrict__ srcs, const DevBuf* __restrict__ dsts) {
    __shared__ uint4 slice[kFastWarps][96];
    const DevBuf D = dsts[blockIdx.y];
    const unsigned char* __restrict__ src = srcs[blockIdx.y];
    constexpr int FR = 16 / SRC;
    const float scale = 1.0f / 0x7fffffff;
    const long long groups = (long long) D.numFrames / FR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* __restrict__ d0 = const_cast<float*>(D.base);
    for (long long base = ((long long) blockIdx.x * kFastWarps + warp) * 32; base < groups; base += (long long) gridDim.x * kThreads) {
        const long long g = base + lane;
        int t[16];
        if (FMT24) {
            const int valid = (int) min((long long) 32, groups - base) * 3;
            const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src + base * 48);
            #pragma unroll
            for (int k = 0; k < 3; ++k) if (lane + 32 * k < valid) slice[warp][lane + 32 * k] = __ldg(s4 + lane + 32 * k);
            __syncwarp();
            if (g < groups) {
                unsigned w[12];
                #pragma unroll
                for (int k = 0; k < 3; ++k) { const uint4 v = slice[warp][3 * lane + k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
                    t[4 * q] = (int) (w0 << 8); t[4 * q + 1] = (int) (((w0 >> 24) | (w1 << 8)) << 8);
                    t[4 * q + 2] = (int) (((w1 >> 16) | (w2 << 16)) << 8); t[4 * q + 3] = (int) (w2 & 0xffffff00u);
                }
            }
            __syncwarp();
        } else if (g < groups) {
            const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src + g * 32);
            const uint4 a = __ldg(s4), b = __ldg(s4 + 1);
            const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            #pragma unroll
            for (int i = 0; i < 8; ++i) { t[2 * i] = (int) (w[i] << 16); t[2 * i + 1] = (int) (w[i] & 0xffff0000u); }
        }
        if (g < groups) {
            float v[16];
            #pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __fmul_rn((float) t[i], scale);
            if (SRC == 2) {
                float4* __restrict__ l4 = reinterpret_cast<float4*>(d0 + g * FR);
                float4* __restrict__ r4 = reinterpret_cast<float4*>(d0 + D.chStride + g * FR);
                l4[0] = make_float4(v[0], v[2], v[4], v[6]);  l4[1] = make_float4(v[8], v[10], v[12], v[14]);
                r4[0] = make_float4(v[1], v[3], v[5], v[7]);  r4[1] = make_float4(v[9], v[11], v[13], v[15]);
            } else {
                #pragma unroll
                for (int c = 0; c < DST; ++c) {
                    float4* __restrict__ x4 = reinterpret_cast<float4*>(d0 + (long long) c * D.chStride + g * FR);
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) x4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    if (blockIdx.x == 0) {                                         // the last frames (fewer than one group)
        const long long f0 = groups * FR;
        const int remFrames = (int) (D.numFrames - f0);
        for (int i = threadIdx.x; i < remFrames * DST; i += kThreads) {
            const int f = i / DST, c = i - f * DST;
            d0[(long long) c * D.chStride + f0 + f] = pcm_load(src, FMT24 ? F9_PCM_S24LE : F9_PCM_S16LE, (f0 + f) * SRC + min(c, SRC - 1));
        }
    }
}

// ---- planar <-> interleaved float (AudioProcessingService.swift:361-365, :524-531) ----------------------
__global__ void __launch_bounds__(kThreads)
interleave_kernel(const float* __restrict__ src, long long srcStride, int numCh, long long frames, float* __restrict__ dst, int kCvtFrames) {
    extern __shared__ float tile[];          // [numCh][kCvtFrames + 1]
    const long long f0 = (long long) blockIdx.x * kCvtFrames;
    const int nf = (int) min((long long) kCvtFrames, frames - f0);
    for (int c = 0; c < numCh; ++c)
        for (int f = threadIdx.x; f < nf; f += kThreads) tile[c * (kCvtFrames + 1) + f] = __ldg(src + (long long) c * srcStride + f0 + f);
    __syncthreads();
    const int total = nf * numCh;
    float* __restrict__ d = dst + f0 * numCh;
    for (int i = threadIdx.x; i < total; i += kThreads) { const int f = i / numCh, c = i - f * numCh; d[i] = tile[c * (kCvtFrames + 1) + f]; }
}
__global__ void __launch_bounds__(kThreads)
deinterleave_kernel(const float* __restrict__ src, int numCh, long long frames, float* __restrict__ dst, long long dstStride, int kCvtFrames) {
    extern __shared__ float tile[];
    const long long f0 = (long long) blockIdx.x * kCvtFrames;
    const int nf = (int) min((long long) kCvtFrames, frames - f0);
    const int total = nf * numCh;
    const float* __restrict__ s = src + f0 * numCh;
    for (int i = threadIdx.x; i < total; i += kThreads) { const int f = i / numCh, c = i - f * numCh; tile[c * (kCvtFrames + 1) + f] = __ldg(s + i); }
    __syncthreads();
    for (int c = 0; c < numCh; ++c)
        for (int f = threadIdx.x; f < nf; f += kThreads) dst[(long long) c * dstStride + f0 + f] = tile[c * (kCvtFrames + 1) + f];
}

template <typename K>
cudaError_t allow_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

}  // namespace

cudaError_t launch_trim(const DevBuf* d_captured, const int* d_latency, const DevBuf* d_out, int n, int maxOutFrames,
                        int maxCh, cudaStream_t s, long long* launches, double* d_dc_partials, const int* d_dc_mask, int dcDefault) {
    if (n <= 0 || maxCh <= 0) return cudaSuccess;
    const int tiles = std::max(1, std::min((maxOutFrames + kThreads * 16 - 1) / (kThreads * 16), 4096));
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = std::min(65535, n - b0);
        dim3 grid(tiles, maxCh, nb);
        if (d_dc_partials) {                                   // fused removeDCOffset: mean of the copied region first
            double* part = d_dc_partials + (size_t) b0 * maxCh * kDcPartials;
            dc_sum_src_kernel<<<dim3(kDcPartials, maxCh, nb), kThreads, 0, s>>>(d_captured + b0, d_latency + b0, d_out + b0, maxCh, part,
                                                                                d_dc_mask ? d_dc_mask + b0 : nullptr, dcDefault);
            ++*launches;
            trim_kernel<true><<<grid, kThreads, 0, s>>>(d_captured + b0, d_latency + b0, d_out + b0, part, d_dc_mask ? d_dc_mask + b0 : nullptr, maxCh);
        } else {
            trim_kernel<false><<<grid, kThreads, 0, s>>>(d_captured + b0, d_latency + b0, d_out + b0, nullptr, nullptr, maxCh);
        }
        ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_remove_dc(const DevBuf* d_bufs, int n, int maxCh, int maxFrames, double* d_partials, cudaStream_t s, long long* launches, int mode) {
    if (n <= 0 || maxCh <= 0) return cudaSuccess;
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = std::min(65535, n - b0);
        double* part = d_partials + (size_t) b0 * maxCh * kDcPartials;
        dc_sum_kernel<<<dim3(kDcPartials, maxCh, nb), kThreads, 0, s>>>(d_bufs + b0, maxCh, part, mode);
        ++*launches;
        const int tiles = std::max(1, std::min((maxFrames + kThreads * 16 - 1) / (kThreads * 16), 4096));
        dc_sub_kernel<<<dim3(tiles, maxCh, nb), kThreads, 0, s>>>(d_bufs + b0, maxCh, part);
        ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_pcm_to_planar(const void* d_src, int fmt, int srcCh, long long frames, float* d_dst,
                                 long long dstStride, int dstCh, cudaStream_t s, long long* launches) {
    if (frames <= 0) return cudaSuccess;
    const int bps = (fmt == F9_PCM_U8) ? 1 : (fmt == F9_PCM_S16LE) ? 2 : (fmt == F9_PCM_S24LE) ? 3 : 4;
    const int kCvtFrames = cvt_frames(srcCh * bps);
    const size_t smem = (size_t) kCvtFrames * srcCh * bps + 16;
    cudaError_t e = allow_smem(pcm_to_planar_kernel, smem);
    if (e != cudaSuccess) return e;
    const long long ctas = (frames + kCvtFrames - 1) / kCvtFrames;
    pcm_to_planar_kernel<<<(unsigned) ctas, kThreads, smem, s>>>((const unsigned char*) d_src, fmt, srcCh, frames, d_dst, dstStride, dstCh, kCvtFrames);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_planar_to_pcm24(const float* d_src, long long srcStride, int numCh, long long frames,
                                   unsigned char* d_dst, cudaStream_t s, long long* launches) {
    if (frames <= 0) return cudaSuccess;
    const int kCvtFrames = cvt_frames(numCh * 3);
    const size_t smem = (size_t) kCvtFrames * numCh * 3 + 16;
    cudaError_t e = allow_smem(planar_to_pcm24_kernel, smem);
    if (e != cudaSuccess) return e;
    const long long ctas = (frames + kCvtFrames - 1) / kCvtFrames;
    planar_to_pcm24_kernel<<<(unsigned) ctas, kThreads, smem, s>>>(d_src, srcStride, numCh, frames, d_dst, kCvtFrames);
    ++*launches;
    return cudaGetLastError();
}

// h_bufs: host copy of the descriptors (for sizing the grid); d_bufs / d_ptrs: the same on the device
cudaError_t launch_planar_to_pcm24_batch(const DevBuf* h_srcs, const DevBuf* d_srcs, unsigned char* const* d_dsts, int n,
                                         cudaStream_t s, long long* launches, unsigned char* const* h_dsts) {
    int maxCh = 0, maxFrames = 0;
    for (int i = 0; i < n; ++i) { maxCh = std::max(maxCh, h_srcs[i].numCh); maxFrames = std::max(maxFrames, h_srcs[i].numFrames); }
    if (n <= 0 || maxCh <= 0 || maxFrames <= 0) return cudaSuccess;
    // fast path: every file mono or every file stereo, planes and payloads on 16 bytes (h_dsts: host copy of the payload pointers)
    bool fast = h_dsts != nullptr && maxCh <= 2 && n <= 65535;
    for (int i = 0; i < n && fast; ++i)
        fast = h_srcs[i].numCh == maxCh && (reinterpret_cast<uintptr_t>(h_srcs[i].base) & 15) == 0 && (h_srcs[i].chStride & 3) == 0 &&
               (reinterpret_cast<uintptr_t>(h_dsts[i]) & 15) == 0;
    if (fast) {
        const long long groups = (long long) maxFrames * maxCh / 16;
        const unsigned ctas = (unsigned) std::max<long long>(1, std::min<long long>((groups + kThreads - 1) / kThreads, std::max(1, 148 * 8 / std::min(n, 148 * 8))));
        if (maxCh == 2) pcm24_pack_fast_kernel<2><<<dim3(ctas, (unsigned) n), kThreads, 0, s>>>(d_srcs, d_dsts);
        else pcm24_pack_fast_kernel<1><<<dim3(ctas, (unsigned) n), kThreads, 0, s>>>(d_srcs, d_dsts);
        ++*launches;
        return cudaGetLastError();
    }
    const int kCvtFrames = cvt_frames(maxCh * 3);
    const size_t smem = (size_t) kCvtFrames * maxCh * 3 + 16;
    cudaError_t e = allow_smem(planar_to_pcm24_batch_kernel, smem);
    if (e != cudaSuccess) return e;
    const unsigned ctas = (unsigned) ((maxFrames + kCvtFrames - 1) / kCvtFrames);
    for (int b0 = 0; b0 < n; b0 += 65535) {
        planar_to_pcm24_batch_kernel<<<dim3(ctas, (unsigned) std::min(65535, n - b0)), kThreads, smem, s>>>(d_srcs + b0, d_dsts + b0, kCvtFrames);
        ++*launches;
    }
    return cudaGetLastError();
}
cudaError_t launch_pcm_to_planar_batch(const unsigned char* const* d_srcs, int fmt, int srcCh, const DevBuf* h_dsts, const DevBuf* d_dsts, int n,
                                       cudaStream_t s, long long* launches, const unsigned char* const* h_srcs) {
    int maxFrames = 0;
    for (int i = 0; i < n; ++i) maxFrames = std::max(maxFrames, h_dsts[i].numFrames);
    if (n <= 0 || maxFrames <= 0) return cudaSuccess;
    // fast path: 24- or 16-bit payload of 1 or 2 channels into 1 or 2 planes (mono duplicated), everything on 16 bytes
    const int dstCh = h_dsts[0].numCh;
    bool fast = h_srcs != nullptr && (fmt == F9_PCM_S24LE || fmt == F9_PCM_S16LE) && srcCh <= 2 && dstCh >= srcCh && dstCh <= 2 && n <= 65535;
    for (int i = 0; i < n && fast; ++i)
        fast = h_dsts[i].numCh == dstCh && (reinterpret_cast<uintptr_t>(h_dsts[i].base) & 15) == 0 && (h_dsts[i].chStride & 3) == 0 &&
               (reinterpret_cast<uintptr_t>(h_srcs[i]) & 15) == 0;
    if (fast) {
        const long long groups = (long long) maxFrames * srcCh / 16;
        const unsigned ctas = (unsigned) std::max<long long>(1, std::min<long long>((groups + kThreads - 1) / kThreads, std::max(1, 148 * 8 / std::min(n, 148 * 8))));
        const dim3 grid(ctas, (unsigned) n);
        const bool f24 = fmt == F9_PCM_S24LE;
        if (srcCh == 2)      { if (f24) pcm_unpack_fast_kernel<2, 2, true><<<grid, kThreads, 0, s>>>(d_srcs, d_dsts); else pcm_unpack_fast_kernel<2, 2, false><<<grid, kThreads, 0, s>>>(d_srcs, d_dsts); }
        else if (dstCh == 2) { if (f24) pcm_unpack_fast_kernel<1, 2, true><<<grid, kThreads, 0, s>>>(d_srcs, d_dsts); else pcm_unpack_fast_kernel<1, 2, false><<<grid, kThreads, 0, s>>>(d_srcs, d_dsts); }
        else                 { if (f24) pcm_unpack_fast_kernel<1, 1, true><<<grid, kThreads, 0, s>>>(d_srcs, d_dsts); else pcm_unpack_fast_kernel<1, 1, false><<<grid, kThreads, 0, s>>>(d_srcs, d_dsts); }
        ++*launches;
        return cudaGetLastError();
    }
    const int bps = (fmt == F9_PCM_U8) ? 1 : (fmt == F9_PCM_S16LE) ? 2 : (fmt == F9_PCM_S24LE) ? 3 : 4;
    const int kCvtFrames = cvt_frames(srcCh * bps);
    const size_t smem = (size_t) kCvtFrames * srcCh * bps + 16;
    cudaError_t e = allow_smem(pcm_to_planar_batch_kernel, smem);
    if (e != cudaSuccess) return e;
    const unsigned ctas = (unsigned) ((maxFrames + kCvtFrames - 1) / kCvtFrames);
    for (int b0 = 0; b0 < n; b0 += 65535) {
        pcm_to_planar_batch_kernel<<<dim3(ctas, (unsigned) std::min(65535, n - b0)), kThreads, smem, s>>>(d_srcs + b0, fmt, srcCh, d_dsts + b0, kCvtFrames);
        ++*launches;
    }
    return cudaGetLastError();
}

cudaError_t launch_interleave(const float* d_src, long long srcStride, int numCh, long long frames, float* d_dst,
                              cudaStream_t s, long long* launches) {
    if (frames <= 0) return cudaSuccess;
    const int kCvtFrames = cvt_frames(numCh * 4);
    const size_t smem = sizeof(float) * (size_t) numCh * (kCvtFrames + 1);
    cudaError_t e = allow_smem(interleave_kernel, smem);
    if (e != cudaSuccess) return e;
    const long long ctas = (frames + kCvtFrames - 1) / kCvtFrames;
    interleave_kernel<<<(unsigned) ctas, kThreads, smem, s>>>(d_src, srcStride, numCh, frames, d_dst, kCvtFrames);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_deinterleave(const float* d_src, int numCh, long long frames, float* d_dst, long long dstStride,
                                cudaStream_t s, long long* launches) {
    if (frames <= 0) return cudaSuccess;
    const int kCvtFrames = cvt_frames(numCh * 4);
    const size_t smem = sizeof(float) * (size_t) numCh * (kCvtFrames + 1);
    cudaError_t e = allow_smem(deinterleave_kernel, smem);
    if (e != cudaSuccess) return e;
    const long long ctas = (frames + kCvtFrames - 1) / kCvtFrames;
    deinterleave_kernel<<<(unsigned) ctas, kThreads, smem, s>>>(d_src, numCh, frames, d_dst, dstStride, kCvtFrames);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace f9
