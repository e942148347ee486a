// Bounded-lag cross-correlation with argmax: candidates fast, verification exact.
//
//     r_c[lag] = sum_i (double) x[i] * (double) y_c[i + lag],   argmax over (channel, lag) of |r| in findPeakPosition's order
//
// (north_star (b); the reference peak-picks an impulse, Source/MainComponent.cpp:950-975 -- its scan order and strict '>' are what
// xc_better keeps).  The exact double sums of every lag (f9_scan.cu: xcorr_partial_kernel) are FP64-bound: 52 ms for config 4's
// 512 stereo recordings against a 4800-sample sweep over +-2^16 lags.  The argmax does not need them all:
//   1. xc_approx_kernel computes every lag approximately on the tensor cores (fp16 operands scaled by powers of two, fp32
//      accumulation; mma.sync m16n8k16 -- the operand of the lag axis is the recording itself read as a Hankel matrix,
//      B[k, n] = y[16 n + k], the stimulus is a 16-row Toeplitz band A[m, k] = x[k - m], so C[m, n] = r[16 n + m]) together with a
//      rigorous bound eps on |r~ - r| for the tile: operand rounding 2 * 2^-11 and accumulation 2^-14 times sum |x_i| |y_i+lag|,
//      itself bounded by ||x||_2 * ||y over the tile's span||_2 (Cauchy-Schwarz), plus the subnormal floors of both operands;
//   2. xc_select_kernel takes LB = max (|r~| - eps), a lower bound of the true maximum, and lists the lags with |r~| + eps >= LB:
//      no other lag can win;
//   3. xc_exact_kernel evaluates the listed lags with the reference chain (fma in i order: bit-exact doubles) and xc_pick_kernel
//      keeps the best in findPeakPosition's order.  Value, channel and lag are therefore those of the exact scan, ties included.
// A buffer whose list overflows (silence, periodic signals: many near-maxima) or whose approximation is not finite takes the
// exact kernel for all its lags (xcorr_partial_kernel with a per-buffer mask).
#include <cuda_fp16.h>

#include "f9_internal.cuh"

namespace f9 {
namespace {

constexpr int kXaThreads = 256;                       // 8 warps
constexpr int kXaNT = 8;                              // n-tiles (8 columns of 16 lags) per warp: 1024 lags per warp
constexpr int kXaLags = (kXaThreads / 32) * kXaNT * 128;   // 8192 lags per CTA
constexpr int kXaKC = 1024;                           // stimulus samples per shared-memory chunk
constexpr int kXsStride = kXaKC + 24 + 48;            // halves per shifted stimulus copy: stride * 2 B = 16 mod 128 (conflict-free ldmatrix rows)
static_assert((kXsStride * 2) % 128 == 16, "stimulus copies must sit 16 bytes apart modulo a bank row");
constexpr int kYsLen = kXaLags + kXaKC + 16;
constexpr float kXScale = 16.0f, kYScale = 1024.0f;   // fp16 operands stay normal down to |x| = 3.8e-6, |y| = 6e-8; overflow (inf) -> exact fallback
constexpr int kXcCandCap = 512;                       // candidate lags per buffer before the exact fallback

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid: (lag tiles, channel, buffer).  rt: approximate r per lag [buffer][maxCh][tiles * kXaLags]; tileMax / tileEn: per CTA the
// largest |r~| (NaN-propagating) and the energy of the recording span the tile reads.
__global__ void __launch_bounds__(kXaThreads)
xc_approx_kernel(const DevBuf* __restrict__ bufs, const float* __restrict__ stim, int stimLen, int lagMin, int lagMax, int maxCh, int tiles,
                 float* __restrict__ rt, float* __restrict__ tileMax, float* __restrict__ tileEn) {
    __shared__ __align__(128) __half xs[8 * kXsStride];
    __shared__ __align__(128) __half ys[kYsLen];
    __shared__ float red[kXaThreads / 32];
    const DevBuf B = bufs[blockIdx.z];
    const int ch = blockIdx.y, tile = blockIdx.x;
    const size_t slot = ((size_t) blockIdx.z * maxCh + ch) * tiles + tile;
    if (ch >= B.numCh) { if (threadIdx.x == 0) { tileMax[slot] = 0.0f; tileEn[slot] = 0.0f; } return; }
    const float* __restrict__ y = B.base + (long long) ch * B.chStride;
    const long long lag0 = (long long) lagMin + (long long) tile * kXaLags;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[kXaNT][4];
    #pragma unroll
    for (int n = 0; n < kXaNT; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f; }
    float energy = 0.0f;
    const int nChunks = (stimLen + kXaKC - 1) / kXaKC;
    for (int c = 0; c < nChunks; ++c) {
        const int i0 = c * kXaKC;
        // stimulus chunk, eight copies shifted by 0..7 samples: XS_s[j] = x[i0 + j - 8 - s] (zero outside the stimulus / the chunk)
        for (int t = threadIdx.x; t < 8 * (kXaKC + 24); t += kXaThreads) {
            const int s = t / (kXaKC + 24), j = t - s * (kXaKC + 24);
            const int ii = j - 8 - s;
            const float v = (ii >= 0 && ii < kXaKC && i0 + ii < stimLen) ? __ldg(stim + i0 + ii) * kXScale : 0.0f;
            xs[s * kXsStride + j] = __float2half_rn(v);
        }
        // recording span of this chunk: ys[j] = y[lag0 + i0 + j]; energy of each sample counted once over the chunks
        const int fresh = (c == nChunks - 1) ? kYsLen : kXaKC;
        for (int j = threadIdx.x; j < kYsLen; j += kXaThreads) {
            const long long g = lag0 + i0 + j;
            const float v = (g >= 0 && g < B.numFrames) ? __ldg(y + g) * kYScale : 0.0f;
            const __half h = __float2half_rn(v);
            ys[j] = h;
            if (j < fresh) { const float f = __half2float(h); energy = fmaf(f, f, energy); }
        }
        __syncthreads();
        const int ksteps = (min(kXaKC, stimLen - i0) + 15 + 15) / 16;          // k reaches i + m, m <= 15
        // ldmatrix row addresses of this lane: A matrix j = lane / 8 (rows 0-7 k0-7 | rows 8-15 k0-7 | rows 0-7 k8-15 | rows 8-15 k8-15)
        const int mj = lane >> 3, rr = lane & 7;
        const uint32_t aBase = smem_addr(xs + rr * kXsStride + ((mj == 0) ? 8 : (mj == 1) ? 0 : (mj == 2) ? 16 : 8));
        // B matrix j: columns n0 + 8 * (j / 2) + rr, k half j % 2
        const uint32_t bBase = smem_addr(ys + 16 * (warp * kXaNT * 8 + (mj >> 1) * 8 + rr) + (mj & 1) * 8);
        #pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
            uint32_t a[4];
            ldmatrix_x4(a, aBase + (uint32_t) (ks * 32));
            #pragma unroll
            for (int np = 0; np < kXaNT / 2; ++np) {
                uint32_t b[4];
                ldmatrix_x4(b, bBase + (uint32_t) (ks * 32 + np * 16 * 16 * 2));
                mma16816(acc[2 * np], a, b[0], b[1]);
                mma16816(acc[2 * np + 1], a, b[2], b[3]);
            }
        }
        __syncthreads();
    }
    // C[m, n] = r~[lag0 + 16 (n0 + n) + m] * (kXScale * kYScale); this lane holds rows g = lane / 4 (+ 8), columns 2 (lane % 4) (+ 1)
    const float unscale = 1.0f / (kXScale * kYScale);
    float* __restrict__ out = rt + slot * (size_t) kXaLags;
    float mx = 0.0f; bool bad = false;
    const int g = lane >> 2, q2 = 2 * (lane & 3);
    #pragma unroll
    for (int n = 0; n < kXaNT; ++n) {
        #pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = g + ((e & 2) ? 8 : 0), col = warp * kXaNT * 8 + n * 8 + q2 + (e & 1);
            const int off = 16 * col + m;
            const float v = acc[n][e] * unscale;
            out[off] = v;
            if (lag0 + off <= lagMax) { bad = bad || !(fabsf(v) <= 3.0e38f); mx = fmaxf(mx, fabsf(v)); }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); energy += __shfl_xor_sync(0xffffffffu, energy, o); }
    bad = __any_sync(0xffffffffu, bad);
    __shared__ float redE[kXaThreads / 32]; __shared__ int redB[kXaThreads / 32];
    if (lane == 0) { red[warp] = mx; redE[warp] = energy; redB[warp] = bad ? 1 : 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m2 = 0.0f, e2 = 0.0f; int b2 = 0;
        for (int w = 0; w < kXaThreads / 32; ++w) { m2 = fmaxf(m2, red[w]); e2 += redE[w]; b2 |= redB[w]; }
        tileMax[slot] = b2 ? __int_as_float(0x7fc00000) : m2;
        tileEn[slot] = e2 / (kYScale * kYScale);
    }
}

struct XcCand { int ch, lag; double v; };

// One CTA per buffer: eps per tile, LB, candidate list (or the fallback flag).
__global__ void __launch_bounds__(256)
xc_select_kernel(const DevBuf* __restrict__ bufs, const float* __restrict__ stim, int stimLen, int lagMin, int lagMax, int maxCh, int tiles,
                 const float* __restrict__ rt, const float* __restrict__ tileMax, const float* __restrict__ tileEn,
                 XcCand* __restrict__ cands, int* __restrict__ candCount, int* __restrict__ needExact) {
    const int b = blockIdx.x;
    const DevBuf B = bufs[b];
    __shared__ double sh[8]; __shared__ float shf[8]; __shared__ int shi[8];
    __shared__ float sLB, sXn, sX1; __shared__ int sBad, sCount;
    // ||x||_2 (and sqrt(S) ||x||_2 >= ||x||_1) in double
    double xx = 0.0;
    for (int i = threadIdx.x; i < stimLen; i += blockDim.x) { const double v = (double) stim[i]; xx += v * v; }
    for (int o = 16; o > 0; o >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = xx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; for (int w = 0; w < 8; ++w) t += sh[w];
        sXn = (float) (sqrt(t) * 1.0001); sX1 = (float) (sqrt(t * (double) stimLen) * 1.0001); sBad = 0; sCount = 0;
    }
    __syncthreads();
    const float xn = sXn, x1 = sX1;
    // eps(tile) = 1.1 * 2^-10 * ||x|| * ||y span|| + subnormal floors: 2^-35 per recording sample times ||x||_1, 2^-29 per tap times ||y span||_1
    auto eps_of = [&](float en) {
        const float yn = sqrtf(en) * 1.002f;
        return 1.1f * 0.0009765625f * xn * yn + x1 * 2.9103830456733704e-11f + yn * sqrtf((float) stimLen) * 1.862645149230957e-9f;
    };
    const int nSlots = B.numCh * tiles;
    float lb = 0.0f; int bad = 0;
    for (int s = threadIdx.x; s < nSlots; s += blockDim.x) {
        const int ch = s / tiles, t = s - ch * tiles;
        const size_t slot = ((size_t) b * maxCh + ch) * tiles + t;
        const float m = tileMax[slot];
        if (!(m <= 3.0e38f)) bad = 1; else lb = fmaxf(lb, m - eps_of(tileEn[slot]));
    }
    for (int o = 16; o > 0; o >>= 1) { lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, o)); bad |= __shfl_xor_sync(0xffffffffu, bad, o); }
    if ((threadIdx.x & 31) == 0) { shf[threadIdx.x >> 5] = lb; shi[threadIdx.x >> 5] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) { float l = 0.0f; int bb = 0; for (int w = 0; w < 8; ++w) { l = fmaxf(l, shf[w]); bb |= shi[w]; } sLB = l; sBad = bb; }
    __syncthreads();
    const float LB = sLB;
    if (sBad) { if (threadIdx.x == 0) { needExact[b] = 1; candCount[b] = 0; } return; }
    // candidates: |r~| + eps >= LB.  Only tiles whose own maximum can reach LB are read.
    for (int s = 0; s < nSlots; ++s) {
        const int ch = s / tiles, t = s - ch * tiles;
        const size_t slot = ((size_t) b * maxCh + ch) * tiles + t;
        const float e = eps_of(tileEn[slot]);
        if (tileMax[slot] + e < LB) continue;
        const float* __restrict__ r = rt + slot * (size_t) kXaLags;
        const long long lag0 = (long long) lagMin + (long long) t * kXaLags;
        for (int o = threadIdx.x; o < kXaLags; o += blockDim.x) {
            if (lag0 + o > lagMax) break;
            if (fabsf(r[o]) + e >= LB) {
                const int at = atomicAdd(&sCount, 1);
                if (at < kXcCandCap) { XcCand c; c.ch = ch; c.lag = (int) (lag0 + o); c.v = 0.0; cands[(size_t) b * kXcCandCap + at] = c; }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = sCount;
        needExact[b] = (n > kXcCandCap || n == 0) ? 1 : 0;       // n == 0 cannot happen (the tile that set LB passes); kept as a guard
        candCount[b] = n > kXcCandCap ? 0 : n;
    }
}

// One thread per candidate: the reference chain, i ascending, one rounding per term (the float product is exact in double).
__global__ void __launch_bounds__(128)
xc_exact_kernel(const DevBuf* __restrict__ bufs, const float* __restrict__ stim, int stimLen, XcCand* __restrict__ cands, const int* __restrict__ candCount) {
    const int b = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= candCount[b]) return;
    const DevBuf B = bufs[b];
    XcCand c = cands[(size_t) b * kXcCandCap + k];
    const float* __restrict__ y = B.base + (long long) c.ch * B.chStride;
    const int i0 = max(0, -c.lag), i1 = min(stimLen, B.numFrames - c.lag);
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) acc = fma((double) __ldg(stim + i), (double) __ldg(y + i + c.lag), acc);
    cands[(size_t) b * kXcCandCap + k].v = fabs(acc);
}

// Best candidate per buffer in findPeakPosition's order; buffers that took the exact fallback keep its result.
__global__ void xc_pick_kernel(const XcCand* __restrict__ cands, const int* __restrict__ candCount, const int* __restrict__ needExact,
                               const XcPartial* __restrict__ fallback, int n, XcPartial* __restrict__ best) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    if (needExact[b]) { best[b] = fallback[b]; return; }
    double bv = 0.0; int bch = -1, blag = 0;
    for (int k = 0; k < candCount[b]; ++k) {
        const XcCand c = cands[(size_t) b * kXcCandCap + k];
        if (c.v > 0.0 && (bch < 0 || c.v > bv || (c.v == bv && (c.ch < bch || (c.ch == bch && c.lag < blag))))) { bv = c.v; bch = c.ch; blag = c.lag; }
    }
    XcPartial o; o.v = bv; o.ch = bch; o.lag = blag; o.pad = 0;
    best[b] = o;
}

}  // namespace

size_t xcorr_fast_scratch_bytes(int n, int maxCh, int lagMin, int lagMax) {
    const int tiles = (lagMax - lagMin + 1 + kXaLags - 1) / kXaLags;
    const size_t slots = (size_t) n * maxCh * tiles;
    return slots * kXaLags * sizeof(float) + 2 * slots * sizeof(float) + (size_t) n * (kXcCandCap * sizeof(XcCand) + 2 * sizeof(int) + sizeof(XcPartial)) + 4096;
}

// d_scratch: xcorr_fast_scratch_bytes; d_partials / d_prefix / total_ctas: the exact kernel's (fallback).  The candidates' values are
// bit-identical to the exact scan's, so d_best is identical to launch_xcorr's.
cudaError_t launch_xcorr_fast(const DevBuf* h_bufs, const DevBuf* d_bufs, int n, int total_ctas, const int* d_prefix, const float* d_stim, int stimLen,
                              int lagMin, int lagMax, XcPartial* d_partials, XcPartial* d_best, void* d_scratch, cudaStream_t s, long long* launches) {
    if (n <= 0) return cudaSuccess;
    int maxCh = 0;
    for (int i = 0; i < n; ++i) maxCh = std::max(maxCh, h_bufs[i].numCh);
    if (maxCh <= 0) return launch_xcorr(d_bufs, n, total_ctas, d_prefix, d_stim, stimLen, lagMin, lagMax, d_partials, d_best, s, launches);
    const int tiles = (lagMax - lagMin + 1 + kXaLags - 1) / kXaLags;
    const size_t slots = (size_t) n * maxCh * tiles;
    char* p = (char*) d_scratch;
    float* rt = (float*) p; p += slots * kXaLags * sizeof(float);
    float* tileMax = (float*) p; p += slots * sizeof(float);
    float* tileEn = (float*) p; p += slots * sizeof(float);
    p = (char*) (((uintptr_t) p + 15) & ~(uintptr_t) 15);
    XcCand* cands = (XcCand*) p; p += (size_t) n * kXcCandCap * sizeof(XcCand);
    XcPartial* fallback = (XcPartial*) p; p += (size_t) n * sizeof(XcPartial);
    int* candCount = (int*) p; p += (size_t) n * sizeof(int);
    int* needExact = (int*) p;
    xc_approx_kernel<<<dim3((unsigned) tiles, (unsigned) maxCh, (unsigned) n), kXaThreads, 0, s>>>(d_bufs, d_stim, stimLen, lagMin, lagMax, maxCh, tiles, rt, tileMax, tileEn);
    ++*launches;
    xc_select_kernel<<<n, 256, 0, s>>>(d_bufs, d_stim, stimLen, lagMin, lagMax, maxCh, tiles, rt, tileMax, tileEn, cands, candCount, needExact);
    ++*launches;
    xc_exact_kernel<<<dim3((kXcCandCap + 127) / 128, (unsigned) n), 128, 0, s>>>(d_bufs, d_stim, stimLen, cands, candCount);
    ++*launches;
    // exact scan of every lag for the buffers that need it (the CTAs of the others return at once)
    cudaError_t e = launch_xcorr(d_bufs, n, total_ctas, d_prefix, d_stim, stimLen, lagMin, lagMax, d_partials, fallback, s, launches, needExact);
    if (e != cudaSuccess) return e;
    xc_pick_kernel<<<(n + 127) / 128, 128, 0, s>>>(cands, candCount, needExact, fallback, n, d_best);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace f9
