// Sample-rate conversion kernels: juce::Interpolators semantics (GenericInterpolator position recurrence in
// closed form, Lagrange / WindowedSinc / CatmullRom / Linear / ZeroOrderHold traits) evaluated in parallel.
//
// Position.  JUCE keeps a double `pos` (subSamplePos, 1.0 after reset()); per output it pushes inputs while
// pos >= 1, evaluates the traits at (float) pos and adds the ratio.  In closed form output n sees
//     T_n = pos0 + n * ratio,   c_n = floor(T_n)  inputs consumed so far,   offset_n = T_n - c_n,
// and reads the `taps` inputs ending at index c_n - 1.  Rational ratios p/q use exact integers
// (n = a*q + k  ->  newest input a*p + B[k], B[k] = floor(k*p/q), phase (k*p) mod q) and per-phase weights
// tabulated on the host; other ratios use a double-double product.  See DESIGN.md "Position arithmetic" for
// why this stays inside the 2^-20 sample tolerance of the sequential recurrence.
#include "f9_internal.cuh"

namespace f9 {
namespace {

constexpr int kRsThreads = 256;

__device__ __forceinline__ int find_seg(const int* __restrict__ prefix, int n, int bid) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= bid) lo = mid; else hi = mid; }
    return lo;
}

struct Pos { long long c; float offset; };

// T = pos0 + n*ratio with a double-double product; c = floor(T), offset = (float)(T - c) in [0, 1].
__device__ __forceinline__ Pos pos_generic(double pos0, double ratio, long long n) {
    const double dn = (double) n;
    const double hi = dn * ratio;
    const double lo = fma(dn, ratio, -hi);
    const double s = pos0 + hi;
    const double bb = s - pos0;
    double err = (pos0 - (s - bb)) + (hi - bb);
    err += lo;
    double fl = floor(s);
    double frac = (s - fl) + err;
    if (frac < 0.0) { fl -= 1.0; frac += 1.0; }
    else if (frac >= 1.0) { fl += 1.0; frac -= 1.0; }
    Pos p; p.c = (long long) fl; p.offset = (float) frac;
    return p;
}

__device__ __forceinline__ float load_in(const Seg& S, long long g) {
    const long long l = g - S.inOffset;
    (void) g;   // indices before the channel start have l < 0 whenever in_offset >= 0; a negative in_offset maps history
    return (l >= 0 && l < S.inAvail) ? __ldg(S.in + l) : 0.0f;
}

// --------------------------------------------------------------------------------------------- polyphase, v1
// One CTA per tile of `tileOut` consecutive outputs of one segment.  The input window (tile span + taps - 1
// samples of halo) is staged in shared memory with coalesced loads; weights W[tap][slot] are read through L1
// with consecutive lanes on consecutive slots.
template <int TAPS>
__global__ void __launch_bounds__(kRsThreads)
poly_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, PolyDev P, int tileOut,
            int adding, float gain) {
    extern __shared__ float xs[];
    const int sidx = find_seg(tilePrefix, nSegs, blockIdx.x);
    const Seg S = segs[sidx];
    const int tile = blockIdx.x - tilePrefix[sidx];
    const long long o0 = (long long) tile * tileOut;
    const int cnt = (int) min((long long) tileOut, S.numOut - o0);
    const long long nFirst = S.n0 + o0, nLast = nFirst + cnt - 1;
    const long long a0 = nFirst / P.q;  const int k0 = (int) (nFirst - a0 * P.q);
    const long long a1 = nLast / P.q;   const int k1 = (int) (nLast - a1 * P.q);
    const int Bk0 = __ldg(P.B + k0);
    const long long mFirst = a0 * P.p + Bk0;
    const long long mLast = a1 * P.p + __ldg(P.B + k1);
    const long long lo = mFirst - (TAPS - 1);
    const int span = (int) (mLast - lo + 1);

    for (int i = threadIdx.x; i < span; i += kRsThreads) xs[i] = load_in(S, lo + i);
    __syncthreads();

    float* __restrict__ out = S.out + o0;
    for (int o = threadIdx.x; o < cnt; o += kRsThreads) {
        const int kk = k0 + o;
        const int arel = kk / P.q;
        const int k = kk - arel * P.q;
        const int rel = arel * P.p + __ldg(P.B + k) - Bk0;       // oldest tap of this output inside xs
        const float* __restrict__ w = P.W + k;
        float acc = 0.0f;
        #pragma unroll 8
        for (int j = 0; j < TAPS; ++j) acc = fmaf(xs[rel + j], __ldg(w + (size_t) j * P.qpad), acc);
        out[o] = adding ? __fadd_rn(out[o], __fmul_rn(gain, acc)) : acc;
    }
}

// --------------------------------------------------------------------------------------------- generic ratio
// Any double ratio (and any pos0: the stateful process() calls land here).  Weights are evaluated per output in
// the scalar interpolator's own operation order with explicit _rn intrinsics (no FMA contraction).
template <int KIND> struct KindTaps;
template <> struct KindTaps<F9_WINDOWED_SINC> { static constexpr int v = 200; };
template <> struct KindTaps<F9_LAGRANGE> { static constexpr int v = 5; };
template <> struct KindTaps<F9_CATMULL_ROM> { static constexpr int v = 4; };
template <> struct KindTaps<F9_LINEAR> { static constexpr int v = 2; };
template <> struct KindTaps<F9_ZERO_ORDER_HOLD> { static constexpr int v = 1; };

__device__ __forceinline__ float eval_sinc(const float* __restrict__ x, float offset, const float* __restrict__ table) {
    // WindowedSincTraits::valueAtOffset; x[0] is the oldest of the 200 ring samples.
    float result = 0.0f, firstFrac = 0.0f, lastSincPosition = -1.0f;
    int index = 0, sign = -1;
    const float base = __fsub_rn(1.0f, offset);
    for (int i = -100; i < 100; ++i) {
        const float sincPosition = __fadd_rn(base, (float) i);
        if (i == -100 || (sincPosition >= 0.0f && lastSincPosition < 0.0f)) {
            const float indexFloat = __fmul_rn(fabsf(sincPosition), 100.0f);
            const float indexFloored = floorf(indexFloat);
            index = (int) indexFloored;
            firstFrac = __fsub_rn(indexFloat, indexFloored);
            sign = (sincPosition < 0.0f) ? -1 : 1;
        }
        if (sincPosition == 0.0f) result = __fadd_rn(result, x[i + 100]);
        else if (sincPosition < 100.0f && sincPosition > -100.0f) {
            const float v1 = __ldg(table + index), v2 = __ldg(table + index + 1);
            const float w = __fadd_rn(v1, __fmul_rn(firstFrac, __fsub_rn(v2, v1)));
            result = __fadd_rn(result, __fmul_rn(x[i + 100], w));
        }
        lastSincPosition = sincPosition;
        index += 100 * sign;
    }
    return result;
}

template <int K> __device__ __forceinline__ float lag_coef(float input, float offset) {
    #pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (j == K) continue;
        input = __fmul_rn(input, __fmul_rn(__fsub_rn((float) (j - 2), offset), 1.0f / (float) (j - K)));
    }
    return input;
}
__device__ __forceinline__ float eval_lagrange(const float* __restrict__ x, float offset) {
    float r = 0.0f;
    r = __fadd_rn(r, lag_coef<0>(x[0], offset));
    r = __fadd_rn(r, lag_coef<1>(x[1], offset));
    r = __fadd_rn(r, lag_coef<2>(x[2], offset));
    r = __fadd_rn(r, lag_coef<3>(x[3], offset));
    r = __fadd_rn(r, lag_coef<4>(x[4], offset));
    return r;
}
__device__ __forceinline__ float eval_catmull(const float* __restrict__ x, float offset) {
    const float y0 = x[0], y1 = x[1], y2 = x[2], y3 = x[3];
    const float halfY0 = __fmul_rn(0.5f, y0), halfY3 = __fmul_rn(0.5f, y3);
    const float t3 = __fsub_rn(__fadd_rn(halfY3, __fmul_rn(1.5f, y1)), __fadd_rn(halfY0, __fmul_rn(1.5f, y2)));
    const float t2 = __fadd_rn(__fsub_rn(__fadd_rn(y0, __fmul_rn(2.0f, y2)), __fadd_rn(halfY3, __fmul_rn(2.5f, y1))), __fmul_rn(offset, t3));
    const float t1 = __fadd_rn(__fsub_rn(__fmul_rn(0.5f, y2), halfY0), __fmul_rn(offset, t2));
    return __fadd_rn(y1, __fmul_rn(offset, t1));
}

template <int KIND>
__global__ void __launch_bounds__(kRsThreads)
generic_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, double ratio, double pos0,
               const float* __restrict__ sincTable, int tileOut, int adding, float gain) {
    constexpr int TAPS = KindTaps<KIND>::v;
    extern __shared__ float xs[];
    const int sidx = find_seg(tilePrefix, nSegs, blockIdx.x);
    const Seg S = segs[sidx];
    const int tile = blockIdx.x - tilePrefix[sidx];
    const long long o0 = (long long) tile * tileOut;
    const int cnt = (int) min((long long) tileOut, S.numOut - o0);
    const long long nFirst = S.n0 + o0;
    const Pos pf = pos_generic(pos0, ratio, nFirst);
    const Pos pl = pos_generic(pos0, ratio, nFirst + cnt - 1);
    const long long lo = (pf.c - 1) - (TAPS - 1) - 1;          // one sample of slack on both sides
    const int span = (int) ((pl.c - 1) + 1 - lo + 1);
    for (int i = threadIdx.x; i < span; i += kRsThreads) xs[i] = load_in(S, lo + i);
    __syncthreads();

    float* __restrict__ out = S.out + o0;
    for (int o = threadIdx.x; o < cnt; o += kRsThreads) {
        const Pos p = pos_generic(pos0, ratio, nFirst + o);
        int rel = (int) ((p.c - 1) - (TAPS - 1) - lo);
        rel = max(0, min(rel, span - TAPS));
        const float* x = xs + rel;
        float v;
        if (KIND == F9_WINDOWED_SINC) v = eval_sinc(x, p.offset, sincTable);
        else if (KIND == F9_LAGRANGE) v = eval_lagrange(x, p.offset);
        else if (KIND == F9_CATMULL_ROM) v = eval_catmull(x, p.offset);
        else if (KIND == F9_LINEAR) v = __fadd_rn(__fmul_rn(x[1], p.offset), __fmul_rn(x[0], __fsub_rn(1.0f, p.offset)));
        else v = x[0];
        out[o] = adding ? __fadd_rn(out[o], __fmul_rn(gain, v)) : v;
    }
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

template <int TAPS>
cudaError_t run_poly(const ResampleLaunch& L, size_t smem, cudaStream_t s) {
    cudaError_t e = set_smem(poly_kernel<TAPS>, smem);
    if (e != cudaSuccess) return e;
    poly_kernel<TAPS><<<L.n_tiles, kRsThreads, smem, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.poly, L.tile_out, L.adding, L.gain);
    return cudaGetLastError();
}
template <int KIND>
cudaError_t run_generic(const ResampleLaunch& L, size_t smem, cudaStream_t s) {
    cudaError_t e = set_smem(generic_kernel<KIND>, smem);
    if (e != cudaSuccess) return e;
    generic_kernel<KIND><<<L.n_tiles, kRsThreads, smem, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.ratio, L.pos0,
                                                             L.d_sinc_table, L.tile_out, L.adding, L.gain);
    return cudaGetLastError();
}

}  // namespace

int choose_tile_out(double ratio) {
    double t = 8000.0 / (ratio > 0.25 ? ratio : 0.25);
    int tile = (int) t / 256 * 256;
    if (tile < 256) tile = 256;
    if (tile > 4096) tile = 4096;
    return tile;
}

cudaError_t launch_resample(const ResampleLaunch& L, cudaStream_t s, long long* launches) {
    if (L.n_tiles <= 0) return cudaSuccess;
    const int taps = interp_memory(L.kind);
    const size_t smem = sizeof(float) * ((size_t) ((double) L.tile_out * L.ratio) + (size_t) taps + 8);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e;
    if (L.rational) {
        switch (L.kind) {
            case F9_WINDOWED_SINC: e = run_poly<200>(L, smem, s); break;
            case F9_LAGRANGE: e = run_poly<5>(L, smem, s); break;
            case F9_CATMULL_ROM: e = run_poly<4>(L, smem, s); break;
            case F9_LINEAR: e = run_poly<2>(L, smem, s); break;
            case F9_ZERO_ORDER_HOLD: e = run_poly<1>(L, smem, s); break;
            default: return cudaErrorInvalidValue;
        }
    } else {
        switch (L.kind) {
            case F9_WINDOWED_SINC: e = run_generic<F9_WINDOWED_SINC>(L, smem, s); break;
            case F9_LAGRANGE: e = run_generic<F9_LAGRANGE>(L, smem, s); break;
            case F9_CATMULL_ROM: e = run_generic<F9_CATMULL_ROM>(L, smem, s); break;
            case F9_LINEAR: e = run_generic<F9_LINEAR>(L, smem, s); break;
            case F9_ZERO_ORDER_HOLD: e = run_generic<F9_ZERO_ORDER_HOLD>(L, smem, s); break;
            default: return cudaErrorInvalidValue;
        }
    }
    if (e == cudaSuccess) ++*launches;
    return e;
}

}  // namespace f9
