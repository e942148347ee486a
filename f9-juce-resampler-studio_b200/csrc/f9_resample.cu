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
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <tuple>

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

// --------------------------------------------------------------------------------------------- short kinds (<= 5 taps)
// Lagrange / CatmullRom / Linear / ZeroOrderHold at a rational ratio: 2-10 FLOP per output against 5-13 bytes, i.e.
// HBM-bound on CUDA cores (north_star's design for the short kernels).  A tile is `tileOut` consecutive outputs of a segment;
// persistent CTAs walk the tiles blockIdx.x, + gridDim.x, ... through a ring of kShortStages shared-memory stages:
//   * the tile's input span (+ taps - 1 of halo), aligned down to 16 bytes, is brought in with cp.async (16 bytes per copy, no
//     register staging) two tiles ahead of the arithmetic; a tile that touches the edge of its segment's window is staged with
//     guarded loads instead (zeros outside the window);
//   * thread t owns the outputs o = t, t + S, t + 2S, ... of a tile, S = c*q a multiple of the period, so its slot -- and with
//     it the TAPS weights and the window offset -- is fixed per tile: an output costs TAPS shared loads, TAPS FFMA and one
//     store, the window advances by c*p samples per step; consecutive lanes write consecutive outputs (coalesced stores);
//   * the geometry of a tile comes from its record (short_tile_table_kernel, one thread per tile, just before this launch).
struct ShortTileRec {
    const float* in; float* out;   // the segment's window; the tile's first output
    long long lA, inAvail;         // 16-byte aligned start of the tile's span as an index into the window; window length
    int cnt, nvec, k0, mis;        // outputs; 16-byte vectors to stage; slot of output 0; the span proper starts at xs[mis]
};

template <int TAPS>
__global__ void __launch_bounds__(256)
short_tile_table_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles, PolyDev P, int tileOut,
                        ShortTileRec* __restrict__ recs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTiles) return;
    const int sidx = find_seg(tilePrefix, nSegs, t);
    const Seg Sg = segs[sidx];
    const int tile = t - tilePrefix[sidx];
    const long long o0 = (long long) tile * tileOut;
    const int cnt = (int) min((long long) tileOut, Sg.numOut - o0);
    const long long nFirst = Sg.n0 + o0, nLast = nFirst + cnt - 1;
    const long long a0 = nFirst / P.q;  const int k0 = (int) (nFirst - a0 * P.q);
    const long long a1 = nLast / P.q;   const int k1 = (int) (nLast - a1 * P.q);
    const long long lo = a0 * P.p + __ldg(P.B + k0) - (TAPS - 1);        // oldest sample of the tile (channel index)
    const int span = (int) (a1 * P.p + __ldg(P.B + k1) - lo + 1);
    const long long l0 = lo - Sg.inOffset;                               // the same as an index into the segment's window
    const int mis = (int) (((long long) (reinterpret_cast<uintptr_t>(Sg.in) >> 2) + l0) & 3);
    ShortTileRec R;
    R.in = Sg.in; R.out = Sg.out + o0; R.lA = l0 - mis; R.inAvail = Sg.inAvail;
    R.cnt = cnt; R.nvec = (span + mis + 3) >> 2; R.k0 = k0; R.mis = mis;
    recs[t] = R;
}

template <int TAPS, int kShortStages>
__global__ void __launch_bounds__(256)
short_kernel(const ShortTileRec* __restrict__ recs, int nTiles, PolyDev P, int S, int stageFloats) {
    extern __shared__ __align__(16) float smem[];
    float* __restrict__ Ws = smem;                                       // [TAPS][qpad]
    int* __restrict__ Bs = reinterpret_cast<int*>(Ws + TAPS * P.qpad);   // [qpad]
    float* __restrict__ ring = reinterpret_cast<float*>(Bs + P.qpad);    // kShortStages x stageFloats
    for (int i = threadIdx.x; i < TAPS * P.qpad; i += blockDim.x) Ws[i] = __ldg(P.W + i);
    for (int i = threadIdx.x; i < P.q; i += blockDim.x) Bs[i] = __ldg(P.B + i);
    const int tq = (int) threadIdx.x / P.q, tr = (int) threadIdx.x - tq * P.q;
    const int step = (S / P.q) * P.p;
    const int myTiles = ((int) blockIdx.x < nTiles) ? (nTiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;

    // Records are fetched a tile ahead of their use (registers), so their L2 latency overlaps the arithmetic of the tile before.
    struct IssueRec { const float* in; long long lA, inAvail; int nvec; };
    struct ComputeRec { float* out; int4 g; };                          // g = cnt, nvec, k0, mis
    auto load_issue = [&](int i) {
        IssueRec R; R.in = nullptr; R.lA = 0; R.inAvail = 0; R.nvec = -1;
        if (i < myTiles) {
            const ShortTileRec* r = recs + blockIdx.x + (size_t) i * gridDim.x;
            R.in = reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(&r->in)));
            R.lA = __ldg(&r->lA); R.inAvail = __ldg(&r->inAvail); R.nvec = __ldg(&r->nvec);
        }
        return R;
    };
    auto load_compute = [&](int i) {
        ComputeRec R; R.out = nullptr; R.g = make_int4(0, 0, 0, 0);
        if (i < myTiles) {
            const ShortTileRec* r = recs + blockIdx.x + (size_t) i * gridDim.x;
            R.out = reinterpret_cast<float*>(__ldg(reinterpret_cast<const unsigned long long*>(&r->out)));
            R.g = __ldg(reinterpret_cast<const int4*>(&r->cnt));
        }
        return R;
    };
    auto issue = [&](int i, const IssueRec& R) {
        if (R.nvec >= 0) {
            float* xs = ring + (i % kShortStages) * stageFloats;
            if (R.lA >= 0 && R.lA + 4LL * R.nvec <= R.inAvail) {
                const float* __restrict__ src = R.in + R.lA;
                const uint32_t dst = (uint32_t) __cvta_generic_to_shared(xs);
                for (int v = threadIdx.x; v < R.nvec; v += blockDim.x)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u * (uint32_t) v), "l"(src + 4 * v) : "memory");
            } else {
                float4* __restrict__ xs4 = reinterpret_cast<float4*>(xs);
                for (int v = threadIdx.x; v < R.nvec; v += blockDim.x) {
                    const long long l = R.lA + 4 * (long long) v;
                    float4 x4;
                    if (l >= 0 && l + 3 < R.inAvail) x4 = __ldg(reinterpret_cast<const float4*>(R.in + l));
                    else {
                        x4.x = (l >= 0 && l < R.inAvail) ? __ldg(R.in + l) : 0.f;             x4.y = (l + 1 >= 0 && l + 1 < R.inAvail) ? __ldg(R.in + l + 1) : 0.f;
                        x4.z = (l + 2 >= 0 && l + 2 < R.inAvail) ? __ldg(R.in + l + 2) : 0.f; x4.w = (l + 3 >= 0 && l + 3 < R.inAvail) ? __ldg(R.in + l + 3) : 0.f;
                    }
                    xs4[v] = x4;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");            // one group per tile, empty or not: uniform counting
    };

    #pragma unroll
    for (int i = 0; i < kShortStages - 1; ++i) issue(i, load_issue(i));
    IssueRec In = load_issue(kShortStages - 1);
    ComputeRec Cn = load_compute(0);
    for (int i = 0; i < myTiles; ++i) {
        asm volatile("cp.async.wait_group %0;" :: "n"(kShortStages - 2) : "memory");   // this thread's copies of tile i have landed
        __syncthreads();                                                 // ... everyone's; and everyone is done with tile i - 1
        const IssueRec Ic = In;     In = load_issue(i + kShortStages);
        const ComputeRec Cc = Cn;   Cn = load_compute(i + 1);
        issue(i + kShortStages - 1, Ic);                                 // into the stage tile i - 1 just left
        if ((int) threadIdx.x < S) {
            float* __restrict__ out = Cc.out;
            const int kk = Cc.g.z + tr;
            const int wrap = kk >= P.q ? 1 : 0;
            const int k = kk - wrap * P.q;
            float w[TAPS];
            #pragma unroll
            for (int j = 0; j < TAPS; ++j) w[j] = Ws[j * P.qpad + k];
            const float* __restrict__ x = ring + (i % kShortStages) * stageFloats + Cc.g.w + (tq + wrap) * P.p + Bs[k] - Bs[Cc.g.z];
            #pragma unroll 4
            for (int o = threadIdx.x; o < Cc.g.x; o += S, x += step) {
                float acc = 0.0f;
                #pragma unroll
                for (int j = 0; j < TAPS; ++j) acc = fmaf(x[j], w[j], acc);
                out[o] = acc;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
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

// --------------------------------------------------------------------------------------------- banded, register-tiled
// WindowedSinc is FP32-FMA bound (400 FLOP per output against 5-13 bytes), so this kernel is built like an SGEMM
// micro-kernel rather than a streaming copy:
//   * outputs are indexed (period a, slot k): n = a*q + k reads the inputs ending at a*p + B[k];
//   * a warp owns one group of TK adjacent slots and 32*TA periods; lane l holds periods l, l+32, ... so the TK
//     weights of a step are the same for the whole warp: one broadcast LDS.128 per 4 weights;
//   * the group walks one shared window of Tmax input samples; slot j's taps sit at offset B[k]-B[g*TK] in it
//     (zero weights elsewhere), so each input register feeds TK FFMAs;
//   * inputs live in shared memory as whole periods with an odd stride (Pstride) so the 32 lanes of a load hit
//     32 different banks for any p (p = 320 would otherwise be a 32-way conflict).
// Per step and warp: TA LDS.32 + TK/4 broadcast LDS.128 wavefronts for TA*TK FFMAs.
// Stage `total` consecutive input samples starting at channel index gStart into period rows of Xs
// (element e -> Xs[(e / p) * Pstride + e % p]).  Global reads are 16-byte vectors on the address's own alignment
// grid (the channel pointer is only float aligned once trimLatency's offset is folded in), four vectors in
// flight per thread; samples outside the segment's window read as zero.
__device__ __forceinline__ void stage_rows(const Seg& S, long long gStart, int total, int p, int Pstride, float* __restrict__ Xs) {
    const long long l0 = gStart - S.inOffset;                       // window index of element 0 (may be negative)
    const long long addrEl = (long long) (reinterpret_cast<uintptr_t>(S.in) >> 2) + l0;
    const int a0 = (int) (((addrEl % 4) + 4) % 4);
    const int head = min(total, (4 - a0) & 3);
    if ((int) threadIdx.x < head) {
        const int e = threadIdx.x;
        Xs[(e / p) * Pstride + (e % p)] = load_in(S, gStart + e);
    }
    const int nq = (total - head) >> 2;
    for (int v0 = threadIdx.x; v0 < nq; v0 += 4 * 256) {
        float4 val[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = v0 + k * 256;
            val[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nq) {
                const long long l = l0 + head + 4LL * v;
                if (l >= 0 && l + 3 < S.inAvail) val[k] = __ldg(reinterpret_cast<const float4*>(S.in + l));
                else {
                    val[k].x = load_in(S, gStart + head + 4 * v);     val[k].y = load_in(S, gStart + head + 4 * v + 1);
                    val[k].z = load_in(S, gStart + head + 4 * v + 2); val[k].w = load_in(S, gStart + head + 4 * v + 3);
                }
            }
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = v0 + k * 256;
            if (v < nq) {
                const int e = head + 4 * v;
                int al = e / p, u = e - al * p;
                float* row = Xs + al * Pstride;
                row[u] = val[k].x; if (++u == p) { u = 0; row += Pstride; }
                row[u] = val[k].y; if (++u == p) { u = 0; row += Pstride; }
                row[u] = val[k].z; if (++u == p) { u = 0; row += Pstride; }
                row[u] = val[k].w;
            }
        }
    }
    const int tail0 = head + 4 * nq;
    if (tail0 + (int) threadIdx.x < total) {
        const int e = tail0 + threadIdx.x;
        Xs[(e / p) * Pstride + (e % p)] = load_in(S, gStart + e);
    }
}

template <int TA, int TK>
__global__ void __launch_bounds__(256, 1)
banded_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles, BandedDev P,
              int PB, int GB, int nGB, int halo, int Pstride, int adding, float gain) {
    extern __shared__ __align__(16) float smem[];
    const int Tmax = P.Tmax, p = P.p, q = P.q;
    float* Cs = smem;                                      // [GB][Tmax][TK]
    float* Xs = smem + (size_t) GB * Tmax * TK;            // [(halo + NPc + 1) periods][Pstride]
    const int NPc = 32 * TA * PB;

    // Persistent CTA: gridDim.x is a multiple of nGB, so every tile this CTA visits has the same group block and
    // the weights are staged exactly once (contiguous in global, 16-byte copies).
    const int gbIdx = blockIdx.x % nGB;
    {
        const float4* __restrict__ src = reinterpret_cast<const float4*>(P.C + (size_t) gbIdx * GB * Tmax * TK);
        float4* dst = reinterpret_cast<float4*>(Cs);
        const int n4 = GB * Tmax * TK / 4;
        for (int i = threadIdx.x; i < n4; i += 256) dst[i] = __ldg(src + i);
    }

  for (int tileId = blockIdx.x; tileId < nTiles; tileId += gridDim.x) {
    const int sidx = find_seg(tilePrefix, nSegs, tileId);
    const Seg S = segs[sidx];
    const int tile = tileId - tilePrefix[sidx];
    const int pbIdx = tile / nGB;
    const long long aFirst = S.n0 / q;
    const long long A0 = aFirst + (long long) pbIdx * NPc;  // first period of this tile

    __syncthreads();                                        // previous tile's readers are done with Xs
    stage_rows(S, (A0 - halo) * (long long) p, (NPc + halo + 1) * p, p, Pstride, Xs);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int item = warp; item < GB * PB; item += 8) {
        const int gl = item % GB, pl = item / GB;
        const int g = gbIdx * GB + gl;
        const int wmin = __ldg(P.wmin + g);
        // window start inside the period grid: wmin = ashift*p + u0 with 0 <= u0 < p (wmin may be negative)
        int ashift = wmin / p; if (ashift * p > wmin) --ashift;
        int u = wmin - ashift * p;
        const float* xp = Xs + (size_t) (halo + pl * 32 * TA + lane + ashift) * Pstride + u;
        const float* cp = Cs + (size_t) gl * Tmax * TK;

        // Accumulators are float2 pairs over adjacent slots: Blackwell issues a scalar FFMA every other cycle but a packed
        // FFMA2 (fma.rn.f32x2) at the same rate, so the pairs are what reaches the FP32 peak.  (x, x) * (c_j, c_j+1).
        float2 acc[TA][TK / 2];
        #pragma unroll
        for (int i = 0; i < TA; ++i)
            #pragma unroll
            for (int j = 0; j < TK / 2; ++j) acc[i][j] = make_float2(0.0f, 0.0f);

        // The window is walked in runs that stay inside one period row, so the hot loop carries no boundary test.
        for (int t = 0; t < Tmax;) {
            const int run = min(p - u, Tmax - t);
            #pragma unroll 2
            for (int r = 0; r < run; ++r) {
                float2 x[TA];
                #pragma unroll
                for (int i = 0; i < TA; ++i) { const float v = xp[(size_t) i * 32 * Pstride + r]; x[i] = make_float2(v, v); }
                float2 c[TK / 2];
                #pragma unroll
                for (int j4 = 0; j4 < TK / 4; ++j4) {
                    const float4 v = *reinterpret_cast<const float4*>(cp + (size_t) r * TK + 4 * j4);
                    c[2 * j4] = make_float2(v.x, v.y); c[2 * j4 + 1] = make_float2(v.z, v.w);
                }
                #pragma unroll
                for (int i = 0; i < TA; ++i)
                    #pragma unroll
                    for (int j = 0; j < TK / 2; ++j) acc[i][j] = __ffma2_rn(x[i], c[j], acc[i][j]);
            }
            t += run; cp += (size_t) run * TK;
            xp += run + (Pstride - p); u = 0;               // next period row (warp-uniform)
        }

        // outputs: n = a*q + g*TK + j
        #pragma unroll
        for (int i = 0; i < TA; ++i) {
            const long long a = A0 + pl * 32 * TA + lane + 32 * i;
            const long long nBase = a * q + (long long) g * TK - S.n0;      // offset into the segment's outputs
            #pragma unroll
            for (int j = 0; j < TK; ++j) {
                const long long o = nBase + j;
                if (g * TK + j < q && o >= 0 && o < S.numOut) {
                    float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
                    if (adding) v = __fadd_rn(S.out[o], __fmul_rn(gain, v));
                    S.out[o] = v;
                }
            }
        }
    }
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
template <int TAPS, int NST>
cudaError_t run_short_st(const ResampleLaunch& L, ShortTileRec* recs, int grid, cudaStream_t s) {
    cudaError_t e = set_smem(short_kernel<TAPS, NST>, L.short_smem);
    if (e != cudaSuccess) return e;
    short_kernel<TAPS, NST><<<grid, L.short_threads, L.short_smem, s>>>(recs, L.n_tiles, L.poly, L.short_S, L.short_stage_floats);
    return cudaGetLastError();
}
template <int TAPS>
cudaError_t run_short(const ResampleLaunch& L, cudaStream_t s, long long* launches) {
    if (!L.d_tile_recs) return cudaErrorInvalidValue;
    ShortTileRec* recs = reinterpret_cast<ShortTileRec*>(L.d_tile_recs);
    if (!L.recs_ready) {
        short_tile_table_kernel<TAPS><<<(L.n_tiles + 255) / 256, 256, 0, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.poly, L.tile_out, recs);
        ++*launches;
    }
    const int perSm = std::max(1, std::min(2048 / L.short_threads, (int) ((227 * 1024) / (L.short_smem + 1024))));
    const int grid = std::min(L.n_tiles, L.sm_count * perSm);
    cudaError_t e;
    switch (L.short_stages) {
        case 2: e = run_short_st<TAPS, 2>(L, recs, grid, s); break;
        case 3: e = run_short_st<TAPS, 3>(L, recs, grid, s); break;
        case 4: e = run_short_st<TAPS, 4>(L, recs, grid, s); break;
        default: return cudaErrorInvalidValue;
    }
    if (e == cudaSuccess) ++*launches;
    return e;
}
static_assert(sizeof(ShortTileRec) == kShortTileRecBytes, "ShortTileRec layout");
template <int KIND>
cudaError_t run_generic(const ResampleLaunch& L, size_t smem, cudaStream_t s) {
    cudaError_t e = set_smem(generic_kernel<KIND>, smem);
    if (e != cudaSuccess) return e;
    generic_kernel<KIND><<<L.n_tiles, kRsThreads, smem, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.ratio, L.pos0,
                                                             L.d_sinc_table, L.tile_out, L.adding, L.gain);
    return cudaGetLastError();
}
template <int TA, int TK>
cudaError_t run_banded(const ResampleLaunch& L, cudaStream_t s) {
    cudaError_t e = set_smem(banded_kernel<TA, TK>, L.banded_smem);
    if (e != cudaSuccess) return e;
    const int perSm = L.banded_smem <= 110 * 1024 ? 2 : 1;           // co-resident CTAs overlap staging with compute
    int grid = std::min(L.n_tiles, std::max(L.sm_count * perSm, L.nGB));
    grid -= grid % L.nGB;
    banded_kernel<TA, TK><<<grid, 256, L.banded_smem, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.band, L.PB, L.GB, L.nGB,
                                                           L.halo, L.Pstride, L.adding, L.gain);
    return cudaGetLastError();
}
template <int TK>
cudaError_t run_banded_ta(const ResampleLaunch& L, cudaStream_t s) {
    switch (L.TA) {
        case 1: return run_banded<1, TK>(L, s);
        case 2: return run_banded<2, TK>(L, s);
        case 4: return run_banded<4, TK>(L, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

int choose_tile_out(double ratio) {
    double t = 8000.0 / (ratio > 0.25 ? ratio : 0.25);
    int tile = (int) t / 256 * 256;
    if (tile < 256) tile = 256;
    if (tile > 4096) tile = 4096;
    return tile;
}

long long resample_ctas_for_segment(const ResampleLaunch& L, long long n0, long long numOut) {
    if (numOut <= 0) return 0;
    if (L.hankel) return hankel_tiles_for_segment(n0, numOut);
    if (L.umma) {
        const long long q = L.um.q;
        const long long aFirst = n0 / q, aLast = (n0 + numOut - 1) / q;
        long long blocks = (aLast - aFirst + 1 + 127) / 128;
        if (L.um_tma && L.um_cta2 && L.um.nGB > 1) blocks += blocks & 1;       // CTA pairs walk two period blocks of the same slot block
        return blocks * L.um.nGB;
    }
    if (!L.banded) return (numOut + L.tile_out - 1) / L.tile_out;
    const long long q = L.band.q;
    const long long aFirst = n0 / q, aLast = (n0 + numOut - 1) / q;
    const long long NPc = 32LL * L.TA * L.PB;
    return ((aLast - aFirst + 1 + NPc - 1) / NPc) * L.nGB;
}

int resample_build_tiles(ResampleLaunch& L, const Seg* segs, int n, std::vector<int>* prefix) {
    prefix->assign((size_t) n + 1, 0);
    long long total = 0;
    if (L.umma) {
        // row pieces start at in + (n0/q + 128*pb + row)*p + U0 - inOffset floats: all on 16 bytes iff p % 4 == 0 (U0 is a
        // multiple of 16) and every segment's (in - inOffset) is
        // ... or every segment's is the same 1 .. 3 floats past one: the tables built for that shift (K origins = -shift mod 16)
        // put the rows back on 16 bytes, where the plain tables would leave the launch to the register loader (1.50 ms against
        // 0.64 ms on config 2's batch with untrimmed captures on 16 bytes: bench.py --unaligned)
        bool aligned = (L.um.p & 3) == 0;
        int shift = -1;
        for (int i = 0; i < n && aligned; ++i) {
            if (segs[i].numOut <= 0) continue;
            if ((reinterpret_cast<uintptr_t>(segs[i].in) & 3) != 0) { aligned = false; break; }
            const int m = (int) (((long long) (reinterpret_cast<uintptr_t>(segs[i].in) >> 2) - segs[i].inOffset) & 3);
            if (shift < 0) shift = m; else if (shift != m) aligned = false;
        }
        static const DiagOpts kNoDiag;
        const DiagOpts& D = L.diag ? *L.diag : kNoDiag;
        if (aligned && shift > 0) {
            UmmaDev shifted;
            if (L.ctx && !D.has("F9_UMMA_NOSHIFT") && L.ctx->get_umma(L.kind, L.um.p, L.um.q, L.um.NB, L.um.GBL, &shifted, shift) == F9_OK &&
                (shifted.poolN > 0) == (L.um.poolN > 0) && shifted.nGB == L.um.nGB) L.um = shifted;
            else aligned = false;
        }
        L.um_aligned = aligned && !D.has("F9_UMMA_UNALIGNED");
        // TMA feed (aligned rows only): tensor maps over the address range the segments read, ring of raw fp32 boxes
        L.um_tma = false;
        if (L.um_aligned && !D.has("F9_UMMA_NOTMA")) {
            if (umma_encode_maps(segs, n, L.um.p, &L.um_maps, D.has("F9_UMMA_NORANGES"))) {
                // CTA pairs halve the weights per SM: worth it when the weights leave a single CTA only a shallow input ring
                // (with several slot blocks only when the TMEM operand ring has four stages: measured slower otherwise, 147/160)
                const bool cta2 = L.um.blk[0].w2Off[0] >= 0 && L.sm_count >= 2 * L.um.nGB && !D.has("F9_UMMA_NOCTA2") &&
                                  (L.um.nGB == 1 || L.um.aSlots == 4) &&
                                  (umma_smem_bytes(L.um.maxEntries, L.um.NB, 4, true) > 227 * 1024 || D.has("F9_UMMA_CTA2"));
                int stages = 2;
                const int maxStages = D.get("F9_UMMA_STAGES", 8);
                while (stages < maxStages && umma_smem_bytes(L.um.maxEntries, L.um.NB, stages + 1, true, cta2) <= 227 * 1024) ++stages;
                stages &= ~1;                                   // even: a ring position always belongs to the same converter team (f9_umma.cu)
                if (umma_smem_bytes(L.um.maxEntries, L.um.NB, stages, true, cta2) <= 227 * 1024) {
                    L.um_tma = true; L.um_cta2 = cta2; L.um_stages = stages; L.um_smem = umma_smem_bytes(L.um.maxEntries, L.um.NB, stages, true, cta2);
                }
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        total += resample_ctas_for_segment(L, segs[i].n0, segs[i].numOut);
        if (total > 0x7fffffffLL) return -1;
        (*prefix)[(size_t) i + 1] = (int) total;
    }
    return (int) total;
}

cudaError_t launch_resample(const ResampleLaunch& L, cudaStream_t s, long long* launches) {
    if (L.n_tiles <= 0) return cudaSuccess;
    cudaError_t e;
    if (L.hankel) return launch_hankel(L, s, launches);
    if (L.umma) return launch_umma(L, s, launches);
    if (L.banded) {
        switch (L.band.TK) {
            case 8:  e = run_banded_ta<8>(L, s); break;
            case 12: e = run_banded_ta<12>(L, s); break;
            case 16: e = run_banded_ta<16>(L, s); break;
            case 20: e = run_banded_ta<20>(L, s); break;
            default: return cudaErrorInvalidValue;
        }
        if (e == cudaSuccess) ++*launches;
        return e;
    }
    const int taps = interp_memory(L.kind);
    const size_t smem = sizeof(float) * ((size_t) ((double) L.tile_out * L.ratio) + (size_t) taps + 8);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (L.rational && L.short_S > 0 && !L.adding) {
        switch (L.kind) {
            case F9_LAGRANGE: e = run_short<5>(L, s, launches); break;
            case F9_CATMULL_ROM: e = run_short<4>(L, s, launches); break;
            case F9_LINEAR: e = run_short<2>(L, s, launches); break;
            case F9_ZERO_ORDER_HOLD: e = run_short<1>(L, s, launches); break;
            default: return cudaErrorInvalidValue;
        }
        return e;
    }
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

// ------------------------------------------------------------------------------------------------ planning (host)
using namespace f9;

int f9_context::get_umma(int kind, long long p, long long q, int NB, int GBL, UmmaDev* out, int shift) {
    UmmaKey key{kind, p, q, NB, GBL, sinc_epoch, shift};
    auto it = umma_cache.find(key);
    if (it != umma_cache.end()) { *out = it->second; return F9_OK; }
    UmmaHost H;
    if (!build_umma(kind, sinc_table.data(), p, q, NB, GBL, &H, shift)) return fail(F9_ERR_INVALID, "umma table build failed");
    UmmaDev D; D.p = H.p; D.q = H.q; D.taps = H.taps; D.NB = H.NB; D.G = H.G; D.GBL = H.GBL; D.nGB = H.nGB; D.maxEntries = H.maxEntries; D.maxNK = H.maxNK;
    for (int b = 0; b < kUmmaMaxBlocks; ++b) D.blk[b] = H.blk[b];
    std::memcpy(D.gStart, H.gStart, sizeof(D.gStart)); std::memcpy(D.gSteps, H.gSteps, sizeof(D.gSteps)); std::memcpy(D.gTile, H.gTile, sizeof(D.gTile));
    D.poolN = H.poolN; D.split = H.split; D.aSlots = H.aSlots;
    uint8_t* dW = nullptr;
    F9_TRY_CUDA(this, cudaMalloc((void**) &dW, H.W.size()));
    F9_TRY_CUDA(this, cudaMemcpy(dW, H.W.data(), H.W.size(), cudaMemcpyHostToDevice));
    D.W = dW;
    umma_cache[key] = D;
    *out = D;
    return F9_OK;
}

int f9_context::get_hankel(int kind, int L, HankelDev* out) {
    const auto key = std::make_tuple(kind, L, sinc_epoch);
    auto it = hankel_cache.find(key);
    if (it != hankel_cache.end()) { *out = it->second; return F9_OK; }
    std::vector<uint8_t> image; int KS = 0;
    if (!build_hankel(kind, sinc_table.data(), L, &image, &KS)) return fail(F9_ERR_INVALID, "hankel table build failed");
    HankelDev D; D.L = L; D.KS = KS; D.rowBytes = 2 * (128 / L);
    D.layout = D.rowBytes == 128 ? 2 : D.rowBytes == 64 ? 4 : D.rowBytes == 32 ? 6 : 0;
    { const int taps = interp_memory(kind), shift = 209 - taps, R = 128 / L;      // lane i's filter is centred on t = shift + i + taps/2
      D.cLo = std::max(0, (shift + taps / 2 - 3) / 16); D.cHi = std::min(KS - 1, (shift + taps / 2 + 3 + R - 1) / 16); }
    D.elems = hankel_tile_elems(L, KS); D.bufBytes = (D.elems * 2 + 1023) / 1024 * 1024;
    uint8_t* dW = nullptr;
    F9_TRY_CUDA(this, cudaMalloc((void**) &dW, image.size()));
    F9_TRY_CUDA(this, cudaMemcpy(dW, image.data(), image.size(), cudaMemcpyHostToDevice));
    D.W = dW;
    hankel_cache[key] = D;
    *out = D;
    return F9_OK;
}

int f9_context::get_banded(int kind, long long p, long long q, int TK, int Gpad, BandedDev* out) {
    BandKey key{kind, p, q, TK, Gpad, sinc_epoch};
    auto it = band_cache.find(key);
    if (it != band_cache.end()) { *out = it->second; return F9_OK; }
    BandedHost H;
    build_banded(kind, sinc_table.data(), p, q, TK, Gpad, &H);
    BandedDev D; D.p = H.p; D.q = H.q; D.taps = H.taps; D.TK = H.TK; D.G = H.G; D.Gpad = H.Gpad; D.Tmax = H.Tmax;
    F9_TRY_CUDA(this, cudaMalloc((void**) &D.C, sizeof(float) * H.C.size()));
    F9_TRY_CUDA(this, cudaMalloc((void**) &D.wmin, sizeof(int) * H.wmin.size()));
    F9_TRY_CUDA(this, cudaMemcpy(D.C, H.C.data(), sizeof(float) * H.C.size(), cudaMemcpyHostToDevice));
    F9_TRY_CUDA(this, cudaMemcpy(D.wmin, H.wmin.data(), sizeof(int) * H.wmin.size(), cudaMemcpyHostToDevice));
    band_cache[key] = D;
    *out = D;
    return F9_OK;
}

// Pick the kernel for (kind, ratio, pos0).  WindowedSinc at a rational ratio from reset state goes to the
// register-tiled banded kernel; other rational cases to poly_kernel; everything else to generic_kernel.
int f9_context::prepare_resample(int kind, double ratio, double pos0, bool allow_rational, ResampleLaunch* L) {
    if (interp_memory(kind) == 0) return fail(F9_ERR_INVALID, "unknown interpolator kind");
    if (!(ratio > 0.0) || !std::isfinite(ratio)) return fail(F9_ERR_INVALID, "speed ratio must be positive and finite");
    *L = ResampleLaunch();
    L->kind = kind; L->ratio = ratio; L->pos0 = pos0; L->diag = &diag; L->ctx = this;
    const DiagOpts& D = diag;
    L->d_sinc_table = d_sinc_table;
    L->tile_out = choose_tile_out(ratio);
    long long p = 0, q = 0;
    if (allow_rational && pos0 == 1.0 && find_rational(ratio, 4096, &p, &q) && p <= (1 << 20)) {
        int rc = get_poly(kind, p, q, &L->poly); if (rc) return rc;
        L->rational = true;
        L->sm_count = sm_count;
        // short kinds: bandwidth-bound on CUDA cores with the slot's weights in registers (F9_SHORT_UMMA=1: tensor-core kernel instead)
        // Integer upsampling of the long kinds: Hankel-operand kernel (every input sample staged and converted once)
        if (p == 1 && (q == 2 || q == 4 || q == 8 || q == 16) && interp_memory(kind) > 5 && !D.has("F9_NO_UMMA") && !D.has("F9_NO_HANKEL")) {
            rc = get_hankel(kind, (int) q, &L->hk); if (rc) return rc;
            if (hankel_smem_bytes(L->hk) <= 227 * 1024) { L->hankel = true; return F9_OK; }
        }
        // Integer decimation (q == 1, p >= 2) stays on the tensor-core kernel when it is enabled: the short kernel's stride-p shared
        // loads conflict p ways there (measured 82 % / 76 % against 93 % / 95 % of the HBM roofline at 2:1 / 4:1); everywhere else
        // the short kernel is as fast or faster (44.1 -> 48 k: 86 % against 66 %; 48 -> 192 k: 84 % against 65 %) and exact fp32.
        const bool decim = q == 1 && p >= 2 && !D.has("F9_NO_UMMA") && !D.has("F9_SHORT_ALL");
        if (interp_memory(kind) <= 5 && q <= 256 && p <= 8192 && !decim && !D.has("F9_SHORT_UMMA") && !D.has("F9_NO_SHORT")) {
            const int c = (int) std::max(1LL, 256 / q);
            const int S = (int) q * c;
            long long I = (4096 * q + (long long) S * p / 2) / ((long long) S * p);
            I = std::max(2LL, std::min(64LL, I));
            if (D.has("F9_SHORT_I")) I = std::max(1, D.get("F9_SHORT_I"));          // experiments: steps per thread
            const long long stageFloats = (((S * I * p + q - 1) / q + interp_memory(kind) + 8) + 3) / 4 * 4;
            const int nst = std::max(2, std::min(4, D.get("F9_SHORT_STAGES", 3)));
            const size_t smem = sizeof(float) * (size_t) (nst * stageFloats + (interp_memory(kind) + 1) * L->poly.qpad);
            L->short_stages = nst;
            if (smem <= 200 * 1024) {
                L->short_S = S; L->short_threads = (S + 31) / 32 * 32;
                L->tile_out = (int) (S * I); L->short_smem = smem; L->short_stage_floats = (int) stageFloats;
                return F9_OK;
            }
        }
        if (!D.has("F9_NO_UMMA") && interp_memory(kind) >= 2) {
            long long bm = 0; int bGBL = 0, bNB = 0;
            umma_choose_plan(interp_memory(kind), p, q, &bm, &bNB, &bGBL, D.get("F9_UMMA_NB", 0));
            if (bm > 0 && D.has("F9_UMMA_GBL")) bGBL = std::max(1, std::min(D.get("F9_UMMA_GBL"), kUmmaMaxGroups));     // experiments: groups per slot block
            if (bm > 0) {
                rc = get_umma(kind, p * bm, q * bm, bNB, bGBL, &L->um);
                if (rc == F9_OK) {
                    int stages = 2;
                    while (stages < 4 && umma_smem_bytes(L->um.maxEntries, L->um.NB, stages + 1) <= 227 * 1024) ++stages;
                    L->um_stages = stages; L->um_smem = umma_smem_bytes(L->um.maxEntries, L->um.NB, stages);
                    L->umma = true;
                    return F9_OK;
                }
                if (rc != F9_ERR_INVALID) return rc;                                // tables this kernel cannot express: CUDA-core paths
            }
        }
        if (kind == F9_WINDOWED_SINC && !D.has("F9_NO_BANDED")) {
            // scale p/q so a group of slots is at least 16 wide, then choose TK / TA / block shape for shared memory
            long long m = 1;
            while (q * m < 16) m *= 2;
            const long long ps = p * m, qs = q * m;
            const int taps = 200;
            const size_t budget = 200 * 1024;
            double bestScore = -1.0; int bTK = 0, bTA = 0, bPB = 0, bGB = 0, bnGB = 0, bHalo = 0, bPs = 0; size_t bSmem = 0;
            int fTK = 0, fTA = 0, fnGB = 0;                      // F9_BANDED_CFG="TK,TA,nGB": force a configuration (experiments)
            if (D.has("F9_BANDED_TK")) { fTK = D.get("F9_BANDED_TK"); fTA = D.get("F9_BANDED_TA"); fnGB = D.get("F9_BANDED_NGB"); }
            for (int TK : {20, 16, 12, 8}) {
                if (fTK && TK != fTK) continue;
                const int G = (int) ((qs + TK - 1) / TK);
                for (int nGB = (G + 7) / 8; nGB <= std::min(G, (G + 7) / 8 + 3); ++nGB) {
                    if (fnGB && nGB != fnGB) continue;
                    const int GB = (G + nGB - 1) / nGB;
                    const int PB = std::max(1, 8 / GB);
                    const int maxShift = (int) (((long long) (TK - 1) * ps) / qs) + 1;
                    const int Tmax = (taps + maxShift + 1) & ~1;
                    const int halo = (int) ((taps - 1 + ps - 1) / ps);
                    const int Ps = (int) ((ps & 1) ? ps : ps + 1);
                    for (int TA : {4, 2, 1}) {
                        if (fTA && TA != fTA) continue;
                        const size_t smem = sizeof(float) * ((size_t) GB * Tmax * TK + (size_t) (32 * TA * PB + halo + 1) * Ps) + 64;
                        if (smem > budget) continue;
                        const double useful = (double) qs * taps / ((double) nGB * GB * TK * Tmax);  // band + slot padding
                        const double warps = (double) (GB * PB) / (8.0 * ((GB * PB + 7) / 8));       // warp balance
                        const double smemRate = std::min(1.0, (TA * TK / 4.0) / (TA + TK / 4.0));    // FFMA clk / smem clk
                        const double issue = (double) (TA * TK) / (TA * TK + TA + TK / 4 + 2);       // FFMA share of issue slots
                        const double restage = 1.0 / (1.0 + 0.15 * (nGB - 1));                       // rows staged once per group block
                        const double overlap = smem <= 110 * 1024 ? 1.25 : 1.0;                      // two CTAs per SM hide the staging
                        const double score = useful * warps * smemRate * issue * restage * overlap;
                        if (score > bestScore) { bestScore = score; bTK = TK; bTA = TA; bPB = PB; bGB = GB; bnGB = nGB; bHalo = halo; bPs = Ps; bSmem = smem; }
                    }
                }
            }
            if (bestScore > 0.0 && ps <= 8192) {
                rc = get_banded(kind, ps, qs, bTK, bGB * bnGB, &L->band); if (rc) return rc;
                L->banded = true; L->TA = bTA; L->PB = bPB; L->GB = bGB; L->nGB = bnGB; L->halo = bHalo; L->Pstride = bPs;
                L->banded_smem = bSmem; L->sm_count = sm_count;
            }
        }
    }
    if (!L->banded && (double) L->tile_out * ratio > 45000.0) return fail(F9_ERR_UNSUPPORTED, "speed ratio too large for the tile buffer");
    return F9_OK;
}
