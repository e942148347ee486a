// Reductions of the detection / trimming path: |x| argmax with the reference's tie-breaking, RMS / peak
// window statistics, the offline reverb-tail scan and the bounded-lag cross-correlation.
// All are HBM-bound streaming reads (4 B per sample); the work is warp-shuffle + block reductions.
#include "f9_internal.cuh"

namespace f9 {

namespace {

constexpr int kPeakThreads = 256;
constexpr int kPeakChunk = 16384;          // frames per CTA (64 KB of input)

__device__ __forceinline__ bool peak_better(float v, int ch, int pos, float bv, int bch, int bpos) {
    // order of MainComponent::findPeakPosition (Source/MainComponent.cpp:950-975): strict '>' while scanning
    // channel 0 first => larger value wins, ties go to the lower channel, then the earlier frame.
    if (v > bv) return true;
    if (v < bv) return false;
    if (ch != bch) return ch < bch;
    return pos < bpos;
}

// ---- stage 1: one CTA per (buffer, channel, chunk) ---------------------------------------------------
// STATS: the same pass also sums the squares (float product rounded to float, widened, double sum: calculateRMS, Source/
// MainComponent.cpp:983-1004) -- findPeakPosition and calculateNoiseFloorDb read the same capture right after each other in the
// latency measurement (:270-279), so one read serves both.
template <bool STATS>
__global__ void __launch_bounds__(kPeakThreads)
peak_partial_kernel(const DevBuf* __restrict__ bufs, const int* __restrict__ prefix, int n, PeakPartial* __restrict__ partials,
                    double* __restrict__ psum) {
    // locate the buffer of this CTA: prefix[b] <= blockIdx.x < prefix[b+1]
    int lo = 0, hi = n;
    const int bid = blockIdx.x;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (prefix[mid] <= bid) lo = mid; else hi = mid; }
    const DevBuf B = bufs[lo];
    const int local = bid - prefix[lo];
    const int chunksPerCh = (B.numFrames + kPeakChunk - 1) / kPeakChunk;
    const int ch = local / chunksPerCh;
    const int chunk = local - ch * chunksPerCh;
    const int start = chunk * kPeakChunk;
    const int len = min(kPeakChunk, B.numFrames - start);
    const float* __restrict__ x = B.base + (long long) ch * B.chStride + start;

    float bv = 0.0f; int bpos = -1;
    double sq = 0.0;
    // head to 16-byte alignment, float4 body, scalar tail; each thread visits increasing indices so a strict
    // '>' keeps its earliest maximum.
    const int mis = (int) ((reinterpret_cast<uintptr_t>(x) >> 2) & 3);
    const int head = min(len, (4 - mis) & 3);
    if ((int) threadIdx.x < head) {
        const float xv0 = x[threadIdx.x];
        float a = fabsf(xv0);
        if (a > bv) { bv = a; bpos = threadIdx.x; }
        if (STATS) sq += (double) __fmul_rn(xv0, xv0);
    }
    const int nvec = (len - head) >> 2;
    const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + head);
    #pragma unroll 4
    for (int i = threadIdx.x; i < nvec; i += kPeakThreads) {
        const float4 v = __ldg(xv + i);
        const int p = head + 4 * i;
        float a;
        a = fabsf(v.x); if (a > bv) { bv = a; bpos = p; }
        a = fabsf(v.y); if (a > bv) { bv = a; bpos = p + 1; }
        a = fabsf(v.z); if (a > bv) { bv = a; bpos = p + 2; }
        a = fabsf(v.w); if (a > bv) { bv = a; bpos = p + 3; }
        if (STATS) sq += ((double) __fmul_rn(v.x, v.x) + (double) __fmul_rn(v.y, v.y)) + ((double) __fmul_rn(v.z, v.z) + (double) __fmul_rn(v.w, v.w));
    }
    const int tail0 = head + 4 * nvec;
    if (tail0 + (int) threadIdx.x < len) {
        // tail indices are larger than everything in the body: strict '>' again keeps the earliest
        const float xt = x[tail0 + threadIdx.x];
        float a = fabsf(xt);
        if (a > bv) { bv = a; bpos = tail0 + threadIdx.x; }
        if (STATS) sq += (double) __fmul_rn(xt, xt);
    }
    // a thread's head index is smaller than its body indices only for threads < head; handled by ordering above.

    // warp reduce (value desc, position asc; -1 = none, loses to any real position with equal value only when v == 0)
    unsigned bposu = (bpos < 0) ? 0x7fffffffu : (unsigned) bpos;
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const unsigned op = __shfl_xor_sync(0xffffffffu, bposu, off);
        if (ov > bv || (ov == bv && op < bposu)) { bv = ov; bposu = op; }
    }
    __shared__ float sv[kPeakThreads / 32];
    __shared__ unsigned sp[kPeakThreads / 32];
    __shared__ double ss[kPeakThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (STATS) {
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);      // fixed tree: deterministic
        if (lane == 0) ss[warp] = sq;
    }
    if (lane == 0) { sv[warp] = bv; sp[warp] = bposu; }
    __syncthreads();
    if (STATS && threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kPeakThreads / 32; ++w) t += ss[w];
        psum[bid] = t;
    }
    if (warp == 0) {
        bv = (lane < kPeakThreads / 32) ? sv[lane] : 0.0f;
        bposu = (lane < kPeakThreads / 32) ? sp[lane] : 0x7fffffffu;
        #pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const unsigned op = __shfl_xor_sync(0xffffffffu, bposu, off);
            if (ov > bv || (ov == bv && op < bposu)) { bv = ov; bposu = op; }
        }
        if (lane == 0) {
            PeakPartial r;
            r.v = bv; r.ch = ch;
            r.pos = (bposu == 0x7fffffffu || bv == 0.0f) ? -1 : (int) bposu + start;
            partials[bid] = r;
        }
    }
}

// ---- calculateRMS bit for bit ------------------------------------------------------------------------
// The reference sums the squares sequentially in double (Source/MainComponent.cpp:985-997); the kernels sum them as a tree.
// All terms are >= 0, so either order is within (n - 1) 2^-53 relative of the exact sum, the two sums within twice that of each
// other, and sqrt(sum / n) within about (n + 4) 2^-53.  (float) sqrt(sum / n) -- what calculateRMS returns -- therefore only
// depends on the order when the double value lies that close to a float rounding boundary (stage 1: about one buffer in a
// thousand at 5 s stereo).  Stage 2 tightens the bound for those: partial sums never exceed the total, so a term that is a
// multiple of the total's ulp is added WITHOUT rounding in either order (the terms are float squares: 24 significant bits);
// only the `small` terms below that grid, the <= 64 steps where the running sum crosses a power of two, and the <= ~100 adds
// of the tree's depth (~100 + n / 16384: a thread's own chain) can round at all, each by at most an ulp of the total, which moves
// sqrt(sum / n) by half of that.  The CTA counts the small terms in parallel and re-tests with (small + 256 + n / 16384) 2^-53:
// about one buffer in 10^4..10^5 is left, and only that one is re-summed in the reference's order by one thread.  The returned
// sum of squares is then the reference's own double.
__device__ __forceinline__ bool rms_rounds_differently(double s, long long count, double relBound) {
    if (!(s != 0.0)) return false;                               // all samples zero: every order gives exactly 0
    const double r = sqrt(s / (double) count);
    return (float) (r * (1.0 - relBound)) != (float) (r * (1.0 + relBound));   // NaN compares unequal: redone in order, NaN again
}
__device__ __forceinline__ bool rms_order_dependent(double s, long long count) {
    return rms_rounds_differently(s, count, (double) (count + 8) * 1.1102230246251565e-16);
}
// Terms of the buffer that can be rounded when added to a partial sum < 4 * 2^ilogb(s): float squares with bits below the ulp of
// the binade above s.  Called by all threads of a CTA; returns this thread's count.  Integer arithmetic on the float's bits.
__device__ __forceinline__ int small_term(float v, int Eb) {
    const unsigned pb = __float_as_uint(__fmul_rn(v, v));
    const int ep = (int) (pb >> 23);
    const unsigned mp = ep ? ((pb & 0x7fffffu) | 0x800000u) : (pb & 0x7fffffu);
    const int shf = (ep ? ep : 1) - 150 - (Eb - 1075);                              // p / grid = mp * 2^shf
    return (mp != 0u && shf < 0 && (-shf >= 32 || (mp & ((1u << -shf) - 1u)) != 0u)) ? 1 : 0;
}
__device__ __forceinline__ long long count_small_terms(const DevBuf& B, double s) {
    const int Eb = (int) (((unsigned long long) __double_as_longlong(s)) >> 52) + 1;       // biased exponent of the binade above s
    int small = 0;                                               // per thread: < 2^31 terms
    for (int c = 0; c < B.numCh; ++c) {
        // one CTA reads a whole channel: 128-bit loads, four of them in flight per thread (scalar loads one at a time made this
        // pass, not the ordered sum behind it, the longest part of a flagged buffer)
        const float* __restrict__ x = B.base + (long long) c * B.chStride;
        const int mis = (int) ((reinterpret_cast<uintptr_t>(x) >> 2) & 3);
        const int head = min(B.numFrames, (4 - mis) & 3);
        if ((int) threadIdx.x < head) small += small_term(__ldg(x + threadIdx.x), Eb);
        const int nvec = (B.numFrames - head) >> 2;
        const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + head);
        #pragma unroll 4
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            const float4 v = __ldg(xv + i);
            small += small_term(v.x, Eb) + small_term(v.y, Eb) + small_term(v.z, Eb) + small_term(v.w, Eb);
        }
        const int tail0 = head + 4 * nvec;
        if (tail0 + (int) threadIdx.x < B.numFrames) small += small_term(__ldg(x + tail0 + threadIdx.x), Eb);
    }
    return small;
}
// The reference's chain, s <- fl(s + p_k) in order, evaluated by one CTA without walking it term by term.  A chain of dependent
// double additions costs ~35 clk per term on this part (measured: 8-12 ms for a 5 s stereo capture), but the chain has structure:
// the terms are >= 0, so s only grows, and while s stays in one binade [2^E, 2^(E+1)) it is an integer M on the grid g = 2^(E-52)
// and a step is integer arithmetic: p_k / g = q + f, M <- M + q + [f > 1/2] + [f == 1/2] ((M + q) mod 2) (round to nearest, ties
// to even).  The increment depends on M only through its parity, so a term is a pair (increment if M is even, increment if M is
// odd) and pairs compose associatively, order preserved.  Every warp reduces its own block of 256 terms on the grid of the
// current binade (8 compositions per lane + a 5-step warp scan, all integer); the blocks' pairs are then applied in order while M
// stays below 2^53 (s has not left the binade: the sum is monotone).  The block that crosses a binade -- at most ~60 per buffer --
// is walked term by term by warp 0 with real double additions, and the next round starts behind it on the new grid.  Every
// operation is exact, so the result IS the sequential sum, bit for bit (forced on every test buffer through F9_RMS_FORCE_ORDER).
// A sum of squares is never negative: the sign bit marks "re-sum this buffer in the reference's order" between the final kernels
// and order_resum_kernel (a NaN sum is marked too and comes out as the chain's NaN).
__device__ __forceinline__ double mark_for_resum(double s) { return __longlong_as_double(__double_as_longlong(s) | (long long) 0x8000000000000000ull); }
__device__ __forceinline__ bool marked_for_resum(double s) { return __double_as_longlong(s) < 0; }
struct IncPair { long long e, o; };                              // increment of M when M is even / odd
__device__ __forceinline__ IncPair inc_compose(IncPair a, IncPair c) {         // a first, then c
    IncPair r;
    r.e = a.e + ((a.e & 1) ? c.o : c.e);
    r.o = a.o + (((1 + a.o) & 1) ? c.o : c.e);
    return r;
}
constexpr int kSeqWarps = 32, kSeqPer = 8, kSeqBlock = 32 * kSeqPer;     // order_resum_kernel: 1024 threads, 8 terms per lane; a round is 8192 terms
struct SeqShared { IncPair pair[kSeqWarps]; int ok[kSeqWarps]; unsigned long long sb; int consumed, nDone; long long Mend; };
// The kSeqPer terms of one lane, channel-major (the reference's scan order), squared in float (:995); zeros past the end.
// All loads are issued before the first use and without a branch between them: a guarded load per term put every load in its own
// basic block, each waiting for the one before it (a round cost one memory latency PER TERM: 3.8 us for 8 terms).
__device__ __forceinline__ void seq_load_terms(const DevBuf& B, long long idx0, long long total, float (&pf)[kSeqPer]) {
    int c = idx0 < total ? (int) (idx0 / B.numFrames) : 0;
    int off = idx0 < total ? (int) (idx0 - (long long) c * B.numFrames) : 0;
    const int left = (int) max(0LL, min((long long) kSeqPer, total - idx0));    // terms of this lane that exist
    float v[kSeqPer];
    #pragma unroll
    for (int j = 0; j < kSeqPer; ++j) {
        const float* ptr = B.base + (j < left ? (long long) c * B.chStride + off : 0LL);      // past the end: any address of the buffer
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v[j]) : "l"(ptr));
        const bool wrap = off + 1 == B.numFrames;
        off = wrap ? 0 : off + 1;
        c += wrap ? 1 : 0;
    }
    #pragma unroll
    for (int j = 0; j < kSeqPer; ++j) pf[j] = j < left ? __fmul_rn(v[j], v[j]) : 0.0f;
}
// Call with all kSeqWarps * 32 threads of the CTA; every thread returns the sum.
// A round covers kSeqWarps blocks of kSeqBlock terms (32 per lane: the scan, the barriers and the in-order application are per round).  The loads of the next round are issued before this round's arithmetic (the round
// almost always advances by all its blocks; a round that stops early reloads), and a block that leaves the binade is cut at the
// LANE whose terms cross it: the lanes before it are applied as integers, that lane's terms take real double additions, and
// the next round starts right behind them on the new grid (the first version walked the whole block and waited for every round's
// loads: 0.66 ms per 5 s stereo capture, a third of it in the ~60 crossings).
__device__ __noinline__ double sum_squares_in_reference_order(const DevBuf& B, SeqShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long total = (long long) B.numCh * B.numFrames;
    const long long mineOff = (long long) warp * kSeqBlock + (long long) lane * kSeqPer;
    unsigned long long sb = 0ull;                                // bits of the running sum s (a non-negative double)
    long long b0 = 0, bNext = -1;                                // start of this round; start the prefetched terms belong to
    float pf[kSeqPer], nf[kSeqPer];
    while (b0 < total) {
        if (bNext == b0) {
            #pragma unroll
            for (int j = 0; j < kSeqPer; ++j) pf[j] = nf[j];
        } else seq_load_terms(B, b0 + mineOff, total, pf);
        bNext = b0 + (long long) kSeqWarps * kSeqBlock;
        if (bNext < total) seq_load_terms(B, bNext + mineOff, total, nf);      // in flight during the round
        const int Eb = (int) (sb >> 52);                          // biased exponent of s
        bool ok = Eb >= 64 && Eb < 2046;                          // s normal and its grid a normal double too
        IncPair mine{0, 0};
        #pragma unroll
        for (int j = 0; j < kSeqPer; ++j) {
            // p = mp * 2^(max(ep, 1) - 150);  grid g = 2^(Eb - 1075);  p / g = mp * 2^shf.  Branch-free (the lanes' exponents differ:
            // the branchy form diverged at every term, ~300 instructions per term) and 32-bit except for the final shift.
            const unsigned pb = __float_as_uint(pf[j]);
            const int ep = (int) ((pb >> 23) & 0xffu);            // biased float exponent (0: zero or subnormal)
            const unsigned mp = (pb & 0x7fffffu) | (ep ? 0x800000u : 0u);
            const int shf = max(ep, 1) - 150 - (Eb - 1075);
            ok = ok && ep != 255 && shf <= 29;                    // Inf / NaN: the plain chain reproduces them; shf > 29: the term alone reaches the next binade
            const int k = min(max(-shf, 0), 31), ls = min(max(shf, 0), 29);     // k >= 26: below a quarter of the grid, rounds away (q = 0, rem < half)
            const unsigned q = mp >> k, rem = mp & ((1u << k) - 1u), half = (1u << k) >> 1;
            const unsigned up = rem > half ? 1u : 0u, tie = (rem == half && k > 0) ? 1u : 0u;    // tie: to even
            IncPair e;
            e.e = (long long) ((unsigned long long) (q + up + (tie & q & 1u)) << ls);
            e.o = (long long) ((unsigned long long) (q + up + (tie & ~q & 1u)) << ls);
            mine = inc_compose(mine, e);
        }
        ok = __all_sync(0xffffffffu, ok);
        #pragma unroll
        for (int off = 1; off < 32; off <<= 1) {                  // ordered reduction over the lanes (inclusive scan)
            IncPair prev;
            prev.e = __shfl_up_sync(0xffffffffu, mine.e, off);
            prev.o = __shfl_up_sync(0xffffffffu, mine.o, off);
            if (lane >= off) mine = inc_compose(prev, mine);
        }
        if (lane == 31) { sh.pair[warp] = mine; sh.ok[warp] = ok ? 1 : 0; }
        __syncthreads();
        // apply the blocks in order while s stays in its binade: a lane per block, an inclusive scan of the pairs (composition is
        // associative), the first block that cannot be applied found with a ballot -- by warp 0 alone, then published (with every
        // thread walking the blocks one after the other this loop was 40 % of the kernel's instructions)
        const long long M0 = (long long) ((sb & 0x000fffffffffffffull) | 0x0010000000000000ull);
        if (warp == 0) {
            const bool live = lane < kSeqWarps && b0 + (long long) lane * kSeqBlock < total;
            IncPair pr = live ? sh.pair[lane] : IncPair{0, 0};
            const bool okb = live && sh.ok[lane] != 0;
            #pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                IncPair prev;
                prev.e = __shfl_up_sync(0xffffffffu, pr.e, off);
                prev.o = __shfl_up_sync(0xffffffffu, pr.o, off);
                if (lane >= off) pr = inc_compose(prev, pr);
            }
            const long long Ml = M0 + ((M0 & 1) ? pr.o : pr.e);           // M after blocks 0 .. lane (monotone in the lane)
            const unsigned bad = __ballot_sync(0xffffffffu, !okb || Ml >= (1LL << 53));    // (blocks past the end count as bad: they stop the run)
            const int nd = bad ? __ffs((int) bad) - 1 : 32;
            const long long Mend = __shfl_sync(0xffffffffu, Ml, nd > 0 ? nd - 1 : 0);
            if (lane == 0) { sh.nDone = nd; sh.Mend = nd > 0 ? Mend : M0; }
        }
        __syncthreads();
        const int done = sh.nDone;
        const long long M = sh.Mend;
        if (done > 0) sb = ((unsigned long long) Eb << 52) | ((unsigned long long) M & 0x000fffffffffffffull);
        const bool crossing = done < kSeqWarps && b0 + (long long) done * kSeqBlock < total;
        int consumed = 0;
        if (crossing) {
            // Block `done` leaves the binade (or s is still zero / subnormal, or a term is not finite).  Warp `done` still holds its
            // terms and its lanes' inclusive prefixes (`mine`), all on the old grid.
            if (warp == done) {
                double s; int took;
                if (ok) {
                    // M after lane l's terms; the sum is monotone, so the first lane at or above 2^53 is the one whose terms cross
                    const long long Ml = M + ((M & 1) ? mine.o : mine.e);
                    const int L = __ffs((int) __ballot_sync(0xffffffffu, Ml >= (1LL << 53))) - 1;     // the block's total is >= 2^53: L >= 0
                    long long Mprev = __shfl_up_sync(0xffffffffu, Ml, 1);
                    if (lane == 0) Mprev = M;
                    const long long Mstart = __shfl_sync(0xffffffffu, Mprev, L);                      // < 2^53: exact state in front of lane L
                    s = __longlong_as_double((long long) (((unsigned long long) Eb << 52) | ((unsigned long long) Mstart & 0x000fffffffffffffull)));
                    #pragma unroll
                    for (int j = 0; j < kSeqPer; ++j) s = __dadd_rn(s, (double) __shfl_sync(0xffffffffu, pf[j], L));
                    took = (L + 1) * kSeqPer;
                } else {
                    // the plain chain over the block's terms (every lane the same s)
                    s = __longlong_as_double((long long) sb);
                    bool zeros = true;                            // leading digital silence: s + 0 = s, nothing to walk
                    #pragma unroll
                    for (int j = 0; j < kSeqPer; ++j) zeros = zeros && pf[j] == 0.0f;
                    #pragma unroll 1
                    for (int l = 0; l < (__all_sync(0xffffffffu, zeros) ? 0 : 32); ++l) {
                        #pragma unroll
                        for (int j = 0; j < kSeqPer; ++j) s = __dadd_rn(s, (double) __shfl_sync(0xffffffffu, pf[j], l));
                    }
                    took = kSeqBlock;
                }
                if (lane == 0) { sh.sb = (unsigned long long) __double_as_longlong(s); sh.consumed = took; }
            }
            __syncthreads();
            sb = sh.sb; consumed = sh.consumed;
        }
        __syncthreads();                                          // sh.pair / sh.ok / sh.sb are rewritten by the next round
        b0 += (long long) done * kSeqBlock + consumed;
    }
    return __longlong_as_double((long long) sb);
}

// ---- stage 2: one CTA per buffer folds its partials in scan order ----------------------------------
__global__ void __launch_bounds__(256)
peak_final_kernel(const PeakPartial* __restrict__ partials, const int* __restrict__ prefix, float threshold, int* __restrict__ out_pos,
                  const double* __restrict__ psum, double* __restrict__ sumsq, float* __restrict__ peakv, const DevBuf* __restrict__ bufs, int forceOrder) {
    const int b = blockIdx.x;
    const int p0 = prefix[b], p1 = prefix[b + 1];
    float bv = 0.0f; int bch = 0x7fffffff, bpos = 0x7fffffff;
    double sq = 0.0;
    __shared__ double sTotal;
    for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
        const PeakPartial r = partials[i];
        if (r.pos >= 0 && peak_better(r.v, r.ch, r.pos, bv, bch, bpos)) { bv = r.v; bch = r.ch; bpos = r.pos; }
        if (psum) sq += psum[i];
    }
    __shared__ double ssq[8];
    if (psum) {
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if ((threadIdx.x & 31) == 0) ssq[threadIdx.x >> 5] = sq;
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oc = __shfl_xor_sync(0xffffffffu, bch, off);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, off);
        if (peak_better(ov, oc, op, bv, bch, bpos)) { bv = ov; bch = oc; bpos = op; }
    }
    __shared__ float sv[8]; __shared__ int sc[8], sp[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = bv; sc[warp] = bch; sp[warp] = bpos; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int) blockDim.x / 32; ++w)
            if (peak_better(sv[w], sc[w], sp[w], bv, bch, bpos)) { bv = sv[w]; bch = sc[w]; bpos = sp[w]; }
        out_pos[b] = (bv > threshold && bpos != 0x7fffffff) ? bpos : -1;
        if (psum) { sTotal = ((ssq[0] + ssq[1]) + (ssq[2] + ssq[3])) + ((ssq[4] + ssq[5]) + (ssq[6] + ssq[7])); if (peakv) peakv[b] = bv; }
    }
    if (psum) {                                                  // bit-exact sum of squares: see rms_order_dependent (all threads of the CTA)
        __syncthreads();
        const DevBuf B = bufs[b];
        const double total = sTotal;
        const long long count = (long long) B.numCh * B.numFrames;
        int redo = forceOrder > 0 ? 2 : (forceOrder == 0 && rms_order_dependent(total, count) ? 1 : 0);        // uniform across the CTA; forceOrder < 0: tree sum as it is
        if (redo == 1) {
            long long small = count_small_terms(B, total);
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) small += __shfl_xor_sync(0xffffffffu, small, off);
            __shared__ long long ssm[8];
            if ((threadIdx.x & 31) == 0) ssm[threadIdx.x >> 5] = small;
            __syncthreads();
            small = ((ssm[0] + ssm[1]) + (ssm[2] + ssm[3])) + ((ssm[4] + ssm[5]) + (ssm[6] + ssm[7]));
            redo = rms_rounds_differently(total, count, (double) (small + 256 + count / 16384) * 1.1102230246251565e-16) ? 2 : 0;
        }
        if (threadIdx.x == 0) sumsq[b] = (redo == 2) ? mark_for_resum(total) : total;        // order_resum_kernel replaces a marked sum
    }
}

// ---- per-buffer sum of squares (double) + peak -----------------------------------------------------
// calculateRMS (Source/MainComponent.cpp:983-1004): float product, widened, double sum.  The tree order here
// differs from the sequential reference; the final kernels re-sum in the reference's order exactly when that could
// change the float the reference returns (rms_order_dependent), so (float) sqrt(sumsq / n) is the reference's value bit for bit.
constexpr int kStatThreads = 256;

__device__ __forceinline__ void block_sum_max(double& s, float& m, double* sh_s, float* sh_m) {
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh_s[warp] = s; sh_m[warp] = m; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        s = (lane < nw) ? sh_s[lane] : 0.0;
        m = (lane < nw) ? sh_m[lane] : 0.0f;
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void accum_range(const float* __restrict__ x, int len, double& s, float& m) {
    // float product rounded to float (no contraction), widened, accumulated in double
    const int mis = (int) ((reinterpret_cast<uintptr_t>(x) >> 2) & 3);
    const int head = min(len, (4 - mis) & 3);
    if ((int) threadIdx.x < head) { const float v = x[threadIdx.x]; s += (double) __fmul_rn(v, v); m = fmaxf(m, fabsf(v)); }
    const int nvec = (len - head) >> 2;
    const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + head);
    #pragma unroll 4
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        const float4 v = __ldg(xv + i);
        s += (double) __fmul_rn(v.x, v.x); s += (double) __fmul_rn(v.y, v.y);
        s += (double) __fmul_rn(v.z, v.z); s += (double) __fmul_rn(v.w, v.w);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    const int tail0 = head + 4 * nvec;
    if (tail0 + (int) threadIdx.x < len) { const float v = x[tail0 + threadIdx.x]; s += (double) __fmul_rn(v, v); m = fmaxf(m, fabsf(v)); }
}

constexpr int kStatChunk = 32768;
constexpr int kStatPartials = kStatPartialsPerBuf;

// grid: (chunks, n).  Partial sums are combined with double atomics?  No: order must be deterministic, so
// stage 1 writes per-chunk partials and stage 2 folds them in index order.
__global__ void __launch_bounds__(kStatThreads)
stats_partial_kernel(const DevBuf* __restrict__ bufs, double* __restrict__ psum, float* __restrict__ pmax, int maxChunks) {
    const DevBuf B = bufs[blockIdx.y];
    const long long total = (long long) B.numCh * B.numFrames;   // channel-major linear index
    const int chunksPerCh = (B.numFrames + kStatChunk - 1) / kStatChunk;
    const int nChunks = chunksPerCh * B.numCh;
    (void) total;
    double s = 0.0; float m = 0.0f;
    __shared__ double sh_s[kStatThreads / 32]; __shared__ float sh_m[kStatThreads / 32];
    for (int c = blockIdx.x; c < nChunks; c += gridDim.x) {
        const int ch = c / chunksPerCh, k = c - ch * chunksPerCh;
        const int start = k * kStatChunk, len = min(kStatChunk, B.numFrames - start);
        accum_range(B.base + (long long) ch * B.chStride + start, len, s, m);
    }
    block_sum_max(s, m, sh_s, sh_m);
    if (threadIdx.x == 0) {
        psum[(size_t) blockIdx.y * maxChunks + blockIdx.x] = s;
        pmax[(size_t) blockIdx.y * maxChunks + blockIdx.x] = m;
    }
}
// One CTA per buffer (grid = n): fold the partials in index order, then the two-stage order test of rms_order_dependent.
__global__ void __launch_bounds__(256)
stats_final_kernel(const double* __restrict__ psum, const float* __restrict__ pmax, int maxChunks,
                   double* __restrict__ sumsq, float* __restrict__ peak, int n, const DevBuf* __restrict__ bufs, int forceOrder) {
    const int b = blockIdx.x;
    if (b >= n) return;
    double s = 0.0; float m = 0.0f;
    for (int i = 0; i < maxChunks; ++i) { s += psum[(size_t) b * maxChunks + i]; m = fmaxf(m, pmax[(size_t) b * maxChunks + i]); }   // every thread: the same value
    const DevBuf B = bufs[b];
    const long long count = (long long) B.numCh * B.numFrames;
    int redo = forceOrder > 0 ? 2 : (forceOrder == 0 && rms_order_dependent(s, count) ? 1 : 0);
    if (redo == 1) {
        long long small = count_small_terms(B, s);
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) small += __shfl_xor_sync(0xffffffffu, small, off);
        __shared__ long long ssm[8];
        if ((threadIdx.x & 31) == 0) ssm[threadIdx.x >> 5] = small;
        __syncthreads();
        small = ((ssm[0] + ssm[1]) + (ssm[2] + ssm[3])) + ((ssm[4] + ssm[5]) + (ssm[6] + ssm[7]));
        redo = rms_rounds_differently(s, count, (double) (small + 256 + count / 16384) * 1.1102230246251565e-16) ? 2 : 0;
    }
    if (threadIdx.x == 0) { sumsq[b] = (redo == 2) ? mark_for_resum(s) : s; peak[b] = m; }
}

// The buffers whose sum the final kernels marked, re-summed in the reference's order: one CTA of 1024 threads each (the walk is
// bound by its own instruction stream: 32 warps take 0.2 ms for a 5 s stereo capture where the final kernels' 8 took 0.6 ms), all
// other CTAs leave at once.
__global__ void __launch_bounds__(kSeqWarps * 32)
order_resum_kernel(const DevBuf* __restrict__ bufs, double* __restrict__ sumsq) {
    const int b = blockIdx.x;
    if (!marked_for_resum(sumsq[b])) return;                     // uniform across the CTA
    __shared__ SeqShared seq;
    const DevBuf B = bufs[b];
    __syncthreads();                                             // every thread has read the mark before thread 0 overwrites it
    const double v = sum_squares_in_reference_order(B, seq);
    if (threadIdx.x == 0) sumsq[b] = v;
}

// ---- reverb-tail windows ------------------------------------------------------------------------------
// One CTA per (poll, buffer).  Poll i tests frames [e-window, e), e = start + (i+1)*hop
// (AudioProcessingService.swift:436-446).  Writes flag -1 skipped / 0 / 1.
__global__ void __launch_bounds__(kStatThreads)
tail_window_kernel(const DevBuf* __restrict__ bufs, const TailParams* __restrict__ params, int max_polls, int* __restrict__ flags) {
    const int b = blockIdx.y, i = blockIdx.x;
    const DevBuf B = bufs[b];
    const TailParams P = params[b];
    const long long e = P.startFrame + (long long) (i + 1) * P.hop;
    int* out = flags + (size_t) b * max_polls + i;
    if (e > B.numFrames) { if (threadIdx.x == 0) *out = -2; return; }      // past the data: no such poll
    if (e < P.window) { if (threadIdx.x == 0) *out = -1; return; }
    double s = 0.0; float m = 0.0f;
    __shared__ double sh_s[kStatThreads / 32]; __shared__ float sh_m[kStatThreads / 32];
    for (int c = 0; c < B.numCh; ++c)
        accum_range(B.base + (long long) c * B.chStride + (e - P.window), P.window, s, m);
    block_sum_max(s, m, sh_s, sh_m);
    if (threadIdx.x != 0) return;
    int below;
    if (P.mode == F9_TAIL_PEAK) {
        // Swift predicate (AudioProcessingService.swift:710-737); the max is order independent => exact.
        if (P.noNf) below = m < 0.0001f;
        else below = (m == 0.0f) ? P.below0 : (m <= P.rstar);
    } else {
        // C++ predicate (Source/MainComponent.cpp:863-882) folded to the linear domain on the host:
        // below <=> rms <= rstar.  rms from the tree sum can differ from the sequential reference by one float
        // ulp only when it lands next to rstar; those windows are re-summed in the reference's order.
        const long long count = (long long) B.numCh * P.window;
        float rms = (float) sqrt(s / (double) count);
        const float up = __int_as_float(__float_as_int(P.rstar) + 1);
        if (P.rstar >= 0.0f && (rms == P.rstar || rms == up)) {
            double ss = 0.0;
            for (int c = 0; c < B.numCh; ++c) {
                const float* x = B.base + (long long) c * B.chStride + (e - P.window);
                for (int k = 0; k < P.window; ++k) ss = __dadd_rn(ss, (double) __fmul_rn(x[k], x[k]));
            }
            rms = (float) sqrt(ss / (double) count);
        }
        below = (P.rstar >= 0.0f) && (rms <= P.rstar);
    }
    *out = below;
}

// One thread per buffer walks its flags: `required` consecutive silent polls => stop frame.
__global__ void tail_runs_kernel(const DevBuf* __restrict__ bufs, const TailParams* __restrict__ params, int n, int max_polls,
                                 const int* __restrict__ flags, long long* __restrict__ stop) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const TailParams P = params[b];
    const DevBuf B = bufs[b];
    int consecutive = 0; long long st = -1;
    for (int i = 0; i < max_polls; ++i) {
        const long long e = P.startFrame + (long long) (i + 1) * P.hop;
        if (e > B.numFrames) break;
        const int f = flags[(size_t) b * max_polls + i];
        if (f < 0) continue;                              // skipped poll leaves the counter alone
        consecutive = f ? consecutive + 1 : 0;
        if (consecutive >= P.required) { st = e; break; }
    }
    stop[b] = st;
}


// ---- bounded-lag cross-correlation -----------------------------------------------------------------
// r_c[lag] = sum_i (double) x[i] * (double) y_c[i+lag], i ascending.  The float*float product is exact in
// double and fma(x, y, acc) rounds once, so a thread that walks i in order reproduces the sequential CPU
// sum bit for bit; the argmax below then needs no guard band.  One thread per lag, one CTA per
// (buffer, channel, tile of kXcLags lags); y staged through shared memory in stimulus-sized chunks.
constexpr int kXcThreads = 256;
constexpr int kXcR = 8;                          // consecutive lags per thread (register tile)
constexpr int kXcLags = kXcThreads * kXcR;       // lags per CTA
constexpr int kXcChunk = 512;                    // stimulus samples per shared-memory chunk
constexpr int kXcPlane = (kXcChunk + kXcLags) / kXcR + 1;

__device__ __forceinline__ bool xc_better(double v, int ch, int lag, double bv, int bch, int blag) {
    return v > bv || (v == bv && (ch < bch || (ch == bch && lag < blag)));
}

// Register-tiled direct form.  A thread owns kXcR consecutive lags and walks the stimulus in order (every sum is the same fma
// chain as the scalar code: bit-exact doubles, hence an exact argmax).  Per stimulus sample it needs one new recording
// sample: x and y are converted to double once when a chunk is staged, and y is stored in kXcR interleaved planes
// (element e in plane e % R at offset e / R) so that the "new sample" loads of a warp are consecutive doubles.  Per tap:
// one broadcast load, one conflict-free load, kXcR DFMAs (the one-lag-per-thread version spent two loads and two
// float->double conversions per DFMA: 4.5 TFLOP/s).
__global__ void __launch_bounds__(kXcThreads)
xcorr_partial_kernel(const DevBuf* __restrict__ bufs, const int* __restrict__ prefix, int n,
                     const float* __restrict__ stim, int stimLen, int lagMin, int lagMax, XcPartial* __restrict__ partials, const int* __restrict__ need) {
    int lo = 0, hi = n;
    const int bid = blockIdx.x;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (prefix[mid] <= bid) lo = mid; else hi = mid; }
    if (need && !need[lo]) return;                              // f9_xcorr.cu: only the buffers whose candidate list overflowed are scanned exactly
    const DevBuf B = bufs[lo];
    const int local = bid - prefix[lo];
    const int nLags = lagMax - lagMin + 1;
    const int tiles = (nLags + kXcLags - 1) / kXcLags;
    const int ch = local / tiles, tile = local - ch * tiles;
    const int lag0 = lagMin + tile * kXcLags;
    const int t = threadIdx.x;
    const float* __restrict__ y = B.base + (long long) ch * B.chStride;

    __shared__ double xs[kXcChunk];
    __shared__ double ys[kXcR][kXcPlane];
    double acc[kXcR];
    #pragma unroll
    for (int r = 0; r < kXcR; ++r) acc[r] = 0.0;
    for (int c0 = 0; c0 < stimLen; c0 += kXcChunk) {
        const int clen = min(kXcChunk, stimLen - c0);
        for (int i = t; i < clen; i += kXcThreads) xs[i] = (double) stim[c0 + i];
        for (int e = t; e < clen + kXcLags; e += kXcThreads) {              // element e = recording sample c0 + lag0 + e
            const long long g = (long long) c0 + lag0 + e;
            ys[e % kXcR][e / kXcR] = (g >= 0 && g < B.numFrames) ? (double) y[g] : 0.0;
        }
        __syncthreads();
        // window w[r] = element i + kXcR * t + r; tap i multiplies w[0 .. R-1], then the window slides by one element
        double w[kXcR];
        #pragma unroll
        for (int r = 0; r < kXcR; ++r) w[r] = ys[r][t];                       // elements kXcR * t + r
        int i = 0;
        for (; i + kXcR <= clen; i += kXcR) {                                  // i stays a multiple of kXcR: static plane indices
            const int q = i / kXcR + t;
            #pragma unroll
            for (int u = 0; u < kXcR; ++u) {
                const double x = xs[i + u];
                #pragma unroll
                for (int r = 0; r < kXcR; ++r) acc[r] = fma(x, w[(u + r) % kXcR], acc[r]);
                w[u] = ys[u][q + 1];                                          // element i + u + kXcR * (t + 1): the window's next sample
            }
        }
        for (; i < clen; ++i) {                                                // chunk tail (clen not a multiple of kXcR)
            const double x = xs[i];
            #pragma unroll
            for (int r = 0; r < kXcR; ++r) {
                const int e = i + kXcR * t + r;
                acc[r] = fma(x, ys[e % kXcR][e / kXcR], acc[r]);
            }
        }
        __syncthreads();
    }
    // best of this thread's lags (ascending lag, strict >: ties keep the lower lag), then warp, then block
    double v = -1.0; int blag = 0x7fffffff;
    #pragma unroll
    for (int r = 0; r < kXcR; ++r) {
        const int lag = lag0 + kXcR * t + r;
        const double a = fabs(acc[r]);
        if (lag <= lagMax && a > v) { v = a; blag = lag; }
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int ol = __shfl_xor_sync(0xffffffffu, blag, off);
        if (ov > v || (ov == v && ol < blag)) { v = ov; blag = ol; }
    }
    __shared__ double sv[kXcThreads / 32]; __shared__ int sl[kXcThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sv[warp] = v; sl[warp] = blag; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wi = 1; wi < kXcThreads / 32; ++wi) if (sv[wi] > v || (sv[wi] == v && sl[wi] < blag)) { v = sv[wi]; blag = sl[wi]; }
        XcPartial r; r.v = v; r.ch = ch; r.lag = blag; r.pad = 0;
        partials[bid] = r;
    }
}

__global__ void xcorr_final_kernel(const XcPartial* __restrict__ partials, const int* __restrict__ prefix, int n, XcPartial* __restrict__ best,
                                   const int* __restrict__ need) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n || (need && !need[b])) return;
    // "maxValue starts at 0, strict >": a candidate must be > 0 to register (bch = -1 otherwise)
    double bv = 0.0; int bch = -1, blag = 0;
    for (int i = prefix[b]; i < prefix[b + 1]; ++i) {
        const XcPartial r = partials[i];
        if (r.v > 0.0 && (bch < 0 || xc_better(r.v, r.ch, r.lag, bv, bch, blag))) { bv = r.v; bch = r.ch; blag = r.lag; }
    }
    XcPartial o; o.v = bv; o.ch = bch; o.lag = blag; o.pad = 0;
    best[b] = o;
}

}  // namespace

// =============================================================================================== launchers
int peak_prefix(const DevBuf* h_bufs, int n, std::vector<int>* prefix) {
    prefix->assign((size_t) n + 1, 0);
    for (int i = 0; i < n; ++i) {
        const int chunks = (h_bufs[i].numFrames + kPeakChunk - 1) / kPeakChunk;
        (*prefix)[(size_t) i + 1] = (*prefix)[(size_t) i] + chunks * h_bufs[i].numCh;
    }
    return (*prefix)[(size_t) n];
}

cudaError_t launch_find_peak(const DevBuf* d_bufs, int n, int total_ctas, const int* d_prefix, float threshold,
                             PeakPartial* d_partials, int* d_out_pos, cudaStream_t s, long long* launches,
                             double* d_psum, double* d_sumsq, float* d_peakv, int forceOrder) {
    if (n <= 0) return cudaSuccess;
    if (total_ctas > 0) {
        if (d_psum) peak_partial_kernel<true><<<total_ctas, kPeakThreads, 0, s>>>(d_bufs, d_prefix, n, d_partials, d_psum);
        else peak_partial_kernel<false><<<total_ctas, kPeakThreads, 0, s>>>(d_bufs, d_prefix, n, d_partials, nullptr);
        ++*launches;
    }
    peak_final_kernel<<<n, 256, 0, s>>>(d_partials, d_prefix, threshold, d_out_pos, d_psum, d_sumsq, d_peakv, d_bufs, forceOrder);
    ++*launches;
    if (d_psum && forceOrder >= 0) { order_resum_kernel<<<n, kSeqWarps * 32, 0, s>>>(d_bufs, d_sumsq); ++*launches; }
    return cudaGetLastError();
}

cudaError_t launch_stats(const DevBuf* d_bufs, int n, double* d_psum, float* d_pmax,
                         double* d_sumsq, float* d_peak, cudaStream_t s, long long* launches, int forceOrder) {
    if (n <= 0) return cudaSuccess;
    dim3 grid(kStatPartials, n);
    stats_partial_kernel<<<grid, kStatThreads, 0, s>>>(d_bufs, d_psum, d_pmax, kStatPartials);
    ++*launches;
    stats_final_kernel<<<n, 256, 0, s>>>(d_psum, d_pmax, kStatPartials, d_sumsq, d_peak, n, d_bufs, forceOrder);
    ++*launches;
    if (forceOrder >= 0) { order_resum_kernel<<<n, kSeqWarps * 32, 0, s>>>(d_bufs, d_sumsq); ++*launches; }
    return cudaGetLastError();
}

cudaError_t launch_tail_scan(const DevBuf* d_bufs, const TailParams* d_params, int n, int max_polls,
                             long long* d_stop, int* d_flags, cudaStream_t s, long long* launches) {
    if (n <= 0) return cudaSuccess;
    if (max_polls > 0) {
        dim3 grid(max_polls, n);
        tail_window_kernel<<<grid, kStatThreads, 0, s>>>(d_bufs, d_params, max_polls, d_flags);
        ++*launches;
    }
    tail_runs_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_bufs, d_params, n, max_polls, d_flags, d_stop);
    ++*launches;
    return cudaGetLastError();
}

int xcorr_prefix(const DevBuf* h_bufs, int n, int lagMin, int lagMax, std::vector<int>* prefix) {
    const int tiles = (lagMax - lagMin + 1 + kXcLags - 1) / kXcLags;
    prefix->assign((size_t) n + 1, 0);
    for (int i = 0; i < n; ++i) (*prefix)[(size_t) i + 1] = (*prefix)[(size_t) i] + tiles * h_bufs[i].numCh;
    return (*prefix)[(size_t) n];
}

cudaError_t launch_xcorr(const DevBuf* d_bufs, int n, int total_ctas, const int* d_prefix, const float* d_stim, int stimLen,
                         int lagMin, int lagMax, XcPartial* d_partials, XcPartial* d_best, cudaStream_t s, long long* launches, const int* d_need) {
    if (n <= 0) return cudaSuccess;
    if (total_ctas > 0) {
        xcorr_partial_kernel<<<total_ctas, kXcThreads, 0, s>>>(d_bufs, d_prefix, n, d_stim, stimLen, lagMin, lagMax, d_partials, d_need);
        ++*launches;
    }
    xcorr_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_partials, d_prefix, n, d_best, d_need);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace f9
