// WindowedSinc (and any other kind) at integer upsampling ratios 1:L, L in {2, 4, 8, 16} -- config 3's 48 -> 192 kHz -- on
// tcgen05 with a *Hankel* operand.
//
// Output n = L*a + k reads x[a - (taps-1) .. a] with the weights of phase k.  Take a column of 128 consecutive outputs,
// n = 128*r + l, l = i*L + k (R = 128/L inputs per column).  Then
//     D[l, r] = sum_t  A[l, t] * X[r, t],     X[r, t] = x[R*r - 208 + t],     A[(i,k), t] = w_k[t - 9 - i]  (0 elsewhere),
// a GEMM with M = 128 (the lane IS the output's offset inside its column: no transpose in the epilogue, a warp stores 128
// contiguous bytes per column), N = columns, K = R + 208 rounded up to 16.  A is a constant weight image.  X is a Hankel
// matrix: row r is the input itself shifted by R samples, i.e. 2R bytes of fp16 = 16 / 32 / 64 / 128 bytes -- exactly the row
// pitch of a K-major shared-memory operand with no / 32 / 64 / 128-byte swizzle.  So the operand is the converted input
// stored ONCE, linearly (element m at swz(2m)), and K step s is the same buffer read from a start address 32*s bytes further
// on (tools/ubench/hankel_probe.cu: the swizzle is a function of the absolute address bits, base_offset 0 is right for every
// start).  Every input sample is fetched from HBM once, converted once and stored once; the polyphase kernel in f9_umma.cu
// stages and converts each sample (taps + p)/p = 6 times at this ratio.
//
// Precision as in f9_umma.cu: x' = 128 x = x0 + x1/2048, w = w0 + w1/2048 (fp16 parts); two accumulators per output, D0 += w0 x0
// and D1 += w1 x0 + w0 x1 (units of 1/2048), out = (D0 + D1/2048) / 128.  The tensor core truncates the fp32 accumulator after every
// MMA by an ulp of its magnitude, so the K steps are issued *tails first* (|w| < 0.11, the accumulator is still small) and the 2-5
// steps that hold the main lobe of some lane's filter (|w| up to 1) last: only those truncate at full scale (measured: inside the
// 2^-20 tolerance for 1:2 .. 1:16: 0.81-0.88 x 2^-20 against the oracle on noise of amplitude 0.5 (at 0 dBFS: <= 0.78 x 2^-20 against the exact value); window order gave 1.06 x 2^-20).  |x| >= 256,
// Inf or NaN raise the flag and hankel_redo_kernel recomputes the launch in fp32.
//
// Roles (448 threads, one persistent CTA per SM; tile = 64 columns = 8192 outputs):
//   setup        the weights (M-side operand, constant) go into TMEM once: 8 columns per K step and part (tcgen05.st), so the
//                MMAs run in TS mode and shared memory holds nothing but the input ring
//   warps 5-12   converters: thread-private cp.async copies of the tile's input span into a raw fp32 ring three tiles ahead (zeros
//                outside the segment's window), then fp16 head / tail split and 16-byte stores into the swizzled buffers of a
//                4-stage operand ring
//   warps 4, 13  issue 3 * K/16 MMAs (M = 128, N = 32) per tile each: warp 4 for columns 0-31, the last warp for columns 32-63, each into
//                its own accumulator sets -- two per half when the weights leave room (K <= 256: 1:4, 1:8, 1:16), so that the
//                epilogue of tile i overlaps the MMAs of tile i+1; one per half at 1:2
//   warps 0-3    epilogue: tcgen05.ld (thread = lane = output offset), combine, st.global (128 contiguous bytes per warp and column)
// Measured (B200, 512 channels of 10 s): 48 -> 192 k 1.07 ms = 70 % of the HBM roofline (polyphase kernel: 1.93 ms, 38.9 %),
// 48 -> 96 k 72 % (48.4 %); with the MMAs switched off the load / convert / store path alone runs at 94 %, the MMAs alone at 90 %
// (F9_HK_DBG; DESIGN.md 4.2 lists what was tried on the gap between the two).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>

#include "f9_internal.cuh"

namespace f9 {
namespace {

constexpr int kHkCols = 64;                 // columns (of 128 outputs) per tile = MMA N
// Compile-time knobs (A/B builds; measured flat for 3-6 raw / operand stages and 4-8 converter warps, worse with 12-16)
#ifndef F9_HK_STAGES
#define F9_HK_STAGES 4
#endif
#ifndef F9_HK_RAW
#define F9_HK_RAW 4
#endif
#ifndef F9_HK_CONV_WARPS
#define F9_HK_CONV_WARPS 8
#endif
constexpr int kHkStages = F9_HK_STAGES;     // converted-input ring (operand buffers)
constexpr int kHkRaw = F9_HK_RAW;           // raw fp32 ring: cp.async prefetch distance of the converters, in tiles
constexpr int kHkThreads = (6 + F9_HK_CONV_WARPS) * 32;   // warps 0-3 epilogue, 4 and the last one MMA issue (one half of the tile's columns each), 5.. converters
constexpr int kHkHalf = kHkCols / 2;        // columns per MMA (N) and accumulator set
// TMEM columns: weight heads at 0, tails at 8 KS (8 columns per K step and image); accumulator sets (D0, D1 of kHkHalf columns
// each) from column 256 (KS <= 16: two sets per column half, the epilogue of one tile overlaps the MMAs of the next) or 272
// (KS = 17, 1:2: one set per half).
template <int KS> struct HkTmem {
    static constexpr int aTail = 8 * KS;
    static constexpr int nBuf = 16 * KS <= 256 ? 2 : 1;
    static constexpr int dBase = 16 * KS <= 256 ? 256 : 272;
    static_assert(dBase + 2 * nBuf * 2 * kHkHalf <= 512 && 16 * KS <= dBase, "TMEM layout");
    static __device__ __forceinline__ int set(int h, int b) { return dBase + 2 * kHkHalf * (h * nBuf + b); }
};
constexpr int kHkConvWarps = F9_HK_CONV_WARPS, kHkFirstConv = 5, kHkIssuer1 = 5 + F9_HK_CONV_WARPS;
constexpr float kHkPre = 128.0f;            // 2^7 pre-scale, as f9_umma.cu
constexpr uint32_t kHkPark = 2000;

struct HankelTileRec {
    const float* in; float* out;   // the segment's window and output
    long long x0, inAvail;         // window index of operand element 0 (may be negative); window length
    long long oBase, numOut;       // output index (relative to out) of column 0 / lane 0; outputs of the segment
};
static_assert(sizeof(HankelTileRec) == kHankelTileRecBytes, "HankelTileRec layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// Bounded: the hardware may park the thread up to kHkPark ns per attempt; a barrier that never completes traps instead of hanging.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    #pragma unroll 1
    for (int i = 0; i < (1 << 22); ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kHkPark) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand descriptor.  layout 0: no swizzle, element (row r, k) at (k/8)*lbo + (r/8)*sbo + (r%8)*16 + (k%8)*2;
// layout 2 / 4 / 6: 128 / 64 / 32-byte swizzle, rows one swizzle width apart, sbo = 8 rows.  base_offset stays 0 (see the probe).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3fff);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) (layout & 7) << 61;
    return d;
}
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {        // fp16 x fp16 -> fp32, both K-major
    return (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ int find_seg(const int* __restrict__ prefix, int n, int bid) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= bid) lo = mid; else hi = mid; }
    return lo;
}
template <typename T> __device__ __forceinline__ T* ldg_ptr(T* const* p) {
    return reinterpret_cast<T*>(__ldg(reinterpret_cast<const unsigned long long*>(p)));
}
__device__ __forceinline__ float combine(uint32_t d0, uint32_t d1) {                  // (D0 + D1 / 2048) / 128: one rounding
    return fmaf(__uint_as_float(d1), 1.0f / (2048.0f * kHkPre), __uint_as_float(d0) * (1.0f / kHkPre));
}

// One thread per tile.  Columns are absolute: column c holds outputs 128c .. 128c+127 of the channel, so a segment that starts
// at n0 owns columns n0/128 .. (n0+numOut-1)/128 and its tiles are runs of kHkCols of them.
__global__ void __launch_bounds__(256)
hankel_tile_table_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles, int R,
                         HankelTileRec* __restrict__ recs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTiles) return;
    const int sidx = find_seg(tilePrefix, nSegs, t);
    const Seg S = segs[sidx];
    const long long r0 = S.n0 / 128 + (long long) (t - tilePrefix[sidx]) * kHkCols;
    HankelTileRec Rr;
    Rr.in = S.in; Rr.out = S.out;
    Rr.x0 = (long long) R * r0 - 208 - S.inOffset; Rr.inAvail = S.inAvail;
    Rr.oBase = 128 * r0 - S.n0; Rr.numOut = S.numOut;
    recs[t] = Rr;
}

template <int KS, int CLO, int CHI>
__global__ void __launch_bounds__(kHkThreads, 1)
hankel_fir_kernel(const HankelTileRec* __restrict__ recs, int nTiles, const __grid_constant__ HankelDev P, unsigned* __restrict__ ovf, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // the swizzle patterns repeat on 1024 bytes
    const uint32_t bufBytes = (uint32_t) P.bufBytes;                // kHkStages x (head buffer, tail buffer), 1024-aligned
    const uint32_t rawBytes = ((uint32_t) P.elems * 4u + 127u) & ~127u;
    uint8_t* raw = ring + kHkStages * 2 * bufBytes;                 // kHkRaw x rawBytes: the tiles' input spans as they come from HBM
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + kHkRaw * rawBytes);
    uint64_t *bFull = bars, *bEmpty = bars + kHkStages, *accFull = bars + 2 * kHkStages, *accEmpty = accFull + 4;     // [half][buffer]
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(accEmpty + 4);
    using TM = HkTmem<KS>;
    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int myTiles = (int) blockIdx.x < nTiles ? (nTiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;

    // ---- one-time setup: barriers, TMEM, the weights into TMEM (the M-side operand of every MMA: lane = output offset)
    if (threadIdx.x == 0) {
        for (int s = 0; s < kHkStages; ++s) { mbar_init(bFull + s, kHkConvWarps); mbar_init(bEmpty + s, 2); }
        for (int a = 0; a < 4; ++a) { mbar_init(accFull + a, 1); mbar_init(accEmpty + a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmemSlot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmemSlot;
    if (warp < 4) {
        const int l = warp * 32 + lane;
        const uint32_t laneBase = (uint32_t) (warp * 32) << 16;
        #pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const uint4* row = reinterpret_cast<const uint4*>(P.W + ((size_t) part * 128 + l) * (size_t) (KS * 32));   // KS * 16 fp16 per lane
            #pragma unroll 1
            for (int s = 0; s < KS; ++s) {
                const uint4 a = __ldg(row + 2 * s), b = __ldg(row + 2 * s + 1);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             :: "r"(tmem + laneBase + (uint32_t) (part * TM::aTail + 8 * s)), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= kHkFirstConv && warp < kHkFirstConv + kHkConvWarps) {
        // =========================================================== converters
        // Every sample is loaded exactly once: thread c owns the groups c, c + 256, ... (8 samples = 32 bytes) of every tile, copies
        // them with cp.async into its private places of a raw fp32 ring kHkRaw - 1 tiles ahead of the conversion (HBM latency under
        // load is longer than a tile), then splits them into fp16 head / tail and stores 16 bytes each into the swizzled operand
        // buffers.  16-byte copies when the span is aligned and inside the segment's window, 4-byte copies with zero fill otherwise.
        const int ctid = (warp - kHkFirstConv) * 32 + lane;
        const int groups = P.elems >> 3;
        constexpr int kG = 3;                                        // groups per thread and tile (elems <= 3 * 256 * 8)
        const uint32_t swzMask = P.rowBytes == 128 ? 7u : P.rowBytes == 64 ? 3u : P.rowBytes == 32 ? 1u : 0u;
        const uint32_t raw0 = smem_u32(raw);
        __half2 hmax = __floats2half2_rn(0.f, 0.f);
        struct TileIn { const float* in; long long x0, inAvail; };
        auto load_rec = [&](int j) {
            TileIn T = {nullptr, 0, 0};
            if (j < myTiles) { const HankelTileRec* r = recs + blockIdx.x + (size_t) j * gridDim.x; T.in = ldg_ptr(&r->in); T.x0 = __ldg(&r->x0); T.inAvail = __ldg(&r->inAvail); }
            return T;
        };
        auto issue = [&](int j, const TileIn& T) {
            if (j < myTiles && !(dbg & 2)) {
                const bool vec = (((long long) (reinterpret_cast<uintptr_t>(T.in) >> 2) + T.x0) & 3) == 0;
                const uint32_t slot = raw0 + (uint32_t) (j % kHkRaw) * rawBytes;
                #pragma unroll
                for (int u = 0; u < kG; ++u) {
                    const int g = ctid + u * kHkConvWarps * 32;
                    if (g < groups) {
                        const long long l = T.x0 + 8LL * g;
                        const uint32_t dst = slot + 32u * (uint32_t) g;
                        if (vec && l >= 0 && l + 7 < T.inAvail) {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(T.in + l) : "memory");
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 16u), "l"(T.in + l + 4) : "memory");
                        } else {
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const bool ok = l + e >= 0 && l + e < T.inAvail;
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst + 4u * e), "l"(ok ? T.in + l + e : T.in), "r"(ok ? 4 : 0) : "memory");
                            }
                        }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");            // one group per tile, empty or not: uniform counting
        };
        #pragma unroll
        for (int j = 0; j < kHkRaw - 1; ++j) issue(j, load_rec(j));
        TileIn Tn = load_rec(kHkRaw - 1);
        for (int i = 0; i < myTiles; ++i) {
            const TileIn Tc = Tn;
            Tn = load_rec(i + kHkRaw);                                       // consumed one iteration later
            issue(i + kHkRaw - 1, Tc);
            asm volatile("cp.async.wait_group %0;" :: "n"(kHkRaw - 1) : "memory");     // this thread's copies of tile i have landed
            const int st = i % kHkStages;
            if (i >= kHkStages) mbar_wait(bEmpty + st, (uint32_t) ((i / kHkStages - 1) & 1));       // the MMAs that read this stage are done
            const uint32_t hb = smem_u32(ring + (size_t) st * 2 * bufBytes), tb = hb + bufBytes;
            const uint32_t slot = raw0 + (uint32_t) (i % kHkRaw) * rawBytes;
            #pragma unroll
            for (int u = 0; u < kG; ++u) {
                const int g = ctid + u * kHkConvWarps * 32;
                if (g < groups && !(dbg & 2)) {
                    float v[8];
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(slot + 32u * (uint32_t) g) : "memory");
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(slot + 32u * (uint32_t) g + 16u) : "memory");
                    uint32_t hd[4], tl[4];
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float x0 = v[2 * e] * kHkPre, x1 = v[2 * e + 1] * kHkPre;
                        const __half2 h = __floats2half2_rn(x0, x1);
                        hmax = __hmax2_nan(hmax, __habs2(h));
                        const float2 hf = __half22float2(h);
                        const __half2 t = __floats2half2_rn((x0 - hf.x) * 2048.0f, (x1 - hf.y) * 2048.0f);      // exact differences
                        hd[e] = *reinterpret_cast<const uint32_t*>(&h); tl[e] = *reinterpret_cast<const uint32_t*>(&t);
                    }
                    const uint32_t a0 = hb + 16u * (uint32_t) g, a1 = tb + 16u * (uint32_t) g;
                    const uint32_t s0 = a0 ^ (((a0 >> 7) & swzMask) << 4), s1 = a1 ^ (((a1 >> 7) & swzMask) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(s0), "r"(hd[0]), "r"(hd[1]), "r"(hd[2]), "r"(hd[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(s1), "r"(tl[0]), "r"(tl[1]), "r"(tl[2]), "r"(tl[3]) : "memory");
                }
            }
            fence_async_smem();                                      // generic-proxy stores -> async-proxy reads of the MMAs
            __syncwarp();
            if (lane == 0) mbar_arrive(bFull + st);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const float2 hm = __half22float2(hmax);
        if (!(hm.x < 32768.0f) || !(hm.y < 32768.0f)) atomicOr(ovf, 1u);
    } else if (warp == 4 || warp == kHkIssuer1) {
        // =========================================================== MMA issue (TS mode: weights from TMEM, input from shared memory)
        // Warp 4 owns columns 0-31 of every tile, the last warp columns 32-63: 3 * KS MMAs of N = 32 each into their own accumulator sets.
        // Two accumulators per output: D0 += w0 x0, D1 += w1 x0 + w0 x1 (units of 1/2048).  The tensor core truncates the fp32
        // accumulator after every MMA by an ulp of its magnitude, so the K steps are issued tails first (|w| < 0.11: the accumulator is
        // still small) and the steps cLo..cHi that hold the main lobe of some lane's filter (|w| up to 1) last: only those 2-5 MMAs
        // truncate at full scale.  (Window order measured 1.06 x 2^-20 on noise at amplitude 0.5; this order passes with margin.)
        const int h = warp == 4 ? 0 : 1;
        const uint32_t el = elect_one();
        const uint32_t idesc = make_idesc(128, kHkHalf);
        const uint32_t sbo = 8u * (uint32_t) P.rowBytes, lay = (uint32_t) P.layout;
        constexpr int cLo = CLO, cHi = CHI;                          // compile-time: the issue loop is straight-line code
        for (int i = 0; i < myTiles; ++i) {
            const int st = i % kHkStages;
            const int b = TM::nBuf == 2 ? (i & 1) : 0, u = TM::nBuf == 2 ? (i >> 1) : i;             // accumulator set of this tile, its use count
            if (u >= 1) mbar_wait(accEmpty + 2 * h + b, (uint32_t) ((u - 1) & 1));                     // the epilogue has drained this set
            mbar_wait(bFull + st, (uint32_t) ((i / kHkStages) & 1));
            tc_fence_after();
            if (el) {
                const uint32_t hb = smem_u32(ring + (size_t) st * 2 * bufBytes) + (uint32_t) (h * kHkHalf * P.rowBytes), tb = hb + bufBytes;
                const uint64_t bH0 = make_desc(hb, 16, sbo, lay), bT0 = make_desc(tb, 16, sbo, lay);
                const uint32_t d0 = tmem + (uint32_t) TM::set(h, b), d1 = d0 + kHkHalf;
                if (!(dbg & 1)) {
                    #pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        #pragma unroll
                        for (int s = 0; s < KS; ++s) {
                            if ((s >= cLo && s <= cHi) != (pass == 1)) continue;
                            const uint64_t bH = bH0 + (uint64_t) (2 * s), bT = bT0 + (uint64_t) (2 * s);     // + 32 bytes per K step: the Hankel shift
                            const uint32_t first = (pass == 0 && s == 0) ? 0u : 1u;                          // step 0 is a tail step (cLo >= 1)
                            umma_ts(d0, tmem + (uint32_t) (8 * s), bH, idesc, first);                        // w0 x0
                            umma_ts(d1, tmem + (uint32_t) (TM::aTail + 8 * s), bH, idesc, first);            // w1 x0
                            umma_ts(d1, tmem + (uint32_t) (8 * s), bT, idesc, 1u);                           // w0 x1
                        }
                    }
                }
                umma_commit(bEmpty + st);                            // the stage may be refilled once both halves' MMAs have read it
                umma_commit(accFull + 2 * h + b);
            }
            __syncwarp();
        }
    } else {
        // =========================================================== epilogue
        const uint32_t laneBase = (uint32_t) (warp * 32) << 16;
        const int l = warp * 32 + lane;                              // TMEM lane = offset of the output inside its column
        const HankelTileRec* rec = recs + blockIdx.x;
        struct TileOut { float* out; long long oBase, numOut; };
        auto load_rec = [&](const HankelTileRec* r) { TileOut T; T.out = ldg_ptr(&r->out); T.oBase = __ldg(&r->oBase); T.numOut = __ldg(&r->numOut); return T; };
        TileOut N = {nullptr, 0, 0};
        if (myTiles > 0) N = load_rec(rec);
        for (int i = 0; i < myTiles; ++i, rec += gridDim.x) {
            const TileOut T = N;
            if (i + 1 < myTiles) N = load_rec(rec + gridDim.x);
            const bool inside = T.oBase >= 0 && T.oBase + 128LL * kHkCols <= T.numOut;
            float* outG = reinterpret_cast<float*>(__cvta_generic_to_global(T.out));
            const int b = TM::nBuf == 2 ? (i & 1) : 0, u = TM::nBuf == 2 ? (i >> 1) : i;
            #pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                mbar_wait(accFull + 2 * h + b, (uint32_t) (u & 1));
                tc_fence_after();
                #pragma unroll 1
                for (int c = 0; c < kHkHalf / 16; ++c) {
                    uint32_t v0[16], v1[16];
                    const uint32_t c0 = tmem + laneBase + (uint32_t) (TM::set(h, b) + c * 16);
                    #define LD16(arr, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                        : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), \
                          "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), "=r"(arr[15]) : "r"(addr))
                    LD16(v0, c0);
                    LD16(v1, c0 + (uint32_t) kHkHalf);
                    #undef LD16
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c == kHkHalf / 16 - 1) {                     // everything has been read: the set may be overwritten
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(accEmpty + 2 * h + b);
                    }
                    const long long o0 = T.oBase + 128LL * (h * kHkHalf + c * 16) + l;
                    if (dbg & 4) {} else if (inside) {
                        float* __restrict__ dst = outG + o0;
                        #pragma unroll
                        for (int j = 0; j < 16; ++j) __stcs(dst + 128 * j, combine(v0[j], v1[j]));
                    } else {
                        #pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const long long o = o0 + 128 * j;
                            if (o >= 0 && o < T.numOut) outG[o] = combine(v0[j], v1[j]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// fp32 recomputation of a launch whose input did not fit the fp16 split: exits at once unless the flag is set.
__global__ void __launch_bounds__(256)
hankel_redo_kernel(const HankelTileRec* __restrict__ recs, int nTiles, int L, PolyDev W, const unsigned* __restrict__ ovf) {
    if (*ovf == 0u) return;
    const int R = 128 / L;
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const HankelTileRec T = recs[t];
        for (int i = threadIdx.x; i < 128 * kHkCols; i += blockDim.x) {
            const long long o = T.oBase + i;
            if (o < 0 || o >= T.numOut) continue;
            const int col = i >> 7, l = i & 127, ii = l / L, k = l - ii * L;
            const long long m = T.x0 + 208 + (long long) R * col + ii;          // window index of the newest input of this output
            float acc = 0.0f;
            for (int j = 0; j < W.taps; ++j) {
                const long long li = m - (W.taps - 1) + j;
                const float x = (li >= 0 && li < T.inAvail) ? __ldg(T.in + li) : 0.0f;
                acc = fmaf(x, __ldg(W.W + (size_t) j * W.qpad + k), acc);
            }
            T.out[o] = acc;
        }
    }
}

}  // namespace

size_t hankel_smem_bytes(const HankelDev& P) {
    return 1024 + (size_t) kHkStages * 2 * P.bufBytes + (size_t) kHkRaw * (((size_t) P.elems * 4 + 127) & ~(size_t) 127) + 256;
}
long long hankel_tiles_for_segment(long long n0, long long numOut) {
    if (numOut <= 0) return 0;
    const long long c0 = n0 / 128, c1 = (n0 + numOut - 1) / 128;
    return (c1 - c0 + 1 + kHkCols - 1) / kHkCols;
}
int hankel_tile_elems(int L, int KS) { return (128 / L) * kHkCols + 16 * KS; }

cudaError_t launch_hankel(const ResampleLaunch& L, cudaStream_t s, long long* launches) {
    if (!L.d_tile_recs || !L.d_ovf) return cudaErrorInvalidValue;
    HankelTileRec* recs = reinterpret_cast<HankelTileRec*>(L.d_tile_recs);
    const size_t smem = hankel_smem_bytes(L.hk);
    cudaError_t e;
    if ((e = cudaMemsetAsync(L.d_ovf, 0, sizeof(unsigned), s)) != cudaSuccess) return e;
    if (!L.recs_ready) {
        hankel_tile_table_kernel<<<(L.n_tiles + 255) / 256, 256, 0, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, 128 / L.hk.L, recs);
        ++*launches;
    }
    const int grid = std::min(L.n_tiles, L.sm_count);
#ifdef F9_DIAG
    const int dbg = getenv("F9_HK_DBG") ? atoi(getenv("F9_HK_DBG")) : 0;      // -DF9_DIAG builds: 1 skip MMAs, 2 skip loads + conversion, 4 skip stores
#else
    const int dbg = 0;
#endif
    #define F9_HK_LAUNCH(ks_, lo_, hi_) do { \
        if (L.hk.KS != ks_ || L.hk.cLo != lo_ || L.hk.cHi != hi_) return cudaErrorInvalidValue; \
        if ((e = cudaFuncSetAttribute(hankel_fir_kernel<ks_, lo_, hi_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return e; \
        hankel_fir_kernel<ks_, lo_, hi_><<<grid, kHkThreads, smem, s>>>(recs, L.n_tiles, L.hk, L.d_ovf, dbg); } while (0)
    switch (L.hk.L) {                                   // (K steps, first / last main-lobe step) of the 200-tap kinds
        case 2:  F9_HK_LAUNCH(17, 6, 10); break;
        case 4:  F9_HK_LAUNCH(15, 6, 8); break;
        case 8:  F9_HK_LAUNCH(14, 6, 7); break;
        case 16: F9_HK_LAUNCH(14, 6, 7); break;
        default: return cudaErrorInvalidValue;
    }
    #undef F9_HK_LAUNCH
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    hankel_redo_kernel<<<std::min(L.n_tiles, 8 * L.sm_count), 256, 0, s>>>(recs, L.n_tiles, L.hk.L, L.poly, L.d_ovf);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace f9
