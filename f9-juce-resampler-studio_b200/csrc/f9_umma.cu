// Tensor-core polyphase FIR for rational ratios (sm_100a: tcgen05.mma with TMEM accumulators).
//
// JUCE's WindowedSinc interpolator is 200 taps per output: 400 FLOP against 5-13 bytes, far above the FP32 ridge, so
// the CUDA-core kernels in f9_resample.cu stop at ~20 % of the FP32 peak and ~7 % of the HBM roofline.  This kernel
// turns the same arithmetic into small GEMMs so that the stage becomes memory-bound:
//
//   * output n = a*q + k (period a, slot k) reads the `taps` inputs ending at a*p + floor(k*p/q) with weights that
//     depend on k only.  For one group of 16 adjacent slots the windows of a period overlap almost completely, so
//         D[a, k] = sum_t  X[a, t] * C_g[t, k],     X[a, t] = x[a*p + U0 + t]
//     is a GEMM with M = periods, N = 16 slots, K = the group's window (taps + 16*p/q + alignment, in steps of 16).
//   * a tile is 128 consecutive periods (M = 128) x one block of <= 14 groups.  K is walked once per tile from the
//     outside: every 16-sample step of X is staged once and used by all groups whose window contains it.
//   * precision: x = x0 + x1/2048, w = w0 + w1/2048 with fp16 parts (x1, w1 stored pre-multiplied by 2048 so they
//     stay normal numbers).  D0 += x0*w0 and D1 += x0*w1 + x1*w0 accumulate in fp32 in TMEM; the dropped x1*w1 term
//     is < 2^-24 relative.  out = D0 + D1/2048.  Samples with |x| >= 2^15 (or NaN/Inf) do not fit the split: the
//     loader raises a flag and umma_redo_kernel recomputes the launch in fp32.
//
// Warp roles (416 threads, one CTA per SM, persistent over tiles):
//   warps 0-3   epilogue: tcgen05.ld the finished accumulators (lane = period), combine D0/D1, transpose through
//               shared memory, coalesced stores
//   warp  4     tensor pipe: tcgen05.cp the staged X steps into TMEM (the A operand is read from TMEM, so an MMA
//               costs N/2 cycles instead of the ~36 an SS-mode MMA spends fetching 4 KB of A), then the MMAs of
//               the host-built schedule; tcgen05.commit signals "stage free" / "group done"
//   warps 5-12  loaders: global -> fp16 split -> shared memory in the canonical K-major (no swizzle) operand layout
// Measured building blocks (tools/ubench/umma_probe.cu, B200): SS MMA M=128 = 32 + N/4 clk, TS MMA = N/2 clk,
// tcgen05.cp 128x256b = 64 clk.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>

#include "f9_internal.cuh"

namespace f9 {
namespace {

constexpr int kRows = 128;                       // periods per tile = MMA M
constexpr int kChunk = kRows * 16 + 32;          // bytes between K chunks (8 samples) of the X operand: 128 rows x 16 B + pad
constexpr int kStageBytes = 8 * kChunk;          // one stage = 32 samples: chunks 0-3 head (x0), 4-7 tail (x1)
constexpr int kEpiPitch = 36;                    // floats per row of the epilogue transpose buffer (32 + 4)
constexpr int kACol = 448;                       // TMEM columns 448..511: ring of 4 K steps x (8 head + 8 tail)
constexpr int kLoaderWarps = 8;
constexpr int kThreads = (4 + 1 + kLoaderWarps) * 32;
constexpr int kSpin = 1 << 26;                   // bounded waits: a protocol bug must not hang the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < kSpin; ++i) if (mbar_try_wait(bar, parity)) return;
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
__device__ __forceinline__ void umma_cp(uint32_t d_tmem, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(d_tmem), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no swizzle operand: element (row r, k) at (k/8)*lbo + (r/8)*sbo + (r%8)*16 + (k%8)*2 bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t) ((saddr >> 4) & 0x3fff) | ((uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t) 1 << 46);
}
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {        // fp16 x fp16 -> fp32, both K-major
    return (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

__device__ __forceinline__ int find_seg(const int* __restrict__ prefix, int n, int bid) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= bid) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ float load_in(const Seg& S, long long l) {          // l: index into the segment's window
    return (l >= 0 && l < S.inAvail) ? __ldg(S.in + l) : 0.0f;
}

struct SmemMap {
    uint8_t* W; uint8_t* ring; float* epi; uint16_t* sched; uint8_t* ksCount;
    uint64_t *full, *empty, *accFull, *accEmpty; uint32_t* tmemSlot;
};
__device__ __forceinline__ SmemMap carve(uint8_t* smem, int maxEntries, int maxNK, int stages) {
    SmemMap m;
    m.W = smem;
    m.ring = m.W + (size_t) maxEntries * 1024;
    m.epi = reinterpret_cast<float*>(m.ring + (size_t) stages * kStageBytes);
    m.sched = reinterpret_cast<uint16_t*>(m.epi + kRows * kEpiPitch);
    m.ksCount = reinterpret_cast<uint8_t*>(m.sched) + ((size_t) maxEntries * 2 + 15) / 16 * 16;
    m.full = reinterpret_cast<uint64_t*>(m.ksCount + ((size_t) maxNK + 15) / 16 * 16);
    m.empty = m.full + stages;
    m.accFull = m.empty + stages;
    m.accEmpty = m.accFull + kUmmaMaxGroups;
    m.tmemSlot = reinterpret_cast<uint32_t*>(m.accEmpty + kUmmaMaxGroups);
    return m;
}

__global__ void __launch_bounds__(kThreads, 1)
umma_fir_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles, const UmmaDev P,
                int stages, unsigned* __restrict__ ovf) {
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemMap sm = carve(smem, P.maxEntries, P.maxNK, stages);
    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int gb = blockIdx.x % P.nGB;
    const UmmaBlockInfo BI = P.blk[gb];
    const int p = P.p, q = P.q;
    const int nStages = BI.nStages;                            // stages per tile (two K steps each)

    // ---- one-time setup: weights + schedule into shared memory, barriers, TMEM
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.W + BI.wOff);
        uint4* dst = reinterpret_cast<uint4*>(sm.W);
        for (int i = threadIdx.x; i < BI.nEntries * 64; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < BI.nEntries; i += kThreads) sm.sched[i] = P.sched[BI.entryOff + i];
        for (int i = threadIdx.x; i < BI.nK; i += kThreads) sm.ksCount[i] = P.ksCount[BI.ksOff + i];
        if (threadIdx.x == 0) {
            for (int s = 0; s < stages; ++s) { mbar_init(sm.full + s, kLoaderWarps); mbar_init(sm.empty + s, 1); }
            for (int g = 0; g < kUmmaMaxGroups; ++g) { mbar_init(sm.accFull + g, 1); mbar_init(sm.accEmpty + g, 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 4) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sm.tmemSlot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_async_smem();                                    // the weights are read by the tensor pipe (async proxy)
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem = *sm.tmemSlot;

    if (warp >= 5) {
        // =========================================================== loaders
        // Every lane owns, per stage, 4 row pieces of 4 samples (rows i*32 + lw*4 + lane/8, samples 4*(lane%8)..+3): one
        // warp instruction reads 4 rows x 128 contiguous bytes.  Loads run two stages ahead of the conversion (three
        // register buffers) so ~32 KB per SM are in flight; the cursor walks (tile, stage) across tile boundaries.
        const int lw = warp - 5;
        const int rsub = lane >> 3, j = lane & 7;
        const uint32_t dstLane = (uint32_t) ((j >> 1) * kChunk + (j & 1) * 8);
        uint32_t amax = 0;                                     // running max of |x| bit patterns
        const int myTiles = blockIdx.x < nTiles ? (nTiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;
        const int total = myTiles * nStages;

        int curTile = blockIdx.x, curSt = 0;                   // prefetch cursor
        Seg S = {}; long long l00 = 0, addr0 = 0; bool aligned = false;
        auto open_tile = [&]() {
            const int sidx = find_seg(tilePrefix, nSegs, curTile);
            S = segs[sidx];
            const int pb = (curTile - tilePrefix[sidx]) / P.nGB;
            const long long A0 = S.n0 / q + (long long) pb * kRows;
            l00 = A0 * p + BI.U0 - S.inOffset;                                  // window index of (row 0, K 0)
            addr0 = (long long) (reinterpret_cast<uintptr_t>(S.in) >> 2) + l00; // its address in floats
            aligned = ((p & 3) == 0) && ((addr0 & 3) == 0);                     // every row piece starts on 16 bytes
        };
        // A buffer holds, per row piece, the 16-byte ALIGNED vector that starts m = (address mod 4) floats before the
        // lane's 4 samples; meta = m of the 4 rows (2 bits each) | bit 8: tile fully aligned (all m = 0).
        struct Buf { float4 v[4]; uint32_t meta; };
        auto issue = [&](Buf& b) {
            if (curSt == 0) open_tile();
            uint32_t meta = aligned ? 0x100u : 0u;
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = i * 32 + lw * 4 + rsub;
                const int m = (int) ((addr0 + (long long) r * p) & 3);
                meta |= (uint32_t) m << (2 * i);
                const long long l = l00 + (long long) r * p + curSt * 32 + j * 4 - m;
                if (l >= 0 && l + 3 < S.inAvail) b.v[i] = __ldg(reinterpret_cast<const float4*>(S.in + l));
                else { b.v[i].x = load_in(S, l); b.v[i].y = load_in(S, l + 1); b.v[i].z = load_in(S, l + 2); b.v[i].w = load_in(S, l + 3); }
            }
            b.meta = meta;
            if (++curSt == nStages) { curSt = 0; curTile += gridDim.x; }
        };
        uint32_t it = 0;                                       // stages produced by this CTA so far
        const int srcLane = (lane & ~7) | ((j + 1) & 7);       // the lane holding the next 16 bytes of this row
        auto process = [&](const Buf& b, const Buf& nb) {
            const int s = (int) (it % (uint32_t) stages);
            const uint32_t ph = (it / (uint32_t) stages) & 1;
            ++it;
            float4 x[4];
            if (b.meta & 0x100u) {
                #pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = b.v[i];
            } else {
                // funnel shift: samples m..m+3 of (own vector, next vector); lane 7's next vector is lane 0's vector of the
                // next stage (already in registers: the loads run two stages ahead)
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 pub = (j == 0) ? nb.v[i] : b.v[i];
                    const float c4 = __shfl_sync(0xffffffffu, pub.x, srcLane);
                    const float c5 = __shfl_sync(0xffffffffu, pub.y, srcLane);
                    const float c6 = __shfl_sync(0xffffffffu, pub.z, srcLane);
                    const uint32_t m = (b.meta >> (2 * i)) & 3u;
                    const float4 v = b.v[i];
                    const bool m1 = m & 1u, m2 = m & 2u;
                    const float e0 = m1 ? v.y : v.x, e1 = m1 ? v.z : v.y, e2 = m1 ? v.w : v.z, e3 = m1 ? c4 : v.w, e4 = m1 ? c5 : c4, e5 = m1 ? c6 : c5;
                    x[i] = make_float4(m2 ? e2 : e0, m2 ? e3 : e1, m2 ? e4 : e2, m2 ? e5 : e3);
                }
            }
            mbar_wait(sm.empty + s, ph ^ 1);
            uint8_t* stage = sm.ring + (size_t) s * kStageBytes;
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = i * 32 + lw * 4 + rsub;
                const float4 xv = x[i];
                amax = max(max(amax, __float_as_uint(xv.x) & 0x7fffffffu), max(__float_as_uint(xv.y) & 0x7fffffffu,
                           max(__float_as_uint(xv.z) & 0x7fffffffu, __float_as_uint(xv.w) & 0x7fffffffu)));
                const __half2 h01 = __floats2half2_rn(xv.x, xv.y), h23 = __floats2half2_rn(xv.z, xv.w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 t01 = __floats2half2_rn((xv.x - f01.x) * 2048.0f, (xv.y - f01.y) * 2048.0f);
                const __half2 t23 = __floats2half2_rn((xv.z - f23.x) * 2048.0f, (xv.w - f23.y) * 2048.0f);
                uint8_t* dst = stage + dstLane + r * 16;
                uint2 hv, tv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                tv.x = *reinterpret_cast<const uint32_t*>(&t01); tv.y = *reinterpret_cast<const uint32_t*>(&t23);
                *reinterpret_cast<uint2*>(dst) = hv;
                *reinterpret_cast<uint2*>(dst + 4 * kChunk) = tv;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.full + s);
        };
        Buf b0, b1, b2;
        #pragma unroll
        for (int i = 0; i < 4; ++i) b0.v[i] = b1.v[i] = b2.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        b0.meta = b1.meta = b2.meta = 0x100u;
        int issued = 0, done = 0;
        if (issued < total) { issue(b0); ++issued; }
        if (issued < total) { issue(b1); ++issued; }
        while (done < total) {
            if (issued < total) { issue(b2); ++issued; }
            process(b0, b1); if (++done >= total) break;
            if (issued < total) { issue(b0); ++issued; }
            process(b1, b2); if (++done >= total) break;
            if (issued < total) { issue(b1); ++issued; }
            process(b2, b0); ++done;
        }
        if (amax >= 0x47000000u) atomicOr(ovf, 1u);            // |x| >= 32768, Inf or NaN somewhere in this CTA's input
    } else if (warp == 4) {
        // =========================================================== tensor pipe
        const uint32_t el = elect_one();
        const uint32_t idescHi = make_idesc(kRows, 32), idescLo = make_idesc(kRows, 16);
        const uint32_t wBase = smem_u32(sm.W);
        uint32_t it = 0, tcount = 0;
        for (int tileId = blockIdx.x; tileId < nTiles; tileId += gridDim.x, ++tcount) {
            int e = 0;
            for (int st = 0; st < nStages; ++st, ++it) {
                const int s = (int) (it % (uint32_t) stages);
                const uint32_t ph = (it / (uint32_t) stages) & 1;
                mbar_wait(sm.full + s, ph);
                tc_fence_after();
                const uint32_t stageAddr = smem_u32(sm.ring + (size_t) s * kStageBytes);
                if (el) {
                    #pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t slot = (uint32_t) (((2 * st + h) & 3) * 16);
                        umma_cp(tmem + kACol + slot, make_desc(stageAddr + h * 2 * kChunk, kChunk, 128));
                        umma_cp(tmem + kACol + slot + 8, make_desc(stageAddr + (4 + h * 2) * kChunk, kChunk, 128));
                    }
                    umma_commit(sm.empty + s);                 // the stage is free once the copies have read it
                }
                __syncwarp();
                for (int h = 0; h < 2; ++h) {
                    const int ks = 2 * st + h;
                    if (ks >= BI.nK) break;
                    const uint32_t aHi = tmem + kACol + (uint32_t) ((ks & 3) * 16);
                    const int cnt = sm.ksCount[ks];
                    for (int c = 0; c < cnt; ++c, ++e) {
                        const uint32_t ent = sm.sched[e];
                        const uint32_t gl = ent & 63u;
                        if (ent & 0x40u) {                     // first K step of this group in this tile: accumulator drained?
                            mbar_wait(sm.accEmpty + gl, (tcount & 1) ^ 1);
                            tc_fence_after();
                        }
                        if (el) {
                            const uint64_t bd = make_desc(wBase + (uint32_t) e * 1024u, 512, 128);
                            umma_ts(tmem + gl * 32, aHi, bd, idescHi, (ent & 0x40u) ? 0u : 1u);   // [D0 | D1] (+)= x0 * [w0 | w1]
                            umma_ts(tmem + gl * 32 + 16, aHi + 8, bd, idescLo, 1u);               //  D1       +=  x1 * w0
                            if (ent & 0x80u) umma_commit(sm.accFull + gl);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // =========================================================== epilogue
        const int row0 = warp * 32 + lane;                     // TMEM lane = period row of this thread
        uint32_t tcount = 0;
        for (int tileId = blockIdx.x; tileId < nTiles; tileId += gridDim.x, ++tcount) {
            const int sidx = find_seg(tilePrefix, nSegs, tileId);
            const Seg S = segs[sidx];
            const int pb = (tileId - tilePrefix[sidx]) / P.nGB;
            const long long A0 = S.n0 / q + (long long) pb * kRows;
            for (int gp = 0; gp * 2 < BI.nGroups; ++gp) {
                #pragma unroll
                for (int gg = 0; gg < 2; ++gg) {
                    const int gl = gp * 2 + gg;
                    float o[16];
                    if (gl < BI.nGroups) {
                        mbar_wait(sm.accFull + gl, tcount & 1);
                        tc_fence_after();
                        uint32_t v[32];
                        const uint32_t taddr = tmem + (uint32_t) (gl * 32) + ((uint32_t) (warp * 32) << 16);
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                     : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(sm.accEmpty + gl);      // the next tile may overwrite this accumulator
                        #pragma unroll
                        for (int c = 0; c < 16; ++c) o[c] = fmaf(__uint_as_float(v[16 + c]), 1.0f / 2048.0f, __uint_as_float(v[c]));
                    } else {
                        #pragma unroll
                        for (int c = 0; c < 16; ++c) o[c] = 0.0f;
                    }
                    float4* dst = reinterpret_cast<float4*>(sm.epi + row0 * kEpiPitch + gg * 16);
                    dst[0] = make_float4(o[0], o[1], o[2], o[3]);   dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                    dst[2] = make_float4(o[8], o[9], o[10], o[11]); dst[3] = make_float4(o[12], o[13], o[14], o[15]);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // rows of 32 slots are contiguous in the output: one coalesced 128-byte store per row
                const int slot = BI.slot0 + gp * 32 + lane;
                for (int r = warp; r < kRows; r += 4) {
                    const long long o = (A0 + r) * q + slot - S.n0;
                    if (slot < q && o >= 0 && o < S.numOut) S.out[o] = sm.epi[r * kEpiPitch + lane];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// fp32 recomputation of a launch whose input did not fit the fp16 split (same tiles, CUDA cores, no staging):
// exits at once unless the flag is set.
__global__ void __launch_bounds__(256)
umma_redo_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles, int nGB, int q16,
                 PolyDev W, const unsigned* __restrict__ ovf) {
    if (*ovf == 0u) return;
    for (int tileId = blockIdx.x; tileId < nTiles; tileId += gridDim.x) {
        if (tileId % nGB != 0) continue;                       // one pass per period block covers all slots
        const int sidx = find_seg(tilePrefix, nSegs, tileId);
        const Seg S = segs[sidx];
        const int pb = (tileId - tilePrefix[sidx]) / nGB;
        const long long nBase = (S.n0 / q16 + (long long) pb * kRows) * q16;          // first output of the tile (absolute)
        const int cnt = kRows * q16;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const long long n = nBase + i, o = n - S.n0;
            if (o < 0 || o >= S.numOut) continue;
            const long long a = n / W.q; const int k = (int) (n - a * W.q);
            const long long m = a * W.p + __ldg(W.B + k) - S.inOffset;
            float acc = 0.0f;
            for (int t = 0; t < W.taps; ++t) acc = fmaf(load_in(S, m - (W.taps - 1) + t), __ldg(W.W + (size_t) t * W.qpad + k), acc);
            S.out[o] = acc;
        }
    }
}

}  // namespace

cudaError_t launch_umma(const ResampleLaunch& L, cudaStream_t s, long long* launches) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(umma_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    int grid = std::min(L.n_tiles, std::max(L.sm_count, L.um.nGB));
    grid -= grid % L.um.nGB;
    if (grid <= 0) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(L.d_ovf, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    umma_fir_kernel<<<grid, kThreads, L.um_smem, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.um, L.um_stages, L.d_ovf);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    umma_redo_kernel<<<std::min(L.n_tiles, 8 * L.sm_count), 256, 0, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.um.nGB, L.um.q,
                                                                          L.poly, L.d_ovf);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    return cudaSuccess;
}

}  // namespace f9
