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
//     is < 2^-24 relative.  out = D0 + D1/2048.  Samples are pre-scaled by 2^7 (see split_store); |x| >= 256 (or NaN/Inf)
//     does not fit the split: the loader raises a flag and umma_redo_kernel recomputes the launch in fp32.
//
// Variants (template <MERGED, TMA, CTA2>):
//   * TMA-fed (rows 16-byte aligned): warp 4 streams boxes of 128 overlapping rows x 32 floats (tensor map with a row stride of
//     p floats) into a ring of raw fp32 stages, warps 10-17 convert (thread = row) and write the TMEM operand with tcgen05.st,
//     warps 5-9 issue the MMAs from host-built lists, warps 0-3 drain the accumulators.
//   * CTA pairs on top of that (tcgen05.mma.cta_group::2): half the weights per CTA, the leader issues for both.
//   * register loader (any alignment): warps 8-15 load, split and stage canonical K-major tiles in shared memory, warp 4 copies
//     them into TMEM with tcgen05.cp, warps 5-7 issue.
// Measured building blocks (tools/ubench/umma_probe.cu, tma_probe.cu, B200): SS MMA M=128 = 32 + N/4 clk, TS MMA = N/2 clk,
// tcgen05.cp 128x256b = 64 clk, box stream 2.7 / 4.6 TB/s of unique input with a ring of 2 / 8 stages.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <cuda_fp16.h>

#include "f9_internal.cuh"

namespace f9 {
namespace {

constexpr int kRows = 128;                       // periods per tile = MMA M
constexpr int kChunk = kRows * 16 + 32;          // bytes between K chunks (8 samples) of the X operand: 128 rows x 16 B + pad
constexpr int kStageBytes = 8 * kChunk;          // one stage = 32 samples: chunks 0-3 head (x0), 4-7 tail (x1)
constexpr int kEpiPitch = 20;                    // floats per row of the epilogue transpose buffer (16 + 4)
constexpr int kASlotsMax = 4;                    // TMEM operand ring: up to 4 slots of one stage (2 K steps x (8 head + 8 tail) columns)
                                                 // in the top columns (UmmaDev::aSlots: 2 or 4, chosen with the plan)
constexpr int kLoaderWarps = 8;
constexpr int kIssuers = kUmmaIssuers;            // MMA-issuing warps (groups are dealt round-robin)
constexpr int kFirstLoader = 4 + 1 + kIssuers;   // warps 0-3 epilogue, 4 copy, 5..7 issue, 8..15 load
constexpr int kThreads = (kFirstLoader + kLoaderWarps) * 32;
// TMA-fed kernels: warps 0-3 epilogue, 4 producer, 5..9 issue, 10.. convert (a quarter of the rows x (part of) one K step each)
constexpr int kIssuersTma = kUmmaIssuersTma;
constexpr int kFirstConv = 4 + 1 + kIssuersTma;
constexpr int kConvTeams = 2;                    // teams of four warps (one per TMEM lane quarter); team i takes the stages gs % kConvTeams == i
constexpr int kConvWarps = 4 * kConvTeams;
constexpr int kConvPerStage = 4;                 // warps that take part in one stage: its barriers count these
constexpr int kThreadsTma = (kFirstConv + kConvWarps) * 32;
constexpr uint32_t kSleepProducerNs = 100, kSleepEpilogueNs = 200;   // mbar_wait_sleep: roles with slack
constexpr uint32_t kParkNs = 1000;                // suspend-time hint of the TMA roles' barrier waits (a hot poll loop cost 40 % of the issue slots)
#ifdef F9_DIAG
constexpr int kSpin = 1 << 22;
#else
constexpr int kSpin = 1 << 26;
#endif
//                   // bounded waits: a protocol bug must not hang the GPU
// -DF9_DIAG builds: event trace of the TMA-fed kernel's roles (F9_UMMA_TRACE=k traces launch k): per warp of CTAs 0 and 1 a list of
// (event, index, clock) for the tiles kTrTile0 .. kTrTile1 - 1 of the CTA, dumped by launch_umma (tools/umma_trace.py reads it).
constexpr int kTrCap = 1024, kTrWarps = 18;
#ifndef F9_TR_TILE0
#define F9_TR_TILE0 6
#endif
constexpr int kTrTile0 = F9_TR_TILE0, kTrTile1 = F9_TR_TILE0 + 4;
#ifdef F9_DIAG
#define TR_INIT(prof, warp) long long* trBuf = ((prof) && blockIdx.x < 2) ? (prof) + ((size_t) blockIdx.x * kTrWarps + (warp)) * kTrCap : nullptr; int trN = 0
#define TR_EV(t, ev, idx) do { if (trBuf && (t) >= kTrTile0 && (t) < kTrTile1 && trN < kTrCap) \
    trBuf[trN++] = ((long long) (ev) << 56) | ((long long) ((idx) & 0xffff) << 40) | (clock64() & 0xffffffffffll); } while (0)
#else
#define TR_INIT(prof, warp) do {} while (0)
#define TR_EV(t, ev, idx) do {} while (0)
#endif

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
#ifdef F9_DIAG
#define F9_TRAP(bar, parity) do { printf("[umma] wait timed out: block %d warp %d barrier +%u parity %u\n", (int) blockIdx.x, (int) (threadIdx.x >> 5), smem_u32(bar), (unsigned) (parity)); __trap(); } while (0)
#else
#define F9_TRAP(bar, parity) __trap()
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    #pragma unroll 1
    for (int i = 0; i < kSpin; ++i) if (mbar_try_wait(bar, parity)) return;
    F9_TRAP(bar, parity);
}
// Same, for waits that are expected to be long (epilogue): the hardware may park the thread for up to `ns` per attempt
// instead of re-issuing the poll, which leaves the issue slots to the loader warps.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t ns) {
    #pragma unroll 1
    for (int i = 0; i < kSpin; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
        if (ok) return;
    }
    F9_TRAP(bar, parity);
}
// For the roles that are NOT on the pair's critical path (producer: a ring of boxes ahead; epilogue: a group every few stages):
// poll, then really sleep.  The suspend hint of try_wait compiles to NANOSLEEP.SYNCS, which returns on any barrier traffic of the SM
// (~23 clk per poll measured: 26 polls per wait), so a dozen waiting warps took a third of the issue slots from the converters.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
    #pragma unroll 1
    for (int i = 0; i < kSpin; ++i) {
        if (mbar_try_wait(bar, parity)) return;
        asm volatile("nanosleep.u32 %0;" :: "r"(ns));
    }
    F9_TRAP(bar, parity);
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void umma_cp(uint32_t d_tmem, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(d_tmem), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {       // p and bytes multiples of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
// ---- CTA pairs (cluster of two): barriers that collect arrivals from both CTAs live in the leader (cluster rank 0)
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// What crosses between the CTAs is TMEM traffic ordered by the tcgen05 fences, not generic memory: the arrive keeps the default
// (CTA-scope) release.  A cluster-scope release compiled to MEMBAR.ALL.GPU + ERRBAR in front of every arrive: 40 % of all stalls.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {                 // same offset in the leader's shared memory
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t ns) {     // acquires the peer CTA's arrivals too
    #pragma unroll 1
    for (int i = 0; i < kSpin; ++i) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
        if (ok) return;
    }
    F9_TRAP(bar, parity);
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {                  // arrives on the barrier at this offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t) 3) : "memory");
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
    uint8_t* W; uint8_t* ring; float* epi; uint4* ops;
    uint64_t *full, *empty, *accFull, *accEmpty, *cpDone, *slotFree; uint32_t* tmemSlot;
};
constexpr int kTmaStageBytes = kRows * 128;         // TMA feed: one stage = a box of 128 rows x 32 floats, 128-byte swizzle
__device__ __forceinline__ SmemMap carve(uint8_t* smem, int maxEntries, int NB, int stages, bool tma, bool cta2) {
    SmemMap m;
    m.W = smem;
    m.ring = m.W + (size_t) maxEntries * NB * (cta2 ? 32 : 64);
    if (tma) m.ring += (1024u - (smem_u32(m.ring) & 1023u)) & 1023u;          // swizzle atoms are 1024 bytes
    m.epi = reinterpret_cast<float*>(m.ring + (size_t) stages * (tma ? kTmaStageBytes : kStageBytes));
    m.ops = reinterpret_cast<uint4*>(m.epi + kRows * kEpiPitch);
    m.full = reinterpret_cast<uint64_t*>(m.ops + maxEntries);
    m.empty = m.full + stages;
    m.accFull = m.empty + stages;
    m.accEmpty = m.accFull + kUmmaMaxGroups;
    m.cpDone = m.accEmpty + kUmmaMaxGroups;
    m.slotFree = m.cpDone + kASlotsMax;
    m.tmemSlot = reinterpret_cast<uint32_t*>(m.slotFree + kASlotsMax);
    return m;
}

// fp16 head / scaled fp16 tail of four samples, written to the operand tile (8 bytes each).  Samples are pre-scaled by
// kPreScale = 2^7 (exact) so that the fp16 head stays a normal number down to |x| = 2^-21 (-126 dBFS): without it signals
// below -84 dBFS would lose head bits to fp16 subnormals.  The price is the range: |x| >= 256 (+48 dBFS) takes the fp32 redo.
constexpr float kPreScale = 128.0f;
__device__ __forceinline__ void split_store(float4 xv, uint8_t* dst, __half2& hmax) {
    xv.x *= kPreScale; xv.y *= kPreScale; xv.z *= kPreScale; xv.w *= kPreScale;
    const __half2 h01 = __floats2half2_rn(xv.x, xv.y), h23 = __floats2half2_rn(xv.z, xv.w);
    hmax = __hmax2_nan(hmax, __hmax2_nan(__habs2(h01), __habs2(h23)));
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 t01 = __floats2half2_rn((xv.x - f01.x) * 2048.0f, (xv.y - f01.y) * 2048.0f);
    const __half2 t23 = __floats2half2_rn((xv.z - f23.x) * 2048.0f, (xv.w - f23.y) * 2048.0f);
    uint2 hv, tv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    tv.x = *reinterpret_cast<const uint32_t*>(&t01); tv.y = *reinterpret_cast<const uint32_t*>(&t23);
    *reinterpret_cast<uint2*>(dst) = hv;
    *reinterpret_cast<uint2*>(dst + 4 * kChunk) = tv;
}

// ------------------------------------------------------------------------------------------------- loader warps
// Every lane owns, per stage, 4 row pieces of 4 samples (rows i*32 + lw*4 + lane/8, samples 4*(lane%8)..+3): one warp
// instruction reads 4 rows x 128 contiguous bytes.  Loads run three stages ahead of the conversion (four register buffers,
// 48 KB per SM in flight); the cursor walks (tile, stage) across tile boundaries.
// ALIGNED: every row piece of every segment starts on a 16-byte boundary (the host checked), so the loaded vector is the
// lane's four samples.  Otherwise a lane loads the aligned vector that starts m = (address mod 4) floats early and the four
// samples are funnelled out of (own vector, next lane's vector); lane 7's "next" is lane 0's vector of the next stage.
struct LoaderArgs {
    const Seg* segs; const int* tilePrefix; int nSegs, nTiles, nGB, p, q, U0, nStages, stages, myTiles;
    uint8_t* ring; uint64_t *full, *empty; unsigned* ovf; long long* prof;
};
template <bool ALIGNED>
__device__ __forceinline__ void loader_role(const LoaderArgs& A, int lw, int lane) {
    const int rsub = lane >> 3, j = lane & 7;
    const uint32_t dstLane = (uint32_t) ((j >> 1) * kChunk + (j & 1) * 8 + (lw * 4 + rsub) * 16);
    __half2 hmax = __floats2half2_rn(0.f, 0.f);                // running max |x0| (NaN-propagating)
    const int total = A.myTiles * A.nStages;
    long long pT0 = 0, pW0 = 0;
    if (A.prof) pT0 = clock64();

    // ---- prefetch cursor: state of the tile being loaded (three stages ahead of the one being converted)
    int curTile = blockIdx.x, curSt = 0;
    const float* rowPtr[4] = {nullptr, nullptr, nullptr, nullptr};            // aligned vector of (row i, stage 0) for this lane
    long long lRow[4] = {0, 0, 0, 0}, lEnd = 0;                               // its window index; S.inAvail
    uint32_t mrow = 0; bool interior = false;
    auto open_tile = [&]() {
        const int sidx = find_seg(A.tilePrefix, A.nSegs, curTile);
        const Seg S = A.segs[sidx];
        const int pb = (curTile - A.tilePrefix[sidx]) / A.nGB;
        const long long A0 = S.n0 / A.q + (long long) pb * kRows;
        const long long l00 = A0 * A.p + A.U0 - S.inOffset;                     // window index of (row 0, K 0)
        const long long addr0 = (long long) (reinterpret_cast<uintptr_t>(S.in) >> 2) + l00;   // its address in floats
        uint32_t mm = 0;
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = i * 32 + lw * 4 + rsub;
            const int m = ALIGNED ? 0 : (int) ((addr0 + (long long) r * A.p) & 3);
            mm |= (uint32_t) m << (2 * i);
            lRow[i] = l00 + (long long) r * A.p + j * 4 - m;
            rowPtr[i] = S.in + lRow[i];
        }
        mrow = mm; lEnd = S.inAvail;
        // interior: every vector any lane loads for this tile lies inside the segment's window
        interior = l00 - 3 >= 0 && l00 + (long long) (kRows - 1) * A.p + (long long) A.nStages * 32 + 3 < S.inAvail;
    };
    struct Buf { float4 v[4]; uint32_t mrow; };
    auto issue = [&](Buf& b) {
        if (curSt == 0) open_tile();
        const int koff = curSt * 32;
        if (interior) {
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(b.v[i].x), "=f"(b.v[i].y), "=f"(b.v[i].z), "=f"(b.v[i].w) : "l"(rowPtr[i] + koff));
        } else {
            // Predicated loads straight into the buffer registers (no branch, no select between a load and its use).
            bool straddle = false;
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long l = lRow[i] + koff;
                const bool inside = l >= 0 && l + 3 < lEnd;
                straddle |= !inside && l + 3 >= 0 && l < lEnd;
                b.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                             : "+f"(b.v[i].x), "+f"(b.v[i].y), "+f"(b.v[i].z), "+f"(b.v[i].w) : "l"(rowPtr[i] + koff), "r"((int) inside));
            }
            if (__any_sync(0xffffffffu, straddle)) {           // a vector crosses the start or the end of the segment's window
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long l = lRow[i] + koff;
                    if (!(l >= 0 && l + 3 < lEnd)) {
                        const float* ptr = rowPtr[i] + koff;
                        b.v[i].x = (l >= 0 && l < lEnd) ? __ldg(ptr) : 0.f;             b.v[i].y = (l + 1 >= 0 && l + 1 < lEnd) ? __ldg(ptr + 1) : 0.f;
                        b.v[i].z = (l + 2 >= 0 && l + 2 < lEnd) ? __ldg(ptr + 2) : 0.f; b.v[i].w = (l + 3 >= 0 && l + 3 < lEnd) ? __ldg(ptr + 3) : 0.f;
                    }
                }
            }
        }
        b.mrow = mrow;
        if (++curSt == A.nStages) { curSt = 0; curTile += gridDim.x; }
    };
    int sIdx = 0; uint32_t sPh = 0;                            // ring slot / phase of the stage being produced
    const int srcLane = (lane & ~7) | ((j + 1) & 7);           // the lane holding the next 16 bytes of this row
    auto process = [&](const Buf& b, const Buf& nb) {
        float4 x[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (ALIGNED) { x[i] = b.v[i]; continue; }
            const float4 pub = (j == 0) ? nb.v[i] : b.v[i];
            const float4 v = b.v[i];
            const float c4 = __shfl_sync(0xffffffffu, pub.x, srcLane);
            const float c5 = __shfl_sync(0xffffffffu, pub.y, srcLane);
            const float c6 = __shfl_sync(0xffffffffu, pub.z, srcLane);
            const uint32_t m = (b.mrow >> (2 * i)) & 3u;
            const bool m1 = m & 1u, m2 = m & 2u;
            const float e0 = m1 ? v.y : v.x, e1 = m1 ? v.z : v.y, e2 = m1 ? v.w : v.z, e3 = m1 ? c4 : v.w, e4 = m1 ? c5 : c4, e5 = m1 ? c6 : c5;
            x[i] = make_float4(m2 ? e2 : e0, m2 ? e3 : e1, m2 ? e4 : e2, m2 ? e5 : e3);
        }
        if (A.prof) { const long long w = clock64(); mbar_wait(A.empty + sIdx, sPh ^ 1); pW0 += clock64() - w; }
        else mbar_wait(A.empty + sIdx, sPh ^ 1);
        uint8_t* dst = A.ring + (size_t) sIdx * kStageBytes + dstLane;
        #pragma unroll
        for (int i = 0; i < 4; ++i) split_store(x[i], dst + i * 32 * 16, hmax);
        // No proxy fence here: it compiles to MEMBAR.ALL.CTA, which would wait for the loads already in flight for the next
        // stages.  The arrive below releases the stores; the tensor warp fences after it has acquired the barrier.
        __syncwarp();
        if (lane == 0) mbar_arrive(A.full + sIdx);
        if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; }
    };
    Buf b0, b1, b2, b3;
    #pragma unroll
    for (int i = 0; i < 4; ++i) b0.v[i] = b1.v[i] = b2.v[i] = b3.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    b0.mrow = b1.mrow = b2.mrow = b3.mrow = 0;
    int issued = 0, done = 0;
    if (issued < total) { issue(b0); ++issued; }
    if (issued < total) { issue(b1); ++issued; }
    if (issued < total) { issue(b2); ++issued; }
    while (done < total) {
        if (issued < total) { issue(b3); ++issued; }
        process(b0, b1); if (++done >= total) break;
        if (issued < total) { issue(b0); ++issued; }
        process(b1, b2); if (++done >= total) break;
        if (issued < total) { issue(b1); ++issued; }
        process(b2, b3); if (++done >= total) break;
        if (issued < total) { issue(b2); ++issued; }
        process(b3, b0); ++done;
    }
    const float2 hm = __half22float2(hmax);
    if (!(hm.x < 32768.0f) || !(hm.y < 32768.0f)) atomicOr(A.ovf, 1u);         // |128 x| >= 32768, Inf or NaN in this CTA's input
    if (A.prof && lw == 0 && lane == 0) { A.prof[blockIdx.x * 16 + 0] = clock64() - pT0; A.prof[blockIdx.x * 16 + 1] = pW0; }
}


// ------------------------------------------------------------------------------------------------- TMA feed
// With 16-byte aligned rows (p % 4 == 0, aligned segments) the loader warps above are replaced by:
//   warp 4, one lane: per stage one cp.async.bulk.tensor of the box [128 rows (p floats apart: the rows overlap in memory) x
//                     32 floats] into a ring of raw fp32 stages (128-byte swizzle), plus a bulk L2 prefetch of the CTA's next tile;
//   warps 8-15:       thread = period row = TMEM lane.  Each warp converts one 16-sample K step of its 32 rows (4 conflict-free
//                     LDS.128, fp16 head / tail split in registers) and writes it straight into the TMEM operand slot with one
//                     tcgen05.st.32x32b.x16: no converted copy in shared memory, no tcgen05.cp, no proxy fence.
// A tile whose boxes would leave its segment's window [0, inAvail) does not go through the ring: the converters read its
// rows from global memory with guards (both sides derive this from the same tile geometry).
// The geometry of a tile comes from its record (UmmaTileRec, written by umma_tile_table_kernel just before this launch).
template <typename T> __device__ __forceinline__ T* ldg_ptr(T* const* p) {
    return reinterpret_cast<T*>(__ldg(reinterpret_cast<const unsigned long long*>(p)));
}
__device__ __forceinline__ int4 ld_rec_tail(const UmmaTileRec* r) { return __ldg(reinterpret_cast<const int4*>(&r->x0)); }   // x0, mapIdx

struct FeedArgs {
    const UmmaTileRec* recs; int p, nStages, stages, myTiles, aCol, aMask, aShift; bool pair;
    uint8_t* ring; uint64_t *full, *empty, *aReady, *slotFree; unsigned* ovf; long long* prof; int dbg;
};
__device__ __forceinline__ void producer_role(const FeedArgs& A, const UmmaTma& TM) {
    int sIdx = 0; uint32_t sPh = 0;
    const uint32_t ring0 = smem_u32(A.ring);
    const uint32_t tileBytes = (uint32_t) (((kRows - 1) * A.p + A.nStages * 32) * 4);
    const uint32_t pfChunk = (tileBytes / (uint32_t) A.nStages + 15u) & ~15u;
    const UmmaTileRec* rec = A.recs + blockIdx.x;
    int4 cur = A.myTiles > 0 ? ld_rec_tail(rec) : make_int4(0, -1, 0, 0);
    TR_INIT(A.prof, 4);
    for (int t = 0; t < A.myTiles; ++t, rec += gridDim.x) {
        // the next tile: its record (x0, mapIdx) for the next iteration and its input range for the L2 prefetch
        int4 nxt = make_int4(0, -1, 0, 0); const char* pf = nullptr;
        if (t + 1 < A.myTiles) {
            const UmmaTileRec* nr = rec + gridDim.x;
            nxt = ld_rec_tail(nr);
            if (nxt.y >= 0) pf = reinterpret_cast<const char*>(ldg_ptr(&nr->in) + __ldg(&nr->l00));
        }
#ifdef F9_DIAG
        if (A.dbg & 8) pf = nullptr;                           // no L2 prefetch
        if ((A.dbg & 16) && t + 2 < A.myTiles) {               // prefetch two tiles ahead
            const UmmaTileRec* nr = rec + 2 * gridDim.x;
            pf = ld_rec_tail(nr).y >= 0 ? reinterpret_cast<const char*>(ldg_ptr(&nr->in) + __ldg(&nr->l00)) : nullptr;
        }
        if ((A.dbg & 64) && cur.y >= 0) cur.x -= (int) (((TM.base0 >> 2) + (unsigned long long) cur.x) & 31ull);     // rows on 128-byte lines (shifted data)
        if (A.dbg & 32) { const int4 f = ld_rec_tail(A.recs + blockIdx.x); if (f.y >= 0 && cur.y >= 0) { cur.x = f.x; cur.y = f.y; } }   // every tile re-reads the CTA's first tile: L2 hits only
#endif
        if (cur.y < 0) {
            // a tile read with guarded loads: its stages still take their ring positions (see converter_role), without a box
            for (int st = 0; st < A.nStages; ++st) {
                mbar_wait_sleep(A.empty + sIdx, sPh ^ 1, kSleepProducerNs);
                mbar_arrive(A.full + sIdx);
                if (pf) l2_prefetch(pf + (size_t) st * pfChunk, pfChunk);
                if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; }
            }
        } else {
            const CUtensorMap* map = &TM.maps[cur.y];
            for (int st = 0; st < A.nStages; ++st) {
                mbar_wait_sleep(A.empty + sIdx, sPh ^ 1, kSleepProducerNs);
                TR_EV(t, 1, st);
#ifdef F9_DIAG
                if (A.dbg & 128) { mbar_arrive(A.full + sIdx); if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; } continue; }   // no box: the converters read stale shared memory
#endif
                mbar_expect_tx(A.full + sIdx, (uint32_t) kTmaStageBytes);
                tma_load_2d(ring0 + (uint32_t) (sIdx * kTmaStageBytes), map, cur.x + st * 32, 0, A.full + sIdx);
                if (pf) l2_prefetch(pf + (size_t) st * pfChunk, pfChunk);
                if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; }
            }
        }
        cur = nxt;
    }
}

// Short windows (Lagrange and the other short kinds at integer decimation: a handful of stages per tile, almost no MMA work): all
// eight converter warps share every stage, a warp converts one 16-sample K step of its 32 rows (tcgen05.st.x16).  The serial chain
// of a stage is shorter this way, which is what matters when the feed is all there is (96 -> 48 k Lagrange: 0.48 ms against
// 0.60 ms with teams).
__device__ __forceinline__ void converter_role_split(const FeedArgs& A, uint32_t tmem, int warp, int lane) {
    constexpr int kConvSplit = 1;
    constexpr int C = 4 / kConvSplit;                          // 16-byte chunks (4 samples) per thread and stage
    const int quarter = warp & 3, part = (warp - kFirstConv) >> 2; // TMEM lane quarter (fixed by the warp id); K step of the stage and its part
    const int h = part / kConvSplit, sub = part % kConvSplit;
    const int rho = quarter * 32 + lane;                       // period row = TMEM lane = row of the box
    const uint32_t tdst = tmem + ((uint32_t) (quarter * 32) << 16) + (uint32_t) (A.aCol + h * 16 + sub * 2 * C);
    const uint32_t ringRow = smem_u32(A.ring) + (uint32_t) (rho * 128);
    const uint32_t sw = (uint32_t) (rho & 7);
    __half2 hmax = __floats2half2_rn(0.f, 0.f);
    const int aMask = A.aMask, aShift = A.aShift;
    int sIdx = 0; uint32_t sPh = 0; int gs = 0;
    struct TileIn { const float* in; long long l00, inAvail; bool viaTma, mask; };
    auto load_rec = [&](const UmmaTileRec* r) {
        TileIn T; T.in = ldg_ptr(&r->in); T.l00 = __ldg(&r->l00); T.inAvail = __ldg(&r->inAvail);
        const int4 tail = ld_rec_tail(r); T.viaTma = tail.y >= 0; T.mask = tail.z != 0; return T;
    };
    const UmmaTileRec* rec = A.recs + blockIdx.x;
    TileIn T = {nullptr, 0, 0, false, false}, N = T;
    if (A.myTiles > 0) N = load_rec(rec);
    for (int t = 0; t < A.myTiles; ++t, rec += gridDim.x) {
        T = N;
        if (t + 1 < A.myTiles) N = load_rec(rec + gridDim.x);                  // consumed one tile later
        const long long lrow = T.l00 + (long long) rho * A.p + h * 16 + sub * 4 * C;
        for (int st = 0; st < A.nStages; ++st, ++gs) {
            float4 v[C];
            mbar_wait_parked(A.full + sIdx, sPh, kParkNs);     // every stage takes a ring position (producer_role), fed by TMA or not
            if (T.viaTma) {
                const uint32_t a = ringRow + (uint32_t) (sIdx * kTmaStageBytes);
                #pragma unroll
                for (int c = 0; c < C; ++c)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w)
                                 : "r"(a + ((((uint32_t) (4 * h + C * sub + c)) ^ sw) << 4)) : "memory");
                if (T.mask) {                                  // the box left the window [0, inAvail): zero what lies outside
                    const long long l0 = lrow + st * 32;
                    const long long lo = -l0, hi = T.inAvail - l0;              // valid element indices e of this thread's samples: lo <= e < hi
                    #pragma unroll
                    for (int c = 0; c < C; ++c) {
                        if (4 * c < lo || 4 * c >= hi) v[c].x = 0.f;
                        if (4 * c + 1 < lo || 4 * c + 1 >= hi) v[c].y = 0.f;
                        if (4 * c + 2 < lo || 4 * c + 2 >= hi) v[c].z = 0.f;
                        if (4 * c + 3 < lo || 4 * c + 3 >= hi) v[c].w = 0.f;
                    }
                }
            } else {
                #pragma unroll
                for (int c = 0; c < C; ++c) {
                    const long long l = lrow + st * 32 + 4 * c;
                    const float* ptr = T.in + l;
                    if (l >= 0 && l + 3 < T.inAvail) v[c] = __ldg(reinterpret_cast<const float4*>(ptr));
                    else {
                        v[c].x = (l >= 0 && l < T.inAvail) ? __ldg(ptr) : 0.f;             v[c].y = (l + 1 >= 0 && l + 1 < T.inAvail) ? __ldg(ptr + 1) : 0.f;
                        v[c].z = (l + 2 >= 0 && l + 2 < T.inAvail) ? __ldg(ptr + 2) : 0.f; v[c].w = (l + 3 >= 0 && l + 3 < T.inAvail) ? __ldg(ptr + 3) : 0.f;
                    }
                }
            }
            uint32_t hd[2 * C], tl[2 * C];
            #pragma unroll
            for (int c = 0; c < C; ++c) {
                float4 xv = v[c];
                xv.x *= kPreScale; xv.y *= kPreScale; xv.z *= kPreScale; xv.w *= kPreScale;
                const __half2 h01 = __floats2half2_rn(xv.x, xv.y), h23 = __floats2half2_rn(xv.z, xv.w);
                hmax = __hmax2_nan(hmax, __hmax2_nan(__habs2(h01), __habs2(h23)));
                hd[2 * c] = *reinterpret_cast<const uint32_t*>(&h01); hd[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                // tail = (x' - head) * 2048, exact: one mixed-precision FMA, head * (-2048) + x' * 2048.  x' * 2048 is an exponent
                // add on the integer pipe (x' = 0 becomes 2^-116, which the fp16 tail rounds to 0; non-finite x' takes the redo).
                float t0, t1, t2, t3;
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tfma.rn.f32.f16 %0, lo, %5, %3;\n\tfma.rn.f32.f16 %1, hi, %5, %4;\n\t}"
                    : "=f"(t0), "=f"(t1) : "r"(hd[2 * c]), "f"(__int_as_float(__float_as_int(xv.x) + (11 << 23))),
                      "f"(__int_as_float(__float_as_int(xv.y) + (11 << 23))), "h"((unsigned short) 0xE800));
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tfma.rn.f32.f16 %0, lo, %5, %3;\n\tfma.rn.f32.f16 %1, hi, %5, %4;\n\t}"
                    : "=f"(t2), "=f"(t3) : "r"(hd[2 * c + 1]), "f"(__int_as_float(__float_as_int(xv.z) + (11 << 23))),
                      "f"(__int_as_float(__float_as_int(xv.w) + (11 << 23))), "h"((unsigned short) 0xE800));
                const __half2 t01 = __floats2half2_rn(t0, t1), t23 = __floats2half2_rn(t2, t3);
                tl[2 * c] = *reinterpret_cast<const uint32_t*>(&t01); tl[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&t23);
            }
            __syncwarp();                                      // the stage's rows are in registers: the ring position may be refilled
            if (lane == 0) mbar_arrive(A.empty + sIdx);
            if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; }
            if (gs > aMask) mbar_wait_parked(A.slotFree + (gs & aMask), (uint32_t) (((gs >> aShift) - 1) & 1), kParkNs);   // the MMAs of the slot's previous stage are done
            tc_fence_after();
            const uint32_t td = tdst + (uint32_t) ((gs & aMask) * 32);           // head columns; the tail sits 8 columns up
            if constexpr (kConvSplit == 1) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             :: "r"(td), "r"(hd[0]), "r"(hd[1]), "r"(hd[2]), "r"(hd[3]), "r"(hd[4 % (2 * C)]), "r"(hd[5 % (2 * C)]), "r"(hd[6 % (2 * C)]), "r"(hd[7 % (2 * C)]),
                                "r"(tl[0]), "r"(tl[1]), "r"(tl[2]), "r"(tl[3]), "r"(tl[4 % (2 * C)]), "r"(tl[5 % (2 * C)]), "r"(tl[6 % (2 * C)]), "r"(tl[7 % (2 * C)]) : "memory");
            } else {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(td), "r"(hd[0]), "r"(hd[1]), "r"(hd[2]), "r"(hd[3]) : "memory");
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(td + 8), "r"(tl[0]), "r"(tl[1]), "r"(tl[2]), "r"(tl[3]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (A.pair) mbar_arrive_leader(A.aReady + (gs & aMask)); else mbar_arrive(A.aReady + (gs & aMask)); }
        }
    }
    const float2 hm = __half22float2(hmax);
    if (!(hm.x < 32768.0f) || !(hm.y < 32768.0f)) atomicOr(A.ovf, 1u);
}

// Two teams of four warps (one per TMEM lane quarter) take the stages alternately: team 0 the even ones, team 1 the odd ones.  A
// warp converts BOTH 16-sample K steps of its 32 rows and writes the whole operand slot with one tcgen05.st.32x32b.x32.  What
// bounded the feed was not the converters' instruction count but the serial chain of a stage inside a warp -- wait for the box,
// LDS, convert, wait for the slot, tcgen05.st, wait::st, arrive: ~900 clk, 17 times per tile in every warp (splitting a stage
// over sixteen warps did not help: 0.74 ms, the chain stays).  With teams a warp walks that chain for every other stage only.
__device__ __forceinline__ void converter_role(const FeedArgs& A, uint32_t tmem, int warp, int lane) {
    const int quarter = warp & 3, team = (warp - kFirstConv) >> 2;             // TMEM lane quarter (fixed by the warp id); the team's stages
    const int rho = quarter * 32 + lane;                       // period row = TMEM lane = row of the box
    const uint32_t tdst = tmem + ((uint32_t) (quarter * 32) << 16) + (uint32_t) A.aCol;
    const uint32_t ringRow = smem_u32(A.ring) + (uint32_t) (rho * 128);
    const uint32_t sw = (uint32_t) (rho & 7);
    const int aMask = A.aMask, aShift = A.aShift;
    int sIdx = 0; uint32_t sPh = 0; int gs = 0;                // ring position / phase of the next TMA stage, global stage count
    struct TileIn { const float* in; long long l00, inAvail; bool viaTma, mask; };
    auto load_rec = [&](const UmmaTileRec* r) {
        TileIn T; T.in = ldg_ptr(&r->in); T.l00 = __ldg(&r->l00); T.inAvail = __ldg(&r->inAvail);
        const int4 tail = ld_rec_tail(r); T.viaTma = tail.y >= 0; T.mask = tail.z != 0; return T;
    };
    const UmmaTileRec* rec = A.recs + blockIdx.x;
    TR_INIT(lane == 0 ? A.prof : nullptr, warp);
    TileIn T = {nullptr, 0, 0, false, false}, N = T;
    if (A.myTiles > 0) N = load_rec(rec);
    for (int t = 0; t < A.myTiles; ++t, rec += gridDim.x) {
        T = N;
        if (t + 1 < A.myTiles) N = load_rec(rec + gridDim.x);                  // consumed one tile later
        const long long lrow = T.l00 + (long long) rho * A.p;
        for (int st = 0; st < A.nStages; ++st, ++gs) {
            if (gs % kConvTeams != team) {                     // another team's stage: only the ring position moves
                if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; }
                continue;
            }
            // EVERY stage takes a ring position, also those of a tile that is not fed by TMA (the producer then completes the
            // position's "full" phase without a box).  The ring depth is even, so a ring position -- its full / empty barriers --
            // always belongs to the same team, which therefore sees every phase of them: a parity wait only tells the current
            // phase from the one before it, and a warp that saw every other phase could take a box still in flight for landed.
            mbar_wait(A.full + sIdx, sPh);                     // plain polls here and on the slot: -2.5 % against parked waits (the teams are
            TR_EV(t, 10, st);                                  // the pair's pace-setters; the other roles keep their suspend hints)
            float4 v[8];
#ifdef F9_DIAG
            if (A.dbg & 256) {                                 // no LDS: constant samples
                #pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
            } else
#endif
            if (T.viaTma) {
                const uint32_t a = ringRow + (uint32_t) (sIdx * kTmaStageBytes);
                #pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w)
                                 : "r"(a + ((((uint32_t) c) ^ sw) << 4)) : "memory");
                if (T.mask) {                                  // the box left the window [0, inAvail): zero what lies outside
                    const long long l0 = lrow + st * 32;
                    const long long lo = -l0, hi = T.inAvail - l0;              // valid element indices e of this thread's samples: lo <= e < hi
                    #pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (4 * c < lo || 4 * c >= hi) v[c].x = 0.f;
                        if (4 * c + 1 < lo || 4 * c + 1 >= hi) v[c].y = 0.f;
                        if (4 * c + 2 < lo || 4 * c + 2 >= hi) v[c].z = 0.f;
                        if (4 * c + 3 < lo || 4 * c + 3 >= hi) v[c].w = 0.f;
                    }
                }
            } else {
                #pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const long long l = lrow + st * 32 + 4 * c;
                    const float* ptr = T.in + l;
                    if (l >= 0 && l + 3 < T.inAvail) v[c] = __ldg(reinterpret_cast<const float4*>(ptr));
                    else {
                        v[c].x = (l >= 0 && l < T.inAvail) ? __ldg(ptr) : 0.f;             v[c].y = (l + 1 >= 0 && l + 1 < T.inAvail) ? __ldg(ptr + 1) : 0.f;
                        v[c].z = (l + 2 >= 0 && l + 2 < T.inAvail) ? __ldg(ptr + 2) : 0.f; v[c].w = (l + 3 >= 0 && l + 3 < T.inAvail) ? __ldg(ptr + 3) : 0.f;
                    }
                }
            }
            // w[0..7] head of K step 0, w[8..15] its tail, w[16..23] head of K step 1, w[24..31] its tail: the slot's 32 columns
            uint32_t w[32];
            #pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 xv = v[c];
                xv.x *= kPreScale; xv.y *= kPreScale; xv.z *= kPreScale; xv.w *= kPreScale;
                const __half2 h01 = __floats2half2_rn(xv.x, xv.y), h23 = __floats2half2_rn(xv.z, xv.w);
                const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
                // tail = (x' - head) * 2048, exact: one mixed-precision FMA, head * (-2048) + x' * 2048.  x' * 2048 is an exponent
                // add on the integer pipe (x' = 0 becomes 2^-116, which the fp16 tail rounds to 0; non-finite x' takes the redo).
                float t0, t1, t2, t3;
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tfma.rn.f32.f16 %0, lo, %5, %3;\n\tfma.rn.f32.f16 %1, hi, %5, %4;\n\t}"
                    : "=f"(t0), "=f"(t1) : "r"(u01), "f"(__int_as_float(__float_as_int(xv.x) + (11 << 23))),
                      "f"(__int_as_float(__float_as_int(xv.y) + (11 << 23))), "h"((unsigned short) 0xE800));
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tfma.rn.f32.f16 %0, lo, %5, %3;\n\tfma.rn.f32.f16 %1, hi, %5, %4;\n\t}"
                    : "=f"(t2), "=f"(t3) : "r"(u23), "f"(__int_as_float(__float_as_int(xv.z) + (11 << 23))),
                      "f"(__int_as_float(__float_as_int(xv.w) + (11 << 23))), "h"((unsigned short) 0xE800));
                const __half2 t01 = __floats2half2_rn(t0, t1), t23 = __floats2half2_rn(t2, t3);
                const int base = (c >> 2) * 16 + (c & 3) * 2;                  // K step c / 4, words 2 (c % 4), + 1 of its head
                w[base] = u01; w[base + 1] = u23;
                w[base + 8] = *reinterpret_cast<const uint32_t*>(&t01); w[base + 9] = *reinterpret_cast<const uint32_t*>(&t23);
            }
            __syncwarp();                                      // the stage's rows are in registers: the ring position may be refilled
            TR_EV(t, 11, st);
            if (lane == 0) mbar_arrive(A.empty + sIdx);
            if (++sIdx == A.stages) { sIdx = 0; sPh ^= 1; }
            if (gs > aMask) mbar_wait(A.slotFree + (gs & aMask), (uint32_t) (((gs >> aShift) - 1) & 1));   // the MMAs of the slot's previous stage are done
            tc_fence_after();
            TR_EV(t, 12, st);
            const uint32_t td = tdst + (uint32_t) ((gs & aMask) * 32);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                         :: "r"(td), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]),
                            "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]),
                            "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            TR_EV(t, 13, st);
            if (lane == 0) { if (A.pair) mbar_arrive_leader(A.aReady + (gs & aMask)); else mbar_arrive(A.aReady + (gs & aMask)); }
            TR_EV(t, 14, st);
        }
    }
    // Samples outside the fp16 split's range are not looked for here: |128 x| >= 65520, Inf and NaN become an infinite or NaN head,
    // every product with it is Inf or NaN (0 * Inf included), so the outputs it reaches are not finite and the EPILOGUE raises the
    // redo flag (one FFMA per output there against five half2 operations per four samples here, on the path that bounds the kernel).
}

// out = (D0A + D0B + D1 / 2048) / kPreScale: the scalings are powers of two, the only roundings are the two additions
__device__ __forceinline__ float combine(uint32_t d0a, uint32_t d0b, uint32_t d1) {
    return fmaf(__uint_as_float(d1), 1.0f / (2048.0f * kPreScale), (__uint_as_float(d0a) + __uint_as_float(d0b)) * (1.0f / kPreScale));
}

// CTA2: two CTAs of a cluster work as a pair (tcgen05.mma.cta_group::2, M = 256): each converts and stores its own tile of 128
// periods, each holds the weights of half the slots of every group (half the shared memory: the input ring gets 7 stages
// instead of 2), the leader issues every MMA for both.  Tile t of the pair is recs[blockIdx.x + i * gridDim.x] as before; the
// table is padded so that the odd CTA's last tile exists (a record without input or output).
template <bool MERGED, bool TMA, bool CTA2>
__global__ void __launch_bounds__(TMA ? kThreadsTma : kThreads, 1)
umma_fir_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles,
                const __grid_constant__ UmmaDev P, const __grid_constant__ UmmaTma TM, const UmmaTileRec* __restrict__ recs, int stages, int alignedAll,
                unsigned* __restrict__ ovf, long long* __restrict__ prof, int dbg) {
    extern __shared__ __align__(128) uint8_t smem[];
#ifdef F9_DIAG
    if (TMA && prof && blockIdx.x < 2 && threadIdx.x == 0)      // kernel entry of the traced CTAs
        prof[((size_t) blockIdx.x * kTrWarps + 4) * kTrCap + kTrCap - 2] = (0x7ell << 56) | (clock64() & 0xffffffffffll);
#endif
    const SmemMap sm = carve(smem, P.maxEntries, P.NB, stages, TMA, CTA2);
    const uint32_t rank = CTA2 ? cluster_rank() : 0u;        // 0: leader (issues the MMAs)
    const int NB = P.NB;                                       // slots per group: MMA N
    const int aMask = P.aSlots - 1, aShift = P.aSlots == 4 ? 2 : 1, aCol = 512 - 32 * P.aSlots;   // TMEM operand ring
    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool teams = P.taps >= 64;                           // long windows: converter teams take the stages alternately (converter_role)
    const int convPerStage = teams ? kConvPerStage : kConvWarps;       // converter warps that take part in one stage
    // slot block of this CTA: tile t belongs to block t % nGB; with CTA pairs tiles are dealt two at a time, (t / 2) % nGB, so
    // that both CTAs of a pair hold the same weights (the grid is a multiple of 2 * nGB: the block never changes)
    const int gb = CTA2 ? (int) (blockIdx.x >> 1) % P.nGB : (int) blockIdx.x % P.nGB;
    const UmmaBlockInfo& BI = P.blk[gb];
    const int p = P.p, q = P.q;
    const int nStages = BI.nStages;                            // stages per tile (two K steps each)
    const int firstTile = CTA2 ? (int) (blockIdx.x & ~1u) : (int) blockIdx.x;     // both CTAs of a pair walk the same number of tiles
    const int myTiles = firstTile < nTiles ? (nTiles - 1 - firstTile) / (int) gridDim.x + 1 : 0;

    // ---- one-time setup: weights into shared memory, barriers, TMEM
    {
        // 90 - 180 KB of weights per CTA: eight 16-byte loads in flight per thread.  One load per loop iteration waited for one L2
        // latency twenty times over: 15 500 clk (7.9 us) from kernel entry to the end of this set-up in an event trace of config 1,
        // a quarter of that launch.
        const uint4* src = reinterpret_cast<const uint4*>(P.W + (CTA2 ? BI.w2Off[rank] : BI.wOff));
        uint4* dst = reinterpret_cast<uint4*>(sm.W);
        const int nW16 = BI.nEntries * NB * (CTA2 ? 2 : 4);
        for (int i0 = threadIdx.x; i0 < nW16; i0 += 8 * (int) blockDim.x) {
            uint4 r[8];
            #pragma unroll
            for (int k = 0; k < 8; ++k) { const int i = i0 + k * (int) blockDim.x; r[k] = i < nW16 ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u); }
            #pragma unroll
            for (int k = 0; k < 8; ++k) { const int i = i0 + k * (int) blockDim.x; if (i < nW16) dst[i] = r[k]; }
        }
        const uint4* osrc = reinterpret_cast<const uint4*>(P.W + (TMA ? BI.opOffT : BI.opOff));
        for (int i = threadIdx.x; i < BI.nEntries; i += blockDim.x) sm.ops[i] = __ldg(osrc + i);
        if (threadIdx.x == 0) {
            // register loader: full <- 8 loader warps, empty <- the copy warp's commit, cpDone <- its commit
            // TMA feed:        full <- the producer's expect_tx, empty <- 8 converter warps, cpDone ("operand ready") <- 8 converter warps
            for (int s = 0; s < stages; ++s) { mbar_init(sm.full + s, TMA ? 1 : kLoaderWarps); mbar_init(sm.empty + s, TMA ? convPerStage : 1); }
            // CTA pairs: "operand ready" and "accumulator drained" collect both CTAs' arrivals in the leader
            for (int g = 0; g < kUmmaMaxGroups; ++g) { mbar_init(sm.accFull + g, 1); mbar_init(sm.accEmpty + g, CTA2 ? 8 : 4); }
            for (int i = 0; i < kASlotsMax; ++i) { mbar_init(sm.cpDone + i, TMA ? (CTA2 ? 2 : 1) * convPerStage : 1); mbar_init(sm.slotFree + i, TMA ? kIssuersTma : kIssuers); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 4) {
            if (CTA2) {
                asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sm.tmemSlot)), "r"(512) : "memory");
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
            } else {
                asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(sm.tmemSlot)), "r"(512) : "memory");
                asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
            }
        }
        fence_async_smem();                                    // the weights are read by the tensor pipe (async proxy)
        tc_fence_before();
        __syncthreads();
        if (CTA2) cluster_sync_all();                          // the peer's barriers and weights are ready before anything crosses over
#ifdef F9_DIAG
        if (TMA && prof && blockIdx.x < 2 && threadIdx.x == 0)  // end of the set-up; both CTAs of a pair leave the cluster barrier together: offset between their clocks
            prof[((size_t) blockIdx.x * kTrWarps + 4) * kTrCap + kTrCap - 1] = (0x7fll << 56) | (clock64() & 0xffffffffffll);
#endif
        tc_fence_after();
    }
    const uint32_t tmem = *sm.tmemSlot;
    // optional cycle accounting (development): per CTA [0] loader-warp-0 total, [1] its wait on "stage free", [2] tensor warp
    // total, [3] its wait on "stage full", [4] its wait on "accumulator drained", [5] epilogue-warp-0 total, [6] its wait on
    // "group done", [7] its time in the store loops
    long long pT0 = 0, pW0 = 0, pW1 = 0, pW2 = 0, pF = 0, pC = 0, pI = 0;
    #define PROF_BEGIN(v) long long v = (!TMA && prof) ? clock64() : 0
    #define PROF_END(acc, v) if (!TMA && prof) acc += clock64() - v
    if (!TMA && prof) pT0 = clock64();

    if (TMA && (warp >= kFirstConv || warp == 4)) {
        // =========================================================== TMA producer / converters
        FeedArgs FA;
        FA.recs = recs; FA.p = p; FA.aCol = aCol; FA.aMask = aMask; FA.aShift = aShift; FA.pair = CTA2;
        FA.nStages = nStages; FA.stages = stages; FA.myTiles = myTiles;
        FA.ring = sm.ring; FA.full = sm.full; FA.empty = sm.empty; FA.aReady = sm.cpDone; FA.slotFree = sm.slotFree; FA.ovf = ovf; FA.prof = prof; FA.dbg = dbg;
        if (warp == 4) { if (lane == 0) producer_role(FA, TM); }
        else if (teams) converter_role(FA, tmem, warp, lane);
        else converter_role_split(FA, tmem, warp, lane);
    } else if (!TMA && warp >= kFirstLoader) {
        // =========================================================== loaders
        LoaderArgs LA;
        LA.segs = segs; LA.tilePrefix = tilePrefix; LA.nSegs = nSegs; LA.nTiles = nTiles; LA.nGB = P.nGB; LA.p = p; LA.q = q;
        LA.U0 = BI.U0; LA.nStages = nStages; LA.stages = stages; LA.myTiles = myTiles;
        LA.ring = sm.ring; LA.full = sm.full; LA.empty = sm.empty; LA.ovf = ovf; LA.prof = prof;
        if (alignedAll) loader_role<true>(LA, warp - kFirstLoader, lane);
        else loader_role<false>(LA, warp - kFirstLoader, lane);
    } else if (warp == 4) {
        // =========================================================== copy warp
        // Stage gs (32 samples of all 128 rows, head + tail) goes from shared memory into TMEM operand slot gs & 1 with four
        // tcgen05.cp.  A single warp issuing copies AND MMAs was bound by its own instruction latency (the uniform datapath
        // runs ~4 clk per instruction: ~300 clk per K step before the first MMA), so the MMAs are issued by other warps:
        // cpDone[slot] tells them the operand is in TMEM, slotFree[slot] tells this warp they have finished reading it.
        const uint32_t el = elect_one();
        const uint64_t aDesc0 = make_desc(smem_u32(sm.ring), kChunk, 128);
        const int totalStages = myTiles * nStages;
        int sIdx = 0; uint32_t sPh = 0;
        for (int gs = 0; gs < totalStages; ++gs) {
            { PROF_BEGIN(w); mbar_wait(sm.full + sIdx, sPh); PROF_END(pW0, w); }
            PROF_BEGIN(wf);
            fence_async_smem();                                // generic-proxy stores of the loaders -> async-proxy reads of tcgen05.cp
            PROF_END(pF, wf);
            if (gs > aMask) { PROF_BEGIN(w); mbar_wait(sm.slotFree + (gs & aMask), (uint32_t) (((gs >> aShift) - 1) & 1)); PROF_END(pW1, w); }
            tc_fence_after();
            PROF_BEGIN(wc);
            if (el) {
                const uint64_t ad = aDesc0 + (uint64_t) ((sIdx * kStageBytes) >> 4);
                const uint32_t slot0 = tmem + (uint32_t) aCol + (uint32_t) ((gs & aMask) * 32);
                if (!(dbg & 2)) {
                umma_cp(slot0, ad);                                             // first K step  head
                umma_cp(slot0 + 8, ad + ((4 * kChunk) >> 4));                   //               tail
                umma_cp(slot0 + 16, ad + ((2 * kChunk) >> 4));                  // second K step head
                umma_cp(slot0 + 24, ad + ((6 * kChunk) >> 4));                  //               tail
                }
                umma_commit(sm.empty + sIdx);                  // the stage is free once the copies have read it
                umma_commit(sm.cpDone + (gs & aMask));
            }
            __syncwarp();
            PROF_END(pC, wc);
            if (++sIdx == stages) { sIdx = 0; sPh ^= 1; }
        }
        if (prof && lane == 0) { prof[blockIdx.x * 16 + 8] = pF; prof[blockIdx.x * 16 + 9] = pC; }
        if (prof && lane == 0) { prof[blockIdx.x * 16 + 2] = clock64() - pT0; prof[blockIdx.x * 16 + 3] = pW0; prof[blockIdx.x * 16 + 4] = pW1; }
    } else if (warp > 4 && (!CTA2 || rank == 0)) {
        // =========================================================== MMA issuers (CTA pairs: the leader's only)
        // Warp w owns the groups gl = w, w + kIssuers, ...: per group the operands advance by constants from K step to K step
        // (weight tiles are stored group-major), so the issue loop is a few uniform adds around three tcgen05.mma.
        const int w = warp - 5;
        const uint32_t el = elect_one();
        const uint32_t nb = (uint32_t) NB;
        const uint32_t idescN = make_idesc(CTA2 ? 2 * kRows : kRows, NB), idesc2N = make_idesc(kRows, 2 * NB);
        const uint64_t wDesc0 = make_desc(smem_u32(sm.W), CTA2 ? 16 * NB : 32 * NB, 128);   // weight tile i: + 4*NB*i (16-byte units; half for pairs); w1 rows at + NB (+ NB/2)
        // The warp walks its host-built list (UmmaOp, shared memory) once per tile.  The list is first rewritten in place into
        // what the issue needs verbatim: {TMEM address of D0 (or its pool slot), TMEM address of D1, low word of the weight
        // tile's descriptor, flags | h << 7 | gl << 8 | waitGl << 16 | stage << 24}, so that an entry costs one 128-bit load
        // (fetched an entry ahead) and one block of predicated tcgen05 instructions.  A single thread's instruction latency is
        // what bounds this role: ~100 instructions per entry (flag decoding, 64-bit descriptor arithmetic) cost ~360 clk.
        uint4* ops = sm.ops + (TMA ? BI.opStartT[w] : BI.opStart[w]);
        int nOps = TMA ? BI.opStartT[w + 1] - BI.opStartT[w] : BI.opStart[w + 1] - BI.opStart[w];
        const uint32_t wLo = (uint32_t) (wDesc0 & 0xffffffffull), wHi = (uint32_t) (wDesc0 >> 32);
        enum : uint32_t { fA0 = 1, fA1 = 2, fM = 4, fDrain = 8, fPool = 16, fPrev = 32, fLast = 64, fH = 128, fPair = 0x1000 };   // gl sits in bits 8-11
        for (int i = lane; i < nOps; i += 32) {
            const uint4 o = ops[i];
            const uint32_t d1c = o.x & 0xffffu, poolc = o.x >> 16, hh = (o.z >> 8) & 0xffu, gl = (o.z >> 16) & 0xffu, fl = o.z >> 24;
            const bool merged = !CTA2 && (MERGED || (fl & kOpMerged)), toPool = !(MERGED || (fl & kOpMerged));
            uint32_t f = 0;
            if (toPool ? (fl & kOpPoolAcc) : (fl & kOpAcc)) f |= fA0;
            if (fl & kOpAcc) f |= fA1;
            if (merged) f |= fM;
            if (fl & kOpWaitDrain) f |= fDrain; else if (fl & kOpWaitPool) f |= fPool;
            if (fl & kOpPoolPrevTile) f |= fPrev;
            if (fl & kOpLast) f |= fLast;
            if (hh) f |= fH;
            ops[i] = make_uint4(tmem + (toPool ? poolc : d1c - nb), tmem + d1c, wLo + (CTA2 ? o.y >> 1 : o.y),
                                f | (gl << 8) | ((o.w & 0xffu) << 16) | ((o.z & 0xffu) << 24));
        }
        __syncwarp();
        // CTA pairs: the two K steps of a stage that a group takes back to back become ONE record (six MMAs under one decode).
        // The single issuing thread's instruction latency bounds this role and, through the operand ring, the kernel: an event
        // trace of a -DF9_DIAG build (tools/umma_trace.py) showed ~500 clk per record against ~100 clk of tensor time, two records
        // per warp and stage, and every other role waiting for it (with boxes, conversion, MMAs and stores all switched off the
        // kernel took as long as with them).
        const uint32_t tileUnits = CTA2 ? 2 * nb : 4 * nb;     // weight tile pitch in 16-byte units (tiles are stored group-major)
        if (CTA2) {
            int nNew = 0;
            if (lane == 0) {
                int i = 0;
                while (i < nOps) {
                    uint4 a = ops[i];
                    if (i + 1 < nOps) {
                        const uint4 b = ops[i + 1];
                        const bool pair = !(a.w & fH) && (b.w & fH) && (a.w >> 24) == (b.w >> 24) && ((a.w >> 8) & 0xfu) == ((b.w >> 8) & 0xfu) &&
                                          !(b.w & (fDrain | fPool)) && b.x == a.x && b.y == a.y && b.z == a.z + tileUnits &&
                                          (b.w & fA0) && (b.w & fA1) && !(a.w & fLast);
                        if (pair) { a.w |= fPair | (b.w & fLast); ops[nNew++] = a; i += 2; continue; }
                    }
                    ops[nNew++] = a; ++i;
                }
            }
            nOps = __shfl_sync(0xffffffffu, nNew, 0);
            __syncwarp();
        }
        const uint32_t go = (dbg & 1) ? 0u : 1u;
        const uint32_t w1Off = CTA2 ? nb / 2 : nb;             // the w1 rows of a tile, in 16-byte units
        const uint32_t accFull0 = smem_u32(sm.accFull);
        int gs = 0;
        TR_INIT((TMA && el) ? prof : nullptr, warp);
        for (int t = 0; t < myTiles; ++t) {
            int k = 0;
            uint4 nx = nOps > 0 ? ops[0] : make_uint4(0, 0, 0, 0xff000000u);
            for (int st = 0; st < nStages; ++st, ++gs) {
                { PROF_BEGIN(wq); if (CTA2) mbar_wait_cluster(sm.cpDone + (gs & aMask), (uint32_t) ((gs >> aShift) & 1), kParkNs); else if (TMA) mbar_wait_parked(sm.cpDone + (gs & aMask), (uint32_t) ((gs >> aShift) & 1), kParkNs); else mbar_wait(sm.cpDone + (gs & aMask), (uint32_t) ((gs >> aShift) & 1)); PROF_END(pW0, wq); }
                tc_fence_after();
                TR_EV(t, 20, st);
                PROF_BEGIN(wi);
                const uint32_t aSlot = tmem + (uint32_t) aCol + (uint32_t) ((gs & aMask) * 32);
                while (k < nOps && (int) (nx.w >> 24) == st) {
                    const uint4 o = nx;
                    ++k;
                    nx = ops[k < nOps ? k : 0];
                    if (o.w & (fDrain | fPool)) {                               // rare: first K step of a group / first step past the split
                        const uint32_t gl = (o.w >> 8) & 0xfu, waitGl = (o.w >> 16) & 0xffu;
                        uint64_t* bar = sm.accEmpty + ((o.w & fDrain) ? gl : waitGl);
                        const uint32_t par = ((o.w & fDrain) || (o.w & fPrev)) ? (uint32_t) ((t & 1) ^ 1) : (uint32_t) (t & 1);
                        TR_EV(t, 22, st);
                        PROF_BEGIN(wd); if (CTA2) mbar_wait_cluster(bar, par, 200); else mbar_wait(bar, par); PROF_END(pW1, wd);
                        tc_fence_after();
                        TR_EV(t, 23, st);
                    }
                    // x0 * w0 -> D0 (N = 2 NB over [D0 | D1] when merged), x0 * w1 -> D1 (unless merged), x1 * w0 -> D1, "group done"
                    // (issued under the elected lane's branch: a per-lane predicate on these warp-uniform instructions makes the
                    // compiler serialise them with an election loop)
                    const uint32_t aHi = aSlot + ((o.w & fH) ? 16u : 0u);
                    const uint32_t lastBar = accFull0 + ((o.w >> 8) & 0xfu) * 8u;
                    if (!el) continue;
                    if (CTA2)
                        // K step h (or 0 of a pair): x0 w0 -> D0, x0 w1 -> D1, x1 w0 -> D1; a pair: the same for K step 1 (operand 16
                        // columns up, the group's next weight tile), always accumulating; then "group done"
                        asm volatile("{\n\t.reg .pred pe, pp, pa0, pa1, pl;\n\t.reg .b32 t;\n\t.reg .b64 b0, b1, b2, b3;\n\t"
                                     "setp.ne.b32 pe, %0, 0;\n\t"
                                     "and.b32 t, %1, 1;\n\tsetp.ne.b32 pa0, t, 0;\n\tand.b32 t, %1, 2;\n\tsetp.ne.b32 pa1, t, 0;\n\t"
                                     "and.b32 t, %1, 4096;\n\tsetp.ne.and.b32 pp, t, 0, pe;\n\t"
                                     "and.b32 t, %1, 64;\n\tsetp.ne.b32 pl, t, 0;\n\t"
                                     "mov.b64 b0, {%2, %3};\n\tadd.u32 t, %2, %4;\n\tmov.b64 b1, {t, %3};\n\t"
                                     "add.u32 t, %2, %10;\n\tmov.b64 b2, {t, %3};\n\tadd.u32 t, t, %4;\n\tmov.b64 b3, {t, %3};\n\t"
                                     "@pe tcgen05.mma.cta_group::2.kind::f16 [%5], [%7], b0, %9, pa0;\n\t"
                                     "@pe tcgen05.mma.cta_group::2.kind::f16 [%6], [%7], b1, %9, pa1;\n\t"
                                     "@pe tcgen05.mma.cta_group::2.kind::f16 [%6], [%8], b0, %9, 1;\n\t"
                                     "@pp tcgen05.mma.cta_group::2.kind::f16 [%5], [%13], b2, %9, 1;\n\t"
                                     "@pp tcgen05.mma.cta_group::2.kind::f16 [%6], [%13], b3, %9, 1;\n\t"
                                     "@pp tcgen05.mma.cta_group::2.kind::f16 [%6], [%14], b2, %9, 1;\n\t"
                                     "@pl tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%11], %12;\n\t}"
                                     :: "r"(go), "r"(o.w), "r"(o.z), "r"(wHi), "r"(w1Off), "r"(o.x), "r"(o.y), "r"(aHi), "r"(aHi + 8u), "r"(idescN),
                                        "r"(tileUnits), "r"(lastBar), "h"((uint16_t) 3), "r"(aHi + 16u), "r"(aHi + 24u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred pe, pn, pa0, pa1, pm, pl;\n\t.reg .b32 t, id;\n\t.reg .b64 b0, b1;\n\t"
                                     "setp.ne.b32 pe, %0, 0;\n\t"
                                     "and.b32 t, %1, 1;\n\tsetp.ne.b32 pa0, t, 0;\n\tand.b32 t, %1, 2;\n\tsetp.ne.b32 pa1, t, 0;\n\t"
                                     "and.b32 t, %1, 4;\n\tsetp.ne.b32 pm, t, 0;\n\tselp.b32 id, %12, %9, pm;\n\tand.pred pn, pe, !pm;\n\t"
                                     "setp.ne.b32 pl, %10, 0;\n\tand.b32 t, %1, 64;\n\tsetp.ne.and.b32 pl, t, 0, pl;\n\t"
                                     "mov.b64 b0, {%2, %3};\n\tadd.u32 t, %2, %4;\n\tmov.b64 b1, {t, %3};\n\t"
                                     "@pe tcgen05.mma.cta_group::1.kind::f16 [%5], [%7], b0, id, pa0;\n\t"
                                     "@pn tcgen05.mma.cta_group::1.kind::f16 [%6], [%7], b1, %9, pa1;\n\t"
                                     "@pe tcgen05.mma.cta_group::1.kind::f16 [%6], [%8], b0, %9, 1;\n\t"
                                     "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t}"
                                     :: "r"(go), "r"(o.w), "r"(o.z), "r"(wHi), "r"(w1Off), "r"(o.x), "r"(o.y), "r"(aHi), "r"(aHi + 8u), "r"(idescN),
                                        "r"(1u), "r"(lastBar), "r"(idesc2N) : "memory");
                }
                if (el) { if (CTA2) umma2_commit_both(sm.slotFree + (gs & aMask)); else umma_commit(sm.slotFree + (gs & aMask)); }   // arrives once this warp's MMAs on the slot have completed
                TR_EV(t, 21, st);
                __syncwarp();
                PROF_END(pI, wi);
            }
        }
        if (prof && w == 0 && lane == 0) { prof[blockIdx.x * 16 + 10] = pW0; prof[blockIdx.x * 16 + 11] = pI; prof[blockIdx.x * 16 + 12] = pW1; }
    } else if (warp < 4) {
        // =========================================================== epilogue
        // Thread = TMEM lane = period row.  Per group and 16-slot chunk: read D0 (+ its pool half) and D1, combine, transpose
        // through the warp's own slice of the staging buffer, store two rows per instruction (16 slots = 64 contiguous bytes).
        const int row0 = warp * 32 + lane;
        const uint32_t laneBase = (uint32_t) (warp * 32) << 16;
        const int chunks = NB / 16;
        int tileId = blockIdx.x;
        struct TileOut { float* out; long long oBase, numOut; };
        auto load_out = [&](int id) {
            TileOut O;
            if (TMA) { const UmmaTileRec* r = recs + id; O.out = ldg_ptr(&r->out); O.oBase = __ldg(&r->oBase); O.numOut = __ldg(&r->numOut); }
            else {
                const int sidx = find_seg(tilePrefix, nSegs, id);
                const Seg S = segs[sidx];
                const int pb = (id - tilePrefix[sidx]) / P.nGB;
                const long long A0 = S.n0 / q + (long long) pb * kRows;
                O.out = S.out; O.oBase = A0 * q + BI.slot0 - S.n0; O.numOut = S.numOut;
            }
            return O;
        };
        TileOut ON = {nullptr, 0, 0};
        if (myTiles > 0) ON = load_out(tileId);
        TR_INIT((TMA && lane == 0) ? prof : nullptr, warp);
        float sticky = 0.0f;                                   // stays 0 while every output is finite: fma(v, 0, sticky) turns Inf / NaN into NaN
        for (int t = 0; t < myTiles; ++t, tileId += gridDim.x) {
            const TileOut S = ON;
            if (t + 1 < myTiles) ON = load_out(tileId + (int) gridDim.x);     // consumed one tile later
            const long long oBase = S.oBase;                                    // output index of (row 0, slot 0 of the block)
            const int blockSlots = min(BI.nGroups * NB, q - BI.slot0);          // real (non-padding) slots of this block
            const bool rowsInside = oBase >= 0 && oBase + (long long) (kRows - 1) * q + blockSlots <= S.numOut;
            float* outG = reinterpret_cast<float*>(__cvta_generic_to_global(S.out));
            for (int gl = 0; gl < BI.nGroups; ++gl) {
                { PROF_BEGIN(w); if (TMA) mbar_wait_sleep(sm.accFull + gl, t & 1, kSleepEpilogueNs); else mbar_wait_parked(sm.accFull + gl, t & 1, 2000); PROF_END(pW0, w); }
                tc_fence_after();
                TR_EV(t, 30, gl);
                for (int h = 0; h < chunks; ++h) {
                    uint32_t v0[16], v1[16], vb[16];
                    const uint32_t c0 = tmem + laneBase + (uint32_t) (gl * 2 * NB + h * 16);
                    #define LD16(arr, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                        : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), \
                          "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), "=r"(arr[15]) : "r"(addr))
                    LD16(v0, c0);
                    LD16(v1, c0 + (uint32_t) NB);
                    if (P.poolN > 0) { LD16(vb, tmem + laneBase + (uint32_t) (P.GBL * 2 * NB + (gl & (P.poolN - 1)) * NB + h * 16)); }
                    else {
                        #pragma unroll
                        for (int c = 0; c < 16; ++c) vb[c] = 0u;
                    }
                    #undef LD16
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (h == chunks - 1) {                     // the whole group has been read: the next tile may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (CTA2) mbar_arrive_leader(sm.accEmpty + gl); else mbar_arrive(sm.accEmpty + gl); }
                    }
                    float4* dst = reinterpret_cast<float4*>(sm.epi + row0 * kEpiPitch);
                    #pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 o4 = make_float4(combine(v0[4 * c], vb[4 * c], v1[4 * c]), combine(v0[4 * c + 1], vb[4 * c + 1], v1[4 * c + 1]),
                                                      combine(v0[4 * c + 2], vb[4 * c + 2], v1[4 * c + 2]), combine(v0[4 * c + 3], vb[4 * c + 3], v1[4 * c + 3]));
                        if (TMA) sticky = fmaf(o4.x, 0.0f, fmaf(o4.y, 0.0f, fmaf(o4.z, 0.0f, fmaf(o4.w, 0.0f, sticky))));
                        dst[c] = o4;
                    }
                    __syncwarp();
                    PROF_BEGIN(wst);
                    const int col = gl * NB + h * 16 + (lane & 15);           // slot inside the block
                    const int rsel = lane >> 4;
                    const float* src = sm.epi + (warp * 32 + rsel) * kEpiPitch + (lane & 15);
                    const long long o0 = oBase + (long long) (warp * 32 + rsel) * q + col;
                    float* dstp = outG + o0;
                    if (dbg & 4) {} else if (rowsInside && gl * NB + h * 16 + 16 <= blockSlots) {
                        float ov[16];                          // all loads first: a volatile store after each load serialised them
                        #pragma unroll
                        for (int r = 0; r < 16; ++r) ov[r] = src[r * 2 * kEpiPitch];
                        #pragma unroll
                        for (int r = 0; r < 16; ++r, dstp += 2 * q)
                            asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(dstp), "f"(ov[r]));
                    } else {
                        const bool slotOk = col < blockSlots;
                        for (int r = 0; r < 16; ++r, dstp += 2 * q) {
                            const long long o = o0 + (long long) r * 2 * q;
                            if (slotOk && o >= 0 && o < S.numOut) *dstp = src[r * 2 * kEpiPitch];
                        }
                    }
                    __syncwarp();
                    PROF_END(pW2, wst);
                }
                TR_EV(t, 31, gl);
            }
        }
        if (TMA && !(sticky == 0.0f)) atomicOr(ovf, 1u);       // an input sample was outside the fp16 split's range (see converter_role)
        if (prof && warp == 0 && lane == 0) { prof[blockIdx.x * 16 + 5] = clock64() - pT0; prof[blockIdx.x * 16 + 6] = pW0; prof[blockIdx.x * 16 + 7] = pW2; }
    }

    tc_fence_before();
    __syncthreads();
#ifdef F9_DIAG
    if (TMA && prof && blockIdx.x < 2 && threadIdx.x == 0)      // all roles of the CTA are done
        prof[((size_t) blockIdx.x * kTrWarps + 4) * kTrCap + kTrCap - 3] = (0x7dll << 56) | (clock64() & 0xffffffffffll);
#endif
    if (CTA2) cluster_sync_all();                              // the peer may still read this CTA's weights / signal its barriers
    tc_fence_after();
    if (warp == 4) {
        if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

// One thread per tile: the record the TMA-fed kernel's roles read instead of searching the segment table.
__global__ void __launch_bounds__(256)
umma_tile_table_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles,
                       const __grid_constant__ UmmaDev P, const __grid_constant__ UmmaTma TM, UmmaTileRec* __restrict__ recs, int pairs) {
    const int tileId = blockIdx.x * blockDim.x + threadIdx.x;
    if (tileId > nTiles) return;
    if (tileId == nTiles) {                                     // padding for CTA pairs: a tile without input or output
        UmmaTileRec R; R.in = nullptr; R.out = nullptr; R.l00 = 0; R.inAvail = 0; R.oBase = 0; R.numOut = 0; R.x0 = 0; R.mapIdx = -1; R.mask = 0; R.pad = 0;
        recs[tileId] = R;
        return;
    }
    // every segment owns a multiple of nGB tiles (of 2 * nGB with CTA pairs, which take two period blocks of one slot block)
    const int sidx = find_seg(tilePrefix, nSegs, tileId);
    const Seg S = segs[sidx];
    const int local = tileId - tilePrefix[sidx];
    const int gbT = pairs ? (local >> 1) % P.nGB : local % P.nGB;
    const int pb = pairs ? 2 * ((local >> 1) / P.nGB) + (local & 1) : local / P.nGB;
    const UmmaBlockInfo& BI = P.blk[gbT];
    const long long A0 = S.n0 / P.q + (long long) pb * kRows;
    UmmaTileRec R;
    R.in = S.in; R.out = S.out; R.inAvail = S.inAvail; R.numOut = S.numOut;
    R.l00 = A0 * P.p + BI.U0 - S.inOffset;
    R.oBase = A0 * P.q + BI.slot0 - S.n0;
    // Through TMA if every box of the tile lies inside the segment's window, or at least inside the allocation around it (then
    // the converters zero what lies outside the window); never a read outside memory the caller owns, never reliance on
    // out-of-bounds fill.
    const bool interior = R.l00 >= 0 && R.l00 + (long long) (kRows - 1) * P.p + (long long) BI.nStages * 32 <= S.inAvail;
    const unsigned long long aLo = (unsigned long long) reinterpret_cast<uintptr_t>(S.in + R.l00);
    const unsigned long long aHi = aLo + 4ull * (unsigned long long) ((kRows - 1) * P.p + BI.nStages * 32);
    bool contained = false;
    for (int r = 0; r < TM.nRanges; ++r) contained |= aLo >= TM.rangeLo[r] && aHi <= TM.rangeHi[r];
    const unsigned long long rel = aLo - TM.base0;
    R.x0 = (int) ((rel & 0xffffffffull) >> 2);
    R.mapIdx = (interior || contained) && aLo >= TM.base0 && (rel >> 32) < (unsigned long long) TM.nMaps ? (int) (rel >> 32) : -1;
    R.mask = interior ? 0 : 1; R.pad = 0;
    recs[tileId] = R;
}

// fp32 recomputation of a launch whose input did not fit the fp16 split (same tiles, CUDA cores, no staging):
// exits at once unless the flag is set.
__global__ void __launch_bounds__(256)
umma_redo_kernel(const Seg* __restrict__ segs, const int* __restrict__ tilePrefix, int nSegs, int nTiles, int nGB, int q16,
                 PolyDev W, const unsigned* __restrict__ ovf) {
    if (*ovf == 0u) return;
    for (int blk = blockIdx.x; blk < nTiles / nGB; blk += gridDim.x) {       // one pass per period block covers all slots
        const int tileId = blk * nGB;                          // a segment owns a multiple of nGB tiles: block index -> segment
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

// ---- tensor maps (host) ---------------------------------------------------------------------------------------
// The driver entry point is fetched through the runtime (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PointerAttrFn)(void*, CUpointer_attribute, CUdeviceptr);
bool umma_encode_maps(const Seg* segs, int n, int p, UmmaTma* out, bool noRanges) {
    // process-wide, resolved once (one host thread per GPU may plan concurrently: call_once orders the writes before every read)
    static EncodeTiledFn encode = nullptr; static PointerAttrFn pattr = nullptr; static std::once_flag once;
    std::call_once(once, [] {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiledFn>(fn);
        else (void) cudaGetLastError();
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            pattr = reinterpret_cast<PointerAttrFn>(fn);
        else (void) cudaGetLastError();
    });
    if (!encode || (p & 3) != 0 || p <= 0) return false;
    std::vector<std::pair<unsigned long long, unsigned long long>> wins;
    for (int i = 0; i < n; ++i) {
        if (segs[i].numOut <= 0 || segs[i].inAvail <= 0) continue;
        const unsigned long long a = (unsigned long long) reinterpret_cast<uintptr_t>(segs[i].in);
        wins.emplace_back(a, a + 4ull * (unsigned long long) segs[i].inAvail);
    }
    if (wins.empty()) return false;
    std::sort(wins.begin(), wins.end());
    unsigned long long lo = wins.front().first, hi = 0;
    for (const auto& w : wins) hi = std::max(hi, w.second);
    // allocations around the windows (usually one: an arena or a framework's pool block)
    out->nRanges = 0;
    if (pattr && !noRanges)
        for (const auto& w : wins) {
            bool known = false;
            for (int r = 0; r < out->nRanges && !known; ++r) known = w.first >= out->rangeLo[r] && w.second <= out->rangeHi[r];
            if (known || out->nRanges == kUmmaMaxMaps) continue;
            CUdeviceptr rs = 0; size_t sz = 0;
            if (pattr(&rs, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR, (CUdeviceptr) w.first) != CUDA_SUCCESS ||
                pattr(&sz, CU_POINTER_ATTRIBUTE_RANGE_SIZE, (CUdeviceptr) w.first) != CUDA_SUCCESS || sz == 0) continue;
            out->rangeLo[out->nRanges] = (unsigned long long) rs; out->rangeHi[out->nRanges] = (unsigned long long) rs + sz; ++out->nRanges;
        }
    // the first tile of a segment starts a window's worth of taps before it, the last one ends up to a tile after it
    unsigned long long base0 = lo > (1ull << 16) ? lo - (1ull << 16) : 0, top = hi + (1ull << 22);
    base0 &= ~15ull;
    const int nm = (int) (((top - base0) >> 32) + 1);
    if (nm > kUmmaMaxMaps) return false;
    // Rows p floats apart, 32 floats per box row: the rows overlap in memory, which the descriptor does not mind.  The extents
    // are only bounds for the coordinates the kernel uses (x < 2^30 + a tile's K span, y < 128); out-of-bounds fill is never
    // relied on: what a box reads outside its segment's window is real memory of the same allocation, zeroed by the converters.
    const cuuint64_t dims[2] = {(cuuint64_t) ((1ull << 30) + 65536), 1024};
    const cuuint64_t strides[1] = {(cuuint64_t) p * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t) kRows}, es[2] = {1, 1};
    for (int k = 0; k < nm; ++k) {
        const CUresult r = encode(&out->maps[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, reinterpret_cast<void*>(base0 + ((unsigned long long) k << 32)), dims, strides,
                                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    out->base0 = base0; out->nMaps = nm;
    return true;
}

cudaError_t launch_umma(const ResampleLaunch& L, cudaStream_t s, long long* launches) {
    // the opt-in shared-memory size is a per-device function attribute (one context per GPU may share this process)
    static std::atomic<unsigned long long> attr_done_mask{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const bool attr_done = dev < 64 && ((attr_done_mask.load(std::memory_order_acquire) >> dev) & 1ull);
    if (!attr_done) {                                          // idempotent: two threads of one device may both set it
        cudaError_t e = cudaFuncSetAttribute(umma_fir_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_fir_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_fir_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_fir_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_fir_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_done_mask.fetch_or(1ull << dev, std::memory_order_release);
    }
    int grid = std::min(L.n_tiles, std::max(L.sm_count, L.um.nGB));
    grid -= grid % L.um.nGB;
    if (grid <= 0) return cudaErrorInvalidValue;
    const bool pairs = L.um_tma && L.um_cta2;
    if (pairs) {                                               // whole pairs, a multiple of 2 * nGB; the table has a padding record for an odd last tile
        const int unit = 2 * L.um.nGB;
        grid = std::min((grid + unit - 1) / unit * unit, L.sm_count / unit * unit);
        if (grid <= 0) return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaMemsetAsync(L.d_ovf, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
#ifdef F9_DIAG
    // -DF9_DIAG builds only (single-threaded development runs): F9_UMMA_PROF=k: per-role cycle accounting of launch k, printed to
    // stderr; F9_UMMA_DBG: 1 skip MMAs, 2 skip copies, 4 skip stores
    static long long* d_prof = nullptr; static int prof_calls = 0;
    const char* profEnv = getenv("F9_UMMA_PROF");
    const bool doProf = profEnv != nullptr && prof_calls++ == atoi(profEnv);
    if (doProf) { cudaMalloc((void**) &d_prof, sizeof(long long) * 16 * grid); cudaMemsetAsync(d_prof, 0, sizeof(long long) * 16 * grid, s); }
    const int dbg = getenv("F9_UMMA_DBG") ? atoi(getenv("F9_UMMA_DBG")) : 0;
    static long long* d_trace = nullptr; static int trace_calls = 0;
    const char* trEnv = getenv("F9_UMMA_TRACE");
    const bool doTrace = trEnv != nullptr && L.um_tma && trace_calls++ == atoi(trEnv);
    const size_t trWords = (size_t) 2 * kTrWarps * kTrCap;
    if (doTrace) { cudaMalloc((void**) &d_trace, sizeof(long long) * trWords); cudaMemsetAsync(d_trace, 0, sizeof(long long) * trWords, s); }
#else
    long long* const d_trace = nullptr; const bool doTrace = false;
    long long* const d_prof = nullptr; const bool doProf = false; const int dbg = 0;
#endif
    #define F9_UMMA_LAUNCH(MERGED, TMA) umma_fir_kernel<MERGED, TMA, false><<<grid, TMA ? kThreadsTma : kThreads, L.um_smem, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, \
        L.um, L.um_maps, L.d_tile_recs, L.um_stages, L.um_aligned ? 1 : 0, L.d_ovf, doTrace ? d_trace : (doProf ? d_prof : nullptr), dbg)
    if (L.um_tma) {
        if (!L.d_tile_recs) return cudaErrorInvalidValue;
        if (!L.recs_ready) {
            umma_tile_table_kernel<<<(L.n_tiles + 256) / 256, 256, 0, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.um, L.um_maps, L.d_tile_recs,
                                                                               pairs && L.um.nGB > 1 ? 1 : 0);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            ++*launches;
        }
        if (pairs) {                                             // clusters of two CTAs
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned) grid); cfg.blockDim = dim3(kThreadsTma); cfg.dynamicSmemBytes = L.um_smem; cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (doProf) {
                int nc = -1; cudaError_t qe = cudaOccupancyMaxActiveClusters(&nc, umma_fir_kernel<false, true, true>, &cfg);
                fprintf(stderr, "[umma pairs] grid %d smem %zu stages %d max active clusters %d (%s)\n", grid, L.um_smem, L.um_stages, nc, cudaGetErrorString(qe));
            }
            e = cudaLaunchKernelEx(&cfg, umma_fir_kernel<false, true, true>, L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.um, L.um_maps,
                                   (const UmmaTileRec*) L.d_tile_recs, L.um_stages, L.um_aligned ? 1 : 0, L.d_ovf, doTrace ? d_trace : (long long*) nullptr, dbg);
            if (e != cudaSuccess) return e;
        }
        else if (L.um.poolN == 0) F9_UMMA_LAUNCH(true, true); else F9_UMMA_LAUNCH(false, true);
    }
    else          { if (L.um.poolN == 0) F9_UMMA_LAUNCH(true, false); else F9_UMMA_LAUNCH(false, false); }
    #undef F9_UMMA_LAUNCH
    if (doProf) {
        std::vector<long long> h((size_t) 16 * grid);
        cudaStreamSynchronize(s); cudaMemcpy(h.data(), d_prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        double a[16] = {0}; for (int i = 0; i < grid; ++i) for (int k = 0; k < 16; ++k) a[k] += (double) h[(size_t) i * 16 + k] / grid;
        if (L.um_tma)
            fprintf(stderr, "[umma prof tma] tiles/CTA %.1f stages/tile %d ring %d | converter0 total %.0f wait-full %.0f wait-slot %.0f st+arrive %.0f | producer total %.0f wait-empty %.0f | epilogue total %.0f wait-done %.0f stores %.0f | issuer0: wait-ready %.0f issue %.0f (wait-drained %.0f)\n",
                    (double) L.n_tiles / grid, L.um.blk[0].nStages, L.um_stages, a[0], a[1], a[8], a[9], a[2], a[3], a[5], a[6], a[7], a[10], a[11], a[12]);
        else
        fprintf(stderr, "[umma prof] tiles/CTA %.1f stages/tile %d | loader total %.0f wait-free %.0f | copy total %.0f wait-full %.0f wait-slot %.0f | epilogue total %.0f wait-done %.0f stores %.0f | copy: fence %.0f cp-issue %.0f | issuer0: wait-cp %.0f issue %.0f (wait-drained %.0f)\n",
                (double) L.n_tiles / grid, L.um.blk[0].nStages, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
    }
#ifdef F9_DIAG
    if (doTrace) {
        std::vector<long long> h(trWords);
        cudaStreamSynchronize(s); cudaMemcpy(h.data(), d_trace, sizeof(long long) * trWords, cudaMemcpyDeviceToHost);
        const char* path = getenv("F9_UMMA_TRACE_FILE");
        if (FILE* f = fopen(path ? path : "umma_trace.bin", "wb")) { fwrite(h.data(), sizeof(long long), trWords, f); fclose(f); }
        fprintf(stderr, "[umma trace] grid %d tiles %d stages/tile %d ring %d aSlots %d pairs %d\n", grid, L.n_tiles, L.um.blk[0].nStages, L.um_stages, L.um.aSlots, (int) pairs);
    }
#endif
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    umma_redo_kernel<<<std::min(L.n_tiles, 8 * L.sm_count), 256, 0, s>>>(L.d_segs, L.d_tile_prefix, L.n_segs, L.n_tiles, L.um.nGB, L.um.q,
                                                                          L.poly, L.d_ovf);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    ++*launches;
    return cudaSuccess;
}

}  // namespace f9
