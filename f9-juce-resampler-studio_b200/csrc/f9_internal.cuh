// Internal declarations shared by the translation units of libf9dsp.so.
// Product code: nothing here may include or link anything under oracle/.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "f9dsp.h"

namespace f9 {

// Variant switches for tests and development (f9_context_set_option): which kernel generation / feed a plan takes.  They are
// per-context values set through the API; the library never reads the environment (a -DF9_DIAG build adds the in-kernel
// ablation / cycle-accounting switches, read from F9_UMMA_DBG, F9_UMMA_PROF, F9_HK_DBG).
struct DiagOpts {
    std::map<std::string, int> m;
    bool has(const char* k) const { return m.find(k) != m.end(); }
    int  get(const char* k, int dflt = 0) const { auto it = m.find(k); return it == m.end() ? dflt : it->second; }
};

// ----------------------------------------------------------------------------- tables (host, f9_tables.cpp)
constexpr int kSincTableSize = 10001;
constexpr int kSincTaps = 200;       // WindowedSinc memory (taps i = -100 .. 99 contribute)
constexpr int kLagrangeTaps = 5;

void make_default_sinc_table(float* t);                 // sinc * Hann stand-in for JUCE's lookupTable
int  interp_memory(int kind);                           // ring size: 200, 5, 4, 2, 1
float interp_latency(int kind);
// Find p/q (q <= max_q) with (double)p/(double)q == ratio bit for bit.  Returns false if none.
bool  find_rational(double ratio, int max_q, long long* p, long long* q);
// Per-tap weights of one output at sub-sample offset `offset` (oldest input first).
// For WindowedSinc this walks JUCE's valueAtOffset index/frac logic tap by tap.
void  tap_weights(int kind, const float* sinc_table, float offset, float* w /* interp_memory(kind) */);

// Polyphase description of a rational ratio p/q: output n = a*q + k reads inputs ending at
// m = a*p + B[k] with weights W[tap][k].
struct PolyHost {
    int p = 0, q = 0, taps = 0, qpad = 0;
    std::vector<int>   B;        // q
    std::vector<float> W;        // taps * qpad, tap-major
};
void build_poly(int kind, const float* sinc_table, long long p, long long q, PolyHost* out);

// host scalars with the reference's rounding (f9_tables.cpp, compiled with -ffp-contract=off)
float largest_rms_below(float thrDb, float floorv);
float largest_peak_below(float thrDb, int* below0);
float noise_floor_db_from_rms(float rms);
float nf_threshold_db(int has_nf, float nf_db, float margin_pct);
float threshold_linear(float db);
float sine_phase_increment(float frequency, float sample_rate);
float sine_phase_after_block(float phase, float inc, int num_samples);
int   run_position_chain(double* pos_io, double ratio, int num_out);
void  position_closed_form(double pos0, double ratio, long long n, long long* c, double* frac_out);

// ----------------------------------------------------------------------------- device-side job records
struct DevBuf {                    // planar float32 buffer
    const float* base; long long chStride; int numCh; int numFrames;
};
struct Seg {                       // one resample segment (mirror of f9_resample_seg)
    const float* in; long long inOffset; long long inAvail;
    float* out;      long long n0;       long long numOut;
};
struct TailParams {
    long long startFrame; int window, hop, required, mode;
    int noNf;           // 1: Swift fallback (peak < 0.0001f) / C++ -80 dB handled through rstar
    float rstar;        // largest value (rms or peak) for which the predicate is true; <0: never
    int below0;         // predicate value at exactly 0 (peak mode: -160 dB branch)
};

struct PeakPartial { float v; int ch; int pos; };
struct XcPartial   { double v; int ch; int lag; int pad; };

// ----------------------------------------------------------------------------- launchers
// All take the stream; none synchronise.  *launches is bumped once per kernel launched.
constexpr int kStatPartialsPerBuf = 64;     // stage-1 CTAs per buffer of launch_stats

int         peak_prefix(const DevBuf* h_bufs, int n, std::vector<int>* prefix);          // returns total CTAs
cudaError_t launch_find_peak(const DevBuf* d_bufs, int n, int total_ctas, const int* d_prefix, float threshold,
                             PeakPartial* d_partials /* total_ctas */, int* d_out_pos, cudaStream_t s, long long* launches,
                             double* d_psum = nullptr /* total_ctas: the same pass sums the squares */, double* d_sumsq = nullptr,
                             float* d_peakv = nullptr, int forceOrder = 0 /* always re-sum in the reference's order (tests) */);

cudaError_t launch_stats(const DevBuf* d_bufs, int n, double* d_psum, float* d_pmax /* n*kStatPartialsPerBuf each */,
                         double* d_sumsq, float* d_peak, cudaStream_t s, long long* launches, int forceOrder = 0);

cudaError_t launch_tail_scan(const DevBuf* d_bufs, const TailParams* d_params, int n, int max_polls,
                             long long* d_stop, int* d_flags /* n*max_polls, required */, cudaStream_t s, long long* launches);

int         xcorr_prefix(const DevBuf* h_bufs, int n, int lagMin, int lagMax, std::vector<int>* prefix);
cudaError_t launch_xcorr(const DevBuf* d_bufs, int n, int total_ctas, const int* d_prefix, const float* d_stim, int stimLen,
                         int lagMin, int lagMax, XcPartial* d_partials /* total_ctas */, XcPartial* d_best /* n */,
                         cudaStream_t s, long long* launches, const int* d_need = nullptr /* per buffer: scan it (nullptr: all) */);
// f9_xcorr.cu: approximate every lag on the tensor cores, verify the candidates exactly; results identical to launch_xcorr
size_t      xcorr_fast_scratch_bytes(int n, int maxCh, int lagMin, int lagMax);
cudaError_t launch_xcorr_fast(const DevBuf* h_bufs, const DevBuf* d_bufs, int n, int total_ctas, const int* d_prefix, const float* d_stim, int stimLen,
                              int lagMin, int lagMax, XcPartial* d_partials, XcPartial* d_best, void* d_scratch, cudaStream_t s, long long* launches);

constexpr int kDcPartials = 32;             // partial sums per (buffer, channel) of the DC passes: d_partials holds n * maxCh * kDcPartials doubles
// d_dc_partials != nullptr: removeDCOffset fused into the trim (buffers with d_dc_mask[i] == 0 are copied unchanged; nullptr = all)
cudaError_t launch_trim(const DevBuf* d_captured, const int* d_latency, const DevBuf* d_out, int n, int maxOutFrames,
                        int maxCh, cudaStream_t s, long long* launches, double* d_dc_partials = nullptr, const int* d_dc_mask = nullptr,
                        int dcDefault = 1 /* mode when there is no mask: 1 parallel double sum, 2 the reference's sequential float sum */);
cudaError_t launch_remove_dc(const DevBuf* d_bufs /* writable */, int n, int maxCh, int maxFrames, double* d_partials /* n*maxCh*kDcPartials */,
                             cudaStream_t s, long long* launches, int mode = 1);
cudaError_t launch_pcm_to_planar(const void* d_src, int fmt, int srcCh, long long frames, float* d_dst,
                                 long long dstStride, int dstCh, cudaStream_t s, long long* launches);
// h_dsts / h_srcs: host copies of the payload pointer arrays (alignment check for the 128-bit fast paths; nullptr = byte-staged kernels)
cudaError_t launch_planar_to_pcm24_batch(const DevBuf* h_srcs, const DevBuf* d_srcs, unsigned char* const* d_dsts, int n,
                                         cudaStream_t s, long long* launches, unsigned char* const* h_dsts = nullptr);
cudaError_t launch_pcm_to_planar_batch(const unsigned char* const* d_srcs, int fmt, int srcCh, const DevBuf* h_dsts, const DevBuf* d_dsts, int n,
                                       cudaStream_t s, long long* launches, const unsigned char* const* h_srcs = nullptr);
cudaError_t launch_planar_to_pcm24(const float* d_src, long long srcStride, int numCh, long long frames,
                                   unsigned char* d_dst, cudaStream_t s, long long* launches);
cudaError_t launch_interleave(const float* d_src, long long srcStride, int numCh, long long frames, float* d_dst,
                              cudaStream_t s, long long* launches);
cudaError_t launch_deinterleave(const float* d_src, int numCh, long long frames, float* d_dst, long long dstStride,
                                cudaStream_t s, long long* launches);

// stimuli (f9_stimulus.cu): generateImpulse / generateSineWave
cudaError_t launch_impulse(const DevBuf* d_bufs /* writable */, int n, int maxCh, int maxFrames, float amplitude, cudaStream_t s, long long* launches);
cudaError_t launch_sine(const DevBuf& buf /* writable */, float phase0, float inc, float amplitude, int n, float* d_phases /* n + 1 */,
                        cudaStream_t s, long long* launches);
cudaError_t launch_sine_swift(float* d_out, int channels, double phase0, double inc, float amplitude, int frames, double* d_phases /* frames + 1 */,
                              cudaStream_t s, long long* launches);

// resampling ---------------------------------------------------------------
struct PolyDev {                   // device copy of PolyHost
    int p = 0, q = 0, taps = 0, qpad = 0;
    int* B = nullptr; float* W = nullptr;
};
// Band-aligned polyphase tables for the register-tiled FIR (banded_kernel): slots are grouped TK at a time, every
// group walks one shared window of Tmax input samples and slot j of the group has its `taps` weights placed at
// offset B[k] - B[g*TK] inside it (zeros elsewhere), so one input register feeds TK accumulators.
struct BandedDev {
    int p = 0, q = 0;              // scaled ratio p/q (q multiplied up so that a group is at least 16 slots)
    int taps = 0, TK = 0, G = 0;   // slots per group, groups (G*TK >= q)
    int Tmax = 0;                  // window steps per group (same for all groups, even)
    float* C = nullptr;            // [Gpad][Tmax][TK]
    int* wmin = nullptr;           // [Gpad] first window sample relative to the period base: B[g*TK] - (taps-1)
    int Gpad = 0;
};
struct BandedHost {
    int p = 0, q = 0, taps = 0, TK = 0, G = 0, Gpad = 0, Tmax = 0;
    std::vector<float> C; std::vector<int> wmin;
};
void build_banded(int kind, const float* sinc_table, long long p, long long q, int TK, int Gpad, BandedHost* out);

// Tensor-core polyphase FIR (umma_fir_kernel, f9_umma.cu).  Outputs are indexed (period a, slot k) as above; a tile is
// 128 periods (the M rows of a tcgen05.mma) x one block of groups of NB = 16 or 32 slots (N).  K runs over the input window of a
// period in steps of 16 samples.  Samples and weights are split into an fp16 head and an fp16 tail scaled by 2^11
// (x = x0 + x1/2048, w = w0 + w1/2048), three products are kept:  D0 += x0*w0,  D1 += x0*w1 + x1*w0,  out = D0 + D1/2048.
constexpr int kUmmaMaxBlocks = 8;        // group blocks per ratio (passes over the same rows)
constexpr int kUmmaMaxGroups = 12;       // groups per block: 2*NB TMEM columns each
constexpr int kUmmaMaxNK = 64;           // K steps per block (period + taps + alignment <= 1024 input samples)
constexpr int kUmmaIssuers = 3;          // MMA-issuing warps of the register-loader kernel: group gl of a block belongs to warp gl % 3
constexpr int kUmmaIssuersTma = 5;       // ... of the TMA-fed kernels (a single thread's issue latency bounds the role: fewer groups per warp)
// One (group, K step) of a tile as the issuing warp sees it: 16 bytes, listed per warp in issue order (stage, group, K step).
// The lists are built on the host (build_umma), live behind the weight tiles and are copied to shared memory with them.
struct UmmaOp {
    uint16_t d1Col;        // TMEM column of the group's D1 accumulator ((2 gl + 1) NB); D0 is NB columns below
    uint16_t poolCol;      // TMEM column of the group's D0B pool slot (split plans)
    uint32_t bOff;         // weight tile offset from the first tile, in 16-byte units (shared-memory descriptor address field)
    uint8_t  stage, h;     // stage of the tile, K step of the stage (0 / 1)
    uint8_t  gl;           // group inside the block
    uint8_t  flags;        // kOp* below
    uint8_t  waitGl;       // kOpWaitPool: the group whose drained accumulators free the pool slot
    uint8_t  pad[3];
};
static_assert(sizeof(UmmaOp) == 16, "UmmaOp is loaded as one 128-bit word");
enum : uint8_t { kOpAcc = 1,          // accumulate into D0/D1 (not the group's first K step)
                 kOpMerged = 2,       // [D0 | D1] (+)= x0 [w0 | w1] as one N = 2 NB MMA (before the split, or no split)
                 kOpPoolAcc = 4,      // past the split: accumulate into the pool slot (not its first step there)
                 kOpWaitDrain = 8,    // first K step: the epilogue must have drained the group's accumulators (previous tile)
                 kOpWaitPool = 16,    // first step past the split: the pool slot's previous user must have been drained
                 kOpPoolPrevTile = 32,// ... and that user belongs to the previous tile
                 kOpLast = 64 };      // last K step: commit "group done"
struct UmmaBlockInfo {
    int U0;            // K index 0 is input sample a*p + U0 (multiple of 16, <= every window start of the block)
    int nK;            // K steps (16 samples each) the block's windows span
    int nStages;       // 32-sample stages staged per tile: covers the last window sample + 3 (loader funnel shift)
    int nGroups;       // groups in this block
    int slot0;         // first slot of the block
    int nEntries;      // weight tiles of the block (one per (group, K step of its window)), NB*64 bytes each from wOff
    int wOff;          // byte offset into W
    int opOff;         // byte offset into W of the block's UmmaOp lists (nEntries records, warp 0's first), 3 issuing warps
    int opStart[kUmmaIssuers + 1];   // record range of each issuing warp
    int opOffT;        // the same for the TMA-fed kernels' 5 issuing warps
    int opStartT[kUmmaIssuersTma + 1];
    int w2Off[2];      // CTA-pair kernel (NB = 32, one block): byte offset into W of the weight halves of cluster rank 0 / 1:
                       // per tile [2 K chunks][16 x w0, 16 x w1*2048 of slots 16*rank ..][8] = 1 KB; -1: not built
};
struct UmmaHost {
    int p = 0, q = 0, taps = 0, NB = 16, G = 0, GBL = 0, nGB = 0;
    int maxEntries = 0, maxNK = 0;
    UmmaBlockInfo blk[kUmmaMaxBlocks] = {};
    // MMA schedule, per group of a block: first K step, number of K steps, index of its first weight tile (tiles are stored
    // group-major so an issuing warp walks a group with constant operand increments).
    uint8_t gStart[kUmmaMaxBlocks][kUmmaMaxGroups] = {}, gSteps[kUmmaMaxBlocks][kUmmaMaxGroups] = {};
    uint16_t gTile[kUmmaMaxBlocks][kUmmaMaxGroups] = {};
    // Accumulation split (poolN > 0): the tensor core truncates the fp32 accumulator after every MMA, so x0*w0 of the K
    // steps after the window's centre (small partial sums) goes to a second accumulator D0B taken from a pool of poolN
    // 16-column slots (group g uses slot g % poolN); the epilogue adds D0A + D0B.  split = first K step of the second part.
    int poolN = 0, split = 0;
    int aSlots = 2;                      // stages of the TMEM operand ring (2 or 4): columns 512 - 32*aSlots .. 511
    std::vector<uint8_t> W;              // fp16 weight tiles [2 K chunks][32 rows: 16 x w0, 16 x w1*2048][8], schedule order
};
struct UmmaDev {
    int p = 0, q = 0, taps = 0, NB = 16, G = 0, GBL = 0, nGB = 0;
    int maxEntries = 0, maxNK = 0;
    UmmaBlockInfo blk[kUmmaMaxBlocks] = {};
    uint8_t gStart[kUmmaMaxBlocks][kUmmaMaxGroups] = {}, gSteps[kUmmaMaxBlocks][kUmmaMaxGroups] = {};
    uint16_t gTile[kUmmaMaxBlocks][kUmmaMaxGroups] = {};
    int poolN = 0, split = 0;
    int aSlots = 2;
    const uint8_t* W = nullptr;
};
// shift (0 .. 3): the plan for segments whose first sample sits `shift` floats past a 16-byte boundary (K origins = -shift mod 16)
bool build_umma(int kind, const float* sinc_table, long long p, long long q, int NB, int GBL, UmmaHost* out, int shift = 0);
size_t umma_smem_bytes(int maxEntries, int NB, int stages, bool tma = false, bool cta2 = false);
// TMA feed of the tensor-core FIR: the input rows of a tile (128 periods, p floats apart, 32 samples per stage) are one box of a
// rank-2 tensor map whose row stride is p floats, i.e. the rows overlap in memory.  One map covers 4 GiB from its base (box
// coordinates are 32-bit element indices); a launch carries up to kUmmaMaxMaps maps, 4 GiB apart, from the lowest input address.
constexpr int kUmmaMaxMaps = 8;
struct alignas(64) UmmaTma {
    CUtensorMap maps[kUmmaMaxMaps];
    unsigned long long base0 = 0;      // device address of maps[0]'s origin (16-byte aligned)
    int nMaps = 0;
    // Device allocations the segments live in (cuPointerGetAttribute): a tile at the start or the end of a segment, whose boxes
    // leave the segment's window, still goes through TMA when the boxes stay inside one of these (the converters zero what lies
    // outside the window); otherwise its rows are read with guarded loads.
    int nRanges = 0;
    unsigned long long rangeLo[kUmmaMaxMaps] = {}, rangeHi[kUmmaMaxMaps] = {};
};
// Encodes the maps for rows p floats apart over the address range the segments read and finds the allocations around them.
// False when the driver entry point is missing or refuses the map (the register loader is used instead).
// Per-tile record of the TMA-fed kernel, written by umma_tile_table_kernel before the FIR launch so that no role searches the
// segment table on its critical path (a role loads the record of its next tile one tile ahead).
struct alignas(16) UmmaTileRec {
    const float* in; float* out;       // segment window origin, segment output
    long long l00, inAvail;            // window index of (row 0, K 0); window length
    long long oBase, numOut;           // output index of (row 0, slot 0 of the block), relative to the segment; segment outputs
    int x0, mapIdx;                    // box coordinate / tensor map of stage 0; mapIdx < 0: the tile does not go through TMA
    int mask, pad;                     // mask: the boxes leave the window [0, inAvail): zero what lies outside after the load
};
bool umma_encode_maps(const Seg* segs, int n, int p, UmmaTma* out, bool noRanges = false);
void umma_choose_plan(int taps, long long p, long long q, long long* m_out, int* NB_out, int* GBL_out, int forceNB = 0);
double umma_cost_per_output(int taps, long long p, long long q, int NB, int GBL, size_t* smem2);   // model used to pick the plan

// Hankel-operand FIR for integer upsampling 1:L (f9_hankel.cu): device weight image + geometry
struct HankelDev {
    int L = 0;                     // upsampling factor (2, 4, 8, 16); R = 128 / L inputs per column of 128 outputs
    int KS = 0;                    // K steps of 16 samples: K = R + 208 rounded up
    int rowBytes = 0, layout = 0;  // operand row pitch 2R bytes = swizzle width; descriptor layout code (0 none, 6 / 4 / 2 = 32 / 64 / 128-byte swizzle)
    int elems = 0, bufBytes = 0;   // fp16 elements per tile buffer (R * 64 + 16 * KS); buffer size rounded up to 1024 bytes
    int cLo = 0, cHi = 0;          // K steps that hold the main lobe of some lane's filter (|w| up to 1): separate accumulator
    const uint8_t* W = nullptr;    // [head, tail][KS][2 chunks][128 rows][8 fp16]: A[(i,k), t] = w_k[t - (209 - taps) - i]
};
constexpr size_t kHankelTileRecBytes = 48;  // sizeof(HankelTileRec), f9_hankel.cu
bool build_hankel(int kind, const float* sinc_table, int L, std::vector<uint8_t>* image, int* KS_out);
size_t hankel_smem_bytes(const HankelDev& P);
long long hankel_tiles_for_segment(long long n0, long long numOut);
int hankel_tile_elems(int L, int KS);

struct ResampleLaunch {
    int kind = 0;
    const DiagOpts* diag = nullptr;   // the context's variant switches (never null after prepare_resample)
    f9_context* ctx = nullptr;        // the planning context (resample_build_tiles fetches address-shifted tables through it)
    bool hankel = false;           // integer upsampling on the Hankel-operand kernel (takes priority)
    HankelDev hk;
    // tensor-core path (takes priority over banded when set; never used with adding)
    bool umma = false;
    UmmaDev um;
    int um_stages = 0; size_t um_smem = 0;
    bool um_aligned = false;       // every row piece of every segment starts on 16 bytes (set by resample_build_tiles)
    bool um_tma = false;           // aligned and the tensor maps were encoded: TMA feed (set by resample_build_tiles)
    bool um_cta2 = false;          // ... and the plan suits CTA pairs (tcgen05.mma.cta_group::2: each CTA holds half of every weight tile)
    UmmaTma um_maps;
    UmmaTileRec* d_tile_recs = nullptr;    // n_tiles records, caller-provided scratch when um_tma (see resample_scratch_bytes)
    bool recs_ready = false;       // the tile records in d_tile_recs are those of this launch's segments already (a plan's second run): skip the table kernel
    cudaStream_t recs_stream = nullptr;   // ... written on this stream (another stream rebuilds them: no ordering between the two)
    unsigned* d_ovf = nullptr;     // device flag: an input sample was outside the fp16 split's range -> fp32 redo.  One per launch
                                   // (the caller carves it next to the segment table: arena for transient launches, the plan's own
                                   // allocation for plans), so launches on different streams never share a flag
    // banded (register-tiled) path
    bool banded = false;
    BandedDev band;
    int TA = 1, PB = 1, GB = 1, nGB = 1, halo = 0, Pstride = 0;
    size_t banded_smem = 0; int sm_count = 148;
    double ratio = 1.0;
    double pos0 = 1.0;             // sub-sample position before output 0 (1.0 = reset state)
    bool rational = false;
    PolyDev poly;                  // valid when rational
    // short kinds (<= 5 taps) at a rational ratio: CUDA-core kernel, slot weights in registers (short_kernel)
    int short_S = 0;               // thread stride in outputs (a multiple of q); 0 = not this path
    int short_threads = 0; size_t short_smem = 0; int short_stage_floats = 0, short_stages = 3;
    const float* d_sinc_table = nullptr;   // generic WindowedSinc path
    const Seg* d_segs = nullptr;
    const int* d_tile_prefix = nullptr;    // n_segs + 1
    int n_segs = 0;
    int n_tiles = 0;
    int tile_out = 0;              // outputs per tile
    int adding = 0; float gain = 1.0f;
};
int         choose_tile_out(double ratio);
cudaError_t launch_resample(const ResampleLaunch& L, cudaStream_t s, long long* launches);
cudaError_t launch_umma(const ResampleLaunch& L, cudaStream_t s, long long* launches);     // f9_umma.cu
cudaError_t launch_hankel(const ResampleLaunch& L, cudaStream_t s, long long* launches);   // f9_hankel.cu
// CTAs one segment needs under launch configuration L (tile_out outputs each, or period blocks x group blocks)
long long   resample_ctas_for_segment(const ResampleLaunch& L, long long n0, long long numOut);
// Fill tile_prefix (n+1 ints) for the segments; returns the CTA total or -1 on overflow.
int         resample_build_tiles(ResampleLaunch& L, const Seg* segs, int n, std::vector<int>* prefix);
// Device scratch the launch needs next to the segment table (set L.d_tile_recs to a buffer of this size; 0 = none)
constexpr size_t kShortTileRecBytes = 48;   // sizeof(ShortTileRec), f9_resample.cu
inline size_t resample_scratch_bytes(const ResampleLaunch& L, int n_tiles) {
    if (L.hankel) return kHankelTileRecBytes * ((size_t) n_tiles + 1);
    if (L.short_S > 0) return kShortTileRecBytes * ((size_t) n_tiles + 1);
    return L.umma && L.um_tma ? sizeof(UmmaTileRec) * ((size_t) n_tiles + 1) : 0;
}
inline bool resample_needs_ovf(const ResampleLaunch& L) { return L.hankel || L.umma; }

}  // namespace f9

// ----------------------------------------------------------------------------- context
struct f9_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;
    f9::DiagOpts diag;                  // f9_context_set_option

    // bump arenas, reset at the start of every blocking call
    char* d_arena = nullptr;  size_t d_cap = 0, d_used = 0;
    char* h_arena = nullptr;  size_t h_cap = 0, h_used = 0;     // pinned

    std::vector<float> sinc_table;      // host copy (10001 + 1 guard)
    float* d_sinc_table = nullptr;
    unsigned sinc_epoch = 0;            // bumped by f9_sinc_table_set

    struct PolyKey { int kind; long long p, q; unsigned epoch; bool operator<(const PolyKey& o) const {
        return std::tie(kind, p, q, epoch) < std::tie(o.kind, o.p, o.q, o.epoch); } };
    std::map<PolyKey, f9::PolyDev> poly_cache;
    struct BandKey { int kind; long long p, q; int TK, Gpad; unsigned epoch; bool operator<(const BandKey& o) const {
        return std::tie(kind, p, q, TK, Gpad, epoch) < std::tie(o.kind, o.p, o.q, o.TK, o.Gpad, o.epoch); } };
    std::map<BandKey, f9::BandedDev> band_cache;
    int   get_banded(int kind, long long p, long long q, int TK, int Gpad, f9::BandedDev* out);
    struct UmmaKey { int kind; long long p, q; int NB, GBL; unsigned epoch; int shift; bool operator<(const UmmaKey& o) const {
        return std::tie(kind, p, q, NB, GBL, epoch, shift) < std::tie(o.kind, o.p, o.q, o.NB, o.GBL, o.epoch, o.shift); } };
    std::map<UmmaKey, f9::UmmaDev> umma_cache;
    std::map<std::tuple<int, int, unsigned>, f9::HankelDev> hankel_cache;      // (kind, L, sinc epoch)
    int   get_hankel(int kind, int L, f9::HankelDev* out);
    int   get_umma(int kind, long long p, long long q, int NB, int GBL, f9::UmmaDev* out, int shift = 0);
    // Choose kernel + tables for (kind, ratio, pos0); fills everything in L except the segment table.
    int   prepare_resample(int kind, double ratio, double pos0, bool allow_rational, f9::ResampleLaunch* L);

    int fail(int code, const char* msg) { err = msg; return code; }
    int fail_cuda(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e); return F9_ERR_CUDA;
    }
    bool  quiescent = true;             // nothing enqueued by this context can still touch the arenas
    // f9_process_batch pipelines chunks over two slots (arena + stream each) so that the upload of chunk k+1 overlaps the
    // kernels and the download of chunk k; swap_slot() exchanges the current arena / stream with the parked one.
    struct ParkedSlot { char* d_arena = nullptr; size_t d_cap = 0, d_used = 0; char* h_arena = nullptr; size_t h_cap = 0, h_used = 0;
                        cudaStream_t stream = nullptr; bool quiescent = true; std::vector<void*> d_spill, h_spill; } parked;
    cudaStream_t alt_stream = nullptr;  // owned; the parked slot's stream = f9_process_batch's upload stream
    cudaStream_t down_stream = nullptr; // owned; f9_process_batch's download stream
    size_t free_mem_seen = 0;           // free device memory + own arenas at f9_process_batch's first call (never asked again: see there)
    std::vector<cudaEvent_t> ev_pool;   // f9_process_batch's events (three per chunk), created on demand, reused by every call
    int   cur_slot = 0;
    void  swap_slot() {
        std::swap(d_arena, parked.d_arena); std::swap(d_cap, parked.d_cap); std::swap(d_used, parked.d_used);
        std::swap(h_arena, parked.h_arena); std::swap(h_cap, parked.h_cap); std::swap(h_used, parked.h_used);
        std::swap(stream, parked.stream); std::swap(quiescent, parked.quiescent); cur_slot ^= 1;
        d_spill.swap(parked.d_spill); h_spill.swap(parked.h_spill);
    }
    // A request the reserve estimate did not cover never writes past the arena: it is served by its own allocation (slow, counted
    // in arena_spills for the tests) that lives until the arena is next reset, i.e. until the work that uses it has been waited for.
    std::vector<void*> d_spill, h_spill; long long arena_spills = 0;
    void  arena_reset() {
        d_used = 0; h_used = 0;
        for (void* p : d_spill) cudaFree(p);
        for (void* p : h_spill) cudaFreeHost(p);
        d_spill.clear(); h_spill.clear();
    }
    int   arena_reserve(size_t d_bytes, size_t h_bytes, bool async_call = false);
    void* d_alloc(size_t bytes) {
        const size_t o = (d_used + 255) & ~size_t(255);
        if (o + bytes <= d_cap) { d_used = o + bytes; return d_arena + o; }
        void* p = nullptr; ++arena_spills;
        if (cudaMalloc(&p, bytes > 0 ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        d_spill.push_back(p); return p;
    }
    void* h_alloc(size_t bytes) {
        const size_t o = (h_used + 255) & ~size_t(255);
        if (o + bytes <= h_cap) { h_used = o + bytes; return h_arena + o; }
        void* p = nullptr; ++arena_spills;
        if (cudaHostAlloc(&p, bytes > 0 ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        h_spill.push_back(p); return p;
    }
    int   get_poly(int kind, long long p, long long q, f9::PolyDev* out);
};

// A job of f9_process_batch restricted to a time segment of its conversion (f9_multi.cpp: long files split over GPUs): the job's
// `captured` window starts at sample in_offset of the trimmed channel (samples outside the window read as zero, as everywhere),
// it produces outputs [n0, n0 + num_out) of the conversion into out[c][0 ..) / out_pcm24[0 ..).  num_out == 0: an ordinary job.
// tail_only: the job only runs its reverb-tail scan (no trim, no conversion, no outputs).
struct f9_job_ext { long long n0 = 0, num_out = 0, in_offset = 0; int tail_only = 0; };
int f9_process_batch_ext(f9_context* ctx, const f9_job* jobs, const f9_job_ext* ext /* may be null */, int n_jobs, f9_result* results);

#define F9_TRY_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return (ctx)->fail_cuda(e__, #call); } while (0)
// end of a blocking entry point: wait for the stream, after which the arenas may be reused from offset 0
#define F9_FINISH(ctx) do { cudaError_t e__ = cudaStreamSynchronize((ctx)->stream); \
    if (e__ != cudaSuccess) return (ctx)->fail_cuda(e__, "cudaStreamSynchronize"); (ctx)->quiescent = true; } while (0)
