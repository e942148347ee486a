// f9_process_batch: the MainComponent/AppState batch job flow (Source/MainComponent.cpp:705-805; Swift
// processFiles AudioProcessingService.swift:66-113, :339-536) for a list of captured recordings:
//     capture -> [reverb-tail scan] -> trimLatency -> [removeDCOffset] -> [sample-rate conversion] -> float / 24-bit PCM
// One H2D pass per capture, batched kernels over the whole chunk, one D2H pass per output.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>

#include "f9_internal.cuh"

using namespace f9;

namespace {

inline long long pad64(long long frames) { return (std::max<long long>(frames, 1) + 63) / 64 * 64; }

struct JobPlan {
    f9_job_ext ext;
    int latency_frames = 0, start = 0, copied = 0, out_frames = 0;
    bool convert = false;
    double ratio = 1.0;
    DevBuf cap{}, trimmed{}, out{};
    unsigned char* d_pcm = nullptr;
    size_t bytes = 0;
    int polls = 0;
};

inline int pcm_bps(int fmt) { return fmt == F9_PCM_U8 ? 1 : fmt == F9_PCM_S16LE ? 2 : fmt == F9_PCM_S24LE ? 3 : 4; }

int validate(const f9_job& j, const f9_job_ext& x) {
    if (j.numCh <= 0 || j.captured_frames < 0 || j.original_length < 0) return F9_ERR_INVALID;
    if (j.src_pcm) { if (j.src_fmt < F9_PCM_U8 || j.src_fmt > F9_PCM_F32LE || j.src_ch <= 0) return F9_ERR_INVALID; }
    else {
        if (!j.captured) return F9_ERR_INVALID;
        for (int c = 0; c < j.numCh; ++c) if (j.captured_frames > 0 && !j.captured[c]) return F9_ERR_INVALID;
    }
    if (!x.tail_only && !j.out && !((j.flags & F9_JOB_PCM24) && j.out_pcm24)) return F9_ERR_INVALID;        // a job must deliver something
    if (x.num_out < 0 || x.n0 < 0 || (x.num_out > 0 && (j.fs_in == j.fs_out || (j.flags & (F9_JOB_REMOVE_DC | F9_JOB_TAIL_SCAN))))) return F9_ERR_INVALID;
    if (!(j.fs_in > 0.0) || !(j.fs_out > 0.0)) return F9_ERR_INVALID;
    if (interp_memory(j.interp_kind) == 0) return F9_ERR_INVALID;
    if ((j.flags & F9_JOB_TAIL_SCAN) && (j.tail_window <= 0 || j.tail_hop <= 0 || j.tail_required <= 0 ||
        (j.tail_mode != F9_TAIL_RMS && j.tail_mode != F9_TAIL_PEAK))) return F9_ERR_INVALID;
    if ((j.flags & F9_JOB_PCM24) && !j.out_pcm24) return F9_ERR_INVALID;
    return F9_OK;
}

// A chunk whose work has been enqueued on its slot's stream; its scalar results are copied out once the streams have been waited for.
struct PendingChunk { std::vector<int> idx; std::vector<int> tailJobs; long long* h_stop = nullptr; };

// Arena bytes of one chunk: device (reused by the slot's next chunk, in stream order) and pinned host staging (descriptor arrays
// and scalar results: NOT reused inside a call, the copies that read or write them may still be queued).
void chunk_bytes(const f9_job* jobs, const std::vector<int>& idx, const std::vector<JobPlan>& plans, size_t* d_out, size_t* h_out, int* maxPollsOut, int* nTailOut) {
    const int n = (int) idx.size();
    size_t d_bytes = 1 << 20, h_bytes = 16 << 10;
    int maxPolls = 0, nTail = 0;
    for (int t = 0; t < n; ++t) {
        d_bytes += plans[(size_t) t].bytes;
        if (jobs[idx[(size_t) t]].flags & F9_JOB_TAIL_SCAN) { ++nTail; maxPolls = std::max(maxPolls, plans[(size_t) t].polls); }
    }
    size_t totalCh = 0, widest = 0;
    for (int t = 0; t < n; ++t) { totalCh += (size_t) jobs[idx[(size_t) t]].numCh; widest = std::max(widest, (size_t) jobs[idx[(size_t) t]].numCh); }
    d_bytes += (size_t) nTail * ((size_t) maxPolls * sizeof(int) + 64) + (size_t) n * 2048 + totalCh * 256;
    d_bytes += (size_t) n * widest * sizeof(double) * kDcPartials;       // DC partial sums: files x widest channel count
    d_bytes += totalCh * 1024 + d_bytes / 64;                  // per-tile records of the tensor-core resampler (64 B per >= 24 KB of output)
    h_bytes += (size_t) n * 2048 + (size_t) nTail * 16 + totalCh * 256;
    *d_out = d_bytes; *h_out = h_bytes; *maxPollsOut = maxPolls; *nTailOut = nTail;
}

// The three engines of the pipeline each get ONE stream, so that each engine's work is a single in-order queue: uploads (payloads,
// then the chunk's descriptor arrays), kernels, downloads; events carry a chunk from one to the next, and from a chunk's download to
// the upload that reuses its arena.  (The step-to-step variance of the e2e leg that this layout was first suspected of -- 33.6 ms
// or anything up to 178 ms per step of config 2, while the same bytes as 512 plain copies with event hand-offs took 32.6-34.8 ms in
// the same process, tools/copy_granularity_probe.py -- was the host's: cudaMemGetInfo at the top of every call, see there.)
struct ChunkStreams { cudaStream_t up, comp, down; cudaEvent_t evUp, evComp, evDown; };

// Enqueue one chunk on the context's current slot (arena); does not wait.  The caller has reserved the arenas.
int run_chunk(f9_context* ctx, const f9_job* jobs, const std::vector<int>& idx, std::vector<JobPlan>& plans, PendingChunk* pc, const ChunkStreams& CS) {
    const int n = (int) idx.size();
    size_t d_bytes = 0, h_bytes = 0;
    int maxPolls = 0, nTail = 0, rc = F9_OK;
    chunk_bytes(jobs, idx, plans, &d_bytes, &h_bytes, &maxPolls, &nTail);
    const cudaStream_t sUp = CS.up, s = CS.comp, sDown = CS.down;
    long long* const launches = &ctx->launches;
    struct DescCopy { void* d; const void* h; size_t bytes; };
    std::vector<DescCopy> desc;                                // descriptor arrays: uploaded behind the payloads, before any kernel
    std::vector<std::function<cudaError_t()>> work;            // the chunk's launches, in order, enqueued once the uploads are
    std::vector<DevBuf> pb; std::vector<unsigned char*> pd;    // (read by a launch below: function scope)

    // ---- upload captures, carve outputs ----
    // Captures given as file bytes (f9_job::src_pcm): the payload is uploaded as it is and deinterleaved / converted on the device,
    // one launch per (format, channel count).  The planes start padA floats past a 16-byte boundary (see below); the payload is
    // placed so that the first frame whose plane address IS aligned starts a 16-byte word of the payload too, which lets the
    // 128-bit kernels serve everything but the `lead` <= 3 frames in front of it (those go through the byte-staged kernel).
    struct PcmGroup { std::vector<const unsigned char*> src; std::vector<DevBuf> dst; };
    std::map<std::pair<int, int>, PcmGroup> pcmMain, pcmHead;
    // File payloads that follow one another in host memory (a reader that holds its files in one buffer: less than 64 bytes between
    // the end of one and the start of the next) go up as ONE copy when the device placement each of them needs (below) is the one the
    // host spacing gives: a copy costs the engine ~8 us whatever its size, 2 ms for config 2's 256 files against 29 ms of transfer.
    std::vector<unsigned char*> dSrcOf((size_t) n, nullptr);
    {
        struct Run { int t0, t1; const unsigned char* h0; const unsigned char* hEnd; size_t off0; };
        std::vector<Run> runs;
        for (int t = 0; t < n; ++t) {
            const f9_job& J = jobs[idx[(size_t) t]];
            const JobPlan& P = plans[(size_t) t];
            if (!J.src_pcm || J.captured_frames <= 0) continue;
            const int padA = (P.convert && P.start > 0) ? ((4 - (P.start & 3)) & 3) : 0;
            const size_t frameBytes = (size_t) J.src_ch * pcm_bps(J.src_fmt), bytes = frameBytes * (size_t) J.captured_frames;
            const int lead = std::min(J.captured_frames, (4 - padA) & 3);
            const size_t off = (16 - (lead * frameBytes) % 16) % 16;             // the payload's device address modulo 16
            const unsigned char* h = (const unsigned char*) J.src_pcm;
            if (!runs.empty() && runs.back().t1 == t && h >= runs.back().hEnd && (size_t) (h - runs.back().hEnd) < 64 &&
                (runs.back().off0 + (size_t) (h - runs.back().h0)) % 16 == off) { runs.back().t1 = t + 1; runs.back().hEnd = h + bytes; }
            else runs.push_back(Run{t, t + 1, h, h + bytes, off});
        }
        for (const Run& R : runs) {
            const size_t span = (size_t) (R.hEnd - R.h0);
            unsigned char* d0 = (unsigned char*) ctx->d_alloc(span + 48);
            if (!d0) return ctx->fail(F9_ERR_NOMEM, "arena");
            d0 += R.off0;
            F9_TRY_CUDA(ctx, cudaMemcpyAsync(d0, R.h0, span, cudaMemcpyHostToDevice, sUp));
            for (int t = R.t0; t < R.t1; ++t) {
                const f9_job& J = jobs[idx[(size_t) t]];
                if (J.src_pcm && J.captured_frames > 0) dSrcOf[(size_t) t] = d0 + ((const unsigned char*) J.src_pcm - R.h0);
            }
        }
    }
    for (int t = 0; t < n; ++t) {
        const f9_job& J = jobs[idx[(size_t) t]];
        JobPlan& P = plans[(size_t) t];
        // trimLatency is fused into the resampler as a pointer offset of `start` frames: place the capture so that this
        // trimmed start (not the capture's first frame) falls on a 16-byte boundary, which lets the FIR's loader use
        // aligned 128-bit loads for every row (f9_umma.cu, loader_role<true>).
        const int padA = (P.convert && P.start > 0) ? ((4 - (P.start & 3)) & 3) : 0;
        const long long cs = pad64(J.captured_frames + 4);
        float* d_cap = (float*) ctx->d_alloc(sizeof(float) * (size_t) cs * J.numCh) + padA;
        if (J.src_pcm) {
            if (J.captured_frames > 0) {
                const size_t frameBytes = (size_t) J.src_ch * pcm_bps(J.src_fmt), bytes = frameBytes * (size_t) J.captured_frames;
                const int lead = std::min(J.captured_frames, (4 - padA) & 3);
                unsigned char* d_src = dSrcOf[(size_t) t];     // placed (with the copy of its run) above
                (void) bytes;
                const auto key = std::make_pair(J.src_fmt, J.src_ch);
                if (lead > 0) { pcmHead[key].src.push_back(d_src); pcmHead[key].dst.push_back(DevBuf{d_cap, cs, J.numCh, lead}); }
                if (J.captured_frames > lead) {
                    pcmMain[key].src.push_back(d_src + lead * frameBytes);
                    pcmMain[key].dst.push_back(DevBuf{d_cap + lead, cs, J.numCh, J.captured_frames - lead});
                }
            }
        } else
        for (int c = 0; c < J.numCh; ++c)
            if (J.captured_frames > 0)
                F9_TRY_CUDA(ctx, cudaMemcpyAsync(d_cap + c * cs, J.captured[c], sizeof(float) * (size_t) J.captured_frames, cudaMemcpyHostToDevice, sUp));
        P.cap = DevBuf{d_cap, cs, J.numCh, J.captured_frames};
        const bool needTrimmed = !P.convert || (J.flags & F9_JOB_REMOVE_DC);
        if (P.convert && needTrimmed) {
            const long long ts = pad64(J.original_length);
            P.trimmed = DevBuf{(float*) ctx->d_alloc(sizeof(float) * (size_t) ts * J.numCh), ts, J.numCh, J.original_length};
        }
        const long long os = pad64(P.out_frames);
        P.out = DevBuf{(float*) ctx->d_alloc(sizeof(float) * (size_t) os * J.numCh), os, J.numCh, P.out_frames};
        if (!P.convert) P.trimmed = P.out;
        if (J.flags & F9_JOB_PCM24) P.d_pcm = (unsigned char*) ctx->d_alloc((size_t) P.out_frames * J.numCh * 3 + 16);
    }

    for (int pass = 0; pass < 2; ++pass)
        for (auto& kv : (pass ? pcmHead : pcmMain)) {
            PcmGroup& G = kv.second;
            const size_t m = G.src.size();
            const unsigned char** d_p = (const unsigned char**) ctx->d_alloc(sizeof(void*) * m); DevBuf* d_b = (DevBuf*) ctx->d_alloc(sizeof(DevBuf) * m);
            const unsigned char** h_p = (const unsigned char**) ctx->h_alloc(sizeof(void*) * m); DevBuf* h_b = (DevBuf*) ctx->h_alloc(sizeof(DevBuf) * m);
            if (!d_p || !d_b || !h_p || !h_b) return ctx->fail(F9_ERR_NOMEM, "arena");
            std::memcpy(h_p, G.src.data(), sizeof(void*) * m); std::memcpy(h_b, G.dst.data(), sizeof(DevBuf) * m);
            desc.push_back(DescCopy{d_p, h_p, sizeof(void*) * m});
            desc.push_back(DescCopy{d_b, h_b, sizeof(DevBuf) * m});
            const int fmt = kv.first.first, sch = kv.first.second;
            const DevBuf* hostDst = G.dst.data(); const unsigned char* const* hostSrc = pass ? nullptr : G.src.data();     // the maps outlive the launches
            work.push_back([=]() { return launch_pcm_to_planar_batch(d_p, fmt, sch, hostDst, d_b, (int) m, s, launches, hostSrc); });
        }

    // ---- reverb-tail scan (Swift :423-453): starts once source + latency frames are captured ----
    long long* h_stop = nullptr; std::vector<int> tailJobs;
    const long long* d_stop_out = nullptr; size_t stopBytes = 0;
    if (nTail > 0) {
        std::vector<DevBuf> tb; std::vector<TailParams> tp;
        for (int t = 0; t < n; ++t) {
            const f9_job& J = jobs[idx[(size_t) t]];
            if (!(J.flags & F9_JOB_TAIL_SCAN)) continue;
            const JobPlan& P = plans[(size_t) t];
            TailParams T{};
            T.startFrame = (long long) J.original_length + std::max(P.latency_frames, 0);
            T.window = J.tail_window; T.hop = J.tail_hop; T.required = J.tail_required; T.mode = J.tail_mode;
            if (J.tail_mode == F9_TAIL_PEAK) {
                if (!J.has_nf) { T.noNf = 1; T.rstar = -1.0f; }
                else T.rstar = largest_peak_below(J.nf_db + (J.nf_db * J.margin_pct / 100.0f), &T.below0);
            } else T.rstar = largest_rms_below(nf_threshold_db(J.has_nf, J.nf_db, J.margin_pct), 1e-10f);
            tb.push_back(P.cap); tp.push_back(T); tailJobs.push_back(t);
        }
        DevBuf* d_tb = (DevBuf*) ctx->d_alloc(sizeof(DevBuf) * tb.size());
        TailParams* d_tp = (TailParams*) ctx->d_alloc(sizeof(TailParams) * tp.size());
        DevBuf* h_tb = (DevBuf*) ctx->h_alloc(sizeof(DevBuf) * tb.size());
        TailParams* h_tp = (TailParams*) ctx->h_alloc(sizeof(TailParams) * tp.size());
        std::memcpy(h_tb, tb.data(), sizeof(DevBuf) * tb.size());
        std::memcpy(h_tp, tp.data(), sizeof(TailParams) * tp.size());
        desc.push_back(DescCopy{d_tb, h_tb, sizeof(DevBuf) * tb.size()});
        desc.push_back(DescCopy{d_tp, h_tp, sizeof(TailParams) * tp.size()});
        int* d_flags = (int*) ctx->d_alloc(sizeof(int) * (size_t) std::max(maxPolls, 1) * tb.size());
        long long* d_stop = (long long*) ctx->d_alloc(sizeof(long long) * tb.size());
        const int nTb = (int) tb.size();
        work.push_back([=]() { return launch_tail_scan(d_tb, d_tp, nTb, maxPolls, d_stop, d_flags, s, launches); });
        h_stop = (long long*) ctx->h_alloc(sizeof(long long) * tb.size());
        d_stop_out = d_stop; stopBytes = sizeof(long long) * tb.size();
    }

    // ---- trim (+DC) where a trimmed buffer is materialised ----
    {
        std::vector<DevBuf> tc, to; std::vector<int> lat, dcMask;
        int maxCh = 0, maxFrames = 0;
        for (int t = 0; t < n; ++t) {
            const f9_job& J = jobs[idx[(size_t) t]];
            const JobPlan& P = plans[(size_t) t];
            if (P.ext.tail_only || (P.convert && !(J.flags & F9_JOB_REMOVE_DC))) continue;      // fused into the resampler by pointer offset
            tc.push_back(P.cap); to.push_back(P.trimmed); lat.push_back(J.latency_samples);
            dcMask.push_back((J.flags & F9_JOB_REMOVE_DC) ? ((J.flags & F9_JOB_DC_REFERENCE_ORDER) ? 2 : 1) : 0);
            maxCh = std::max(maxCh, J.numCh); maxFrames = std::max(maxFrames, J.original_length);
        }
        if (!tc.empty()) {
            const size_t m = tc.size();
            DevBuf* d_c = (DevBuf*) ctx->d_alloc(sizeof(DevBuf) * m); DevBuf* d_o = (DevBuf*) ctx->d_alloc(sizeof(DevBuf) * m);
            int* d_l = (int*) ctx->d_alloc(sizeof(int) * m);
            DevBuf* h_c = (DevBuf*) ctx->h_alloc(sizeof(DevBuf) * m); DevBuf* h_o = (DevBuf*) ctx->h_alloc(sizeof(DevBuf) * m);
            int* h_l = (int*) ctx->h_alloc(sizeof(int) * m);
            std::memcpy(h_c, tc.data(), sizeof(DevBuf) * m); std::memcpy(h_o, to.data(), sizeof(DevBuf) * m); std::memcpy(h_l, lat.data(), sizeof(int) * m);
            desc.push_back(DescCopy{d_c, h_c, sizeof(DevBuf) * m});
            desc.push_back(DescCopy{d_o, h_o, sizeof(DevBuf) * m});
            desc.push_back(DescCopy{d_l, h_l, sizeof(int) * m});
            // removeDCOffset is fused into the trim for the files that asked for it (mask), so their trimmed buffer is written once
            bool anyDc = false;
            for (size_t i = 0; i < m; ++i) anyDc = anyDc || dcMask[i] != 0;
            double* d_part = nullptr; int* d_mask = nullptr;
            if (anyDc) {
                d_part = (double*) ctx->d_alloc(sizeof(double) * m * (size_t) maxCh * kDcPartials);
                d_mask = (int*) ctx->d_alloc(sizeof(int) * m);
                int* h_mask = (int*) ctx->h_alloc(sizeof(int) * m);
                std::memcpy(h_mask, dcMask.data(), sizeof(int) * m);
                desc.push_back(DescCopy{d_mask, h_mask, sizeof(int) * m});
            }
            work.push_back([=]() { return launch_trim(d_c, d_l, d_o, (int) m, maxFrames, maxCh, s, launches, d_part, d_mask); });
        }
    }

    // ---- sample-rate conversion, one launch per (kind, ratio) group ----
    {
        std::map<std::pair<int, double>, std::vector<Seg>> groups;
        for (int t = 0; t < n; ++t) {
            const f9_job& J = jobs[idx[(size_t) t]];
            const JobPlan& P = plans[(size_t) t];
            if (!P.convert || P.out_frames <= 0) continue;
            auto& g = groups[std::make_pair(J.interp_kind, P.ratio)];
            for (int c = 0; c < J.numCh; ++c) {
                Seg S{};
                if (J.flags & F9_JOB_REMOVE_DC) { S.in = P.trimmed.base + c * P.trimmed.chStride; S.inAvail = J.original_length; }
                else { S.in = P.cap.base + c * P.cap.chStride + std::max(P.start, 0); S.inAvail = P.copied; }
                S.inOffset = P.ext.num_out > 0 ? P.ext.in_offset : 0;
                S.out = const_cast<float*>(P.out.base) + c * P.out.chStride; S.n0 = P.ext.num_out > 0 ? P.ext.n0 : 0; S.numOut = P.out_frames;
                g.push_back(S);
            }
        }
        for (auto& kv : groups) {
            ResampleLaunch L;
            rc = ctx->prepare_resample(kv.first.first, kv.first.second, 1.0, true, &L); if (rc) return rc;
            std::vector<Seg>& segs = kv.second;
            std::vector<int> prefix;
            if (resample_build_tiles(L, segs.data(), (int) segs.size(), &prefix) < 0) return ctx->fail(F9_ERR_INVALID, "too many tiles");
            Seg* d_s = (Seg*) ctx->d_alloc(sizeof(Seg) * segs.size()); int* d_p = (int*) ctx->d_alloc(sizeof(int) * prefix.size());
            Seg* h_s = (Seg*) ctx->h_alloc(sizeof(Seg) * segs.size()); int* h_p = (int*) ctx->h_alloc(sizeof(int) * prefix.size());
            std::memcpy(h_s, segs.data(), sizeof(Seg) * segs.size()); std::memcpy(h_p, prefix.data(), sizeof(int) * prefix.size());
            desc.push_back(DescCopy{d_s, h_s, sizeof(Seg) * segs.size()});
            desc.push_back(DescCopy{d_p, h_p, sizeof(int) * prefix.size()});
            L.d_segs = d_s; L.d_tile_prefix = d_p; L.n_segs = (int) segs.size(); L.n_tiles = prefix.back();
            if (const size_t sb = resample_scratch_bytes(L, L.n_tiles)) L.d_tile_recs = (UmmaTileRec*) ctx->d_alloc(sb);
            if (resample_needs_ovf(L)) L.d_ovf = (unsigned*) ctx->d_alloc(sizeof(unsigned));
            work.push_back([=]() { return launch_resample(L, s, launches); });
        }
    }

    // ---- 24-bit payload + downloads ----
    {   // one launch packs every file that asked for the WAV payload
        for (int t = 0; t < n; ++t) {
            const f9_job& J = jobs[idx[(size_t) t]];
            const JobPlan& P = plans[(size_t) t];
            if ((J.flags & F9_JOB_PCM24) && P.out_frames > 0) { DevBuf b = P.out; b.numCh = J.numCh; b.numFrames = P.out_frames; pb.push_back(b); pd.push_back(P.d_pcm); }
        }
        if (!pb.empty()) {
            DevBuf* d_b = (DevBuf*) ctx->d_alloc(sizeof(DevBuf) * pb.size()); unsigned char** d_p = (unsigned char**) ctx->d_alloc(sizeof(void*) * pd.size());
            DevBuf* h_b = (DevBuf*) ctx->h_alloc(sizeof(DevBuf) * pb.size()); unsigned char** h_p = (unsigned char**) ctx->h_alloc(sizeof(void*) * pd.size());
            std::memcpy(h_b, pb.data(), sizeof(DevBuf) * pb.size()); std::memcpy(h_p, pd.data(), sizeof(void*) * pd.size());
            desc.push_back(DescCopy{d_b, h_b, sizeof(DevBuf) * pb.size()});
            desc.push_back(DescCopy{d_p, h_p, sizeof(void*) * pd.size()});
            const DevBuf* hostB = pb.data(); unsigned char* const* hostP = pd.data(); const int nPb = (int) pb.size();
            work.push_back([=]() { return launch_planar_to_pcm24_batch(hostB, d_b, d_p, nPb, s, launches, hostP); });
        }
    }
    // uploads done -> kernels -> downloads
    for (const DescCopy& dc : desc) F9_TRY_CUDA(ctx, cudaMemcpyAsync(dc.d, dc.h, dc.bytes, cudaMemcpyHostToDevice, sUp));
    F9_TRY_CUDA(ctx, cudaEventRecord(CS.evUp, sUp));
    F9_TRY_CUDA(ctx, cudaStreamWaitEvent(s, CS.evUp, 0));
    for (auto& fn : work) F9_TRY_CUDA(ctx, fn());
    F9_TRY_CUDA(ctx, cudaEventRecord(CS.evComp, s));
    F9_TRY_CUDA(ctx, cudaStreamWaitEvent(sDown, CS.evComp, 0));
    if (h_stop) F9_TRY_CUDA(ctx, cudaMemcpyAsync(h_stop, d_stop_out, stopBytes, cudaMemcpyDeviceToHost, sDown));
    for (int t = 0; t < n; ++t) {
        const f9_job& J = jobs[idx[(size_t) t]];
        const JobPlan& P = plans[(size_t) t];
        if ((J.flags & F9_JOB_PCM24) && P.out_frames > 0)
            F9_TRY_CUDA(ctx, cudaMemcpyAsync(J.out_pcm24, P.d_pcm, (size_t) P.out_frames * J.numCh * 3, cudaMemcpyDeviceToHost, sDown));
        if (J.out && P.out_frames > 0)
            for (int c = 0; c < J.numCh; ++c)
                F9_TRY_CUDA(ctx, cudaMemcpyAsync(J.out[c], P.out.base + c * P.out.chStride, sizeof(float) * (size_t) P.out_frames, cudaMemcpyDeviceToHost, sDown));
    }
    F9_TRY_CUDA(ctx, cudaEventRecord(CS.evDown, sDown));
    pc->idx = idx; pc->tailJobs = tailJobs; pc->h_stop = h_stop;
    return F9_OK;
}

}  // namespace

extern "C" int f9_process_batch(f9_context* ctx, const f9_job* jobs, int n_jobs, f9_result* results) {
    return f9_process_batch_ext(ctx, jobs, nullptr, n_jobs, results);
}

int f9_process_batch_ext(f9_context* ctx, const f9_job* jobs, const f9_job_ext* ext, int n_jobs, f9_result* results) {
    if (!ctx) return F9_ERR_INVALID;
    if (n_jobs < 0 || (n_jobs > 0 && (!jobs || !results))) return ctx->fail(F9_ERR_INVALID, "bad job array");
    const auto tCall = std::chrono::steady_clock::now();
    F9_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    // Free device memory, asked ONCE per context: cudaMemGetInfo goes through the resource manager, and on a shared host that call
    // sometimes took 10-90 ms (the whole enqueue is 2 ms otherwise: option F9_BATCH_TIMING) -- the step-to-step variance of the
    // e2e leg.  The figure only caps the chunk size on small devices; the arenas themselves report failure if memory runs out.
    if (ctx->free_mem_seen == 0) {
        size_t freeNow = 0, totalB = 0;
        F9_TRY_CUDA(ctx, cudaMemGetInfo(&freeNow, &totalB));
        ctx->free_mem_seen = std::max<size_t>(freeNow + ctx->d_cap + ctx->parked.d_cap, 1);
    }
    const size_t freeB = ctx->free_mem_seen;
    // Chunks are pipelined over two arenas and three streams (uploads, kernels, downloads: see ChunkStreams): chunk k + 1 uploads
    // while chunk k computes and downloads (PCIe is full duplex and the copy engines are independent), so a large batch costs about
    // max(upload, download) instead of their sum.  The whole call is ENQUEUED without waiting for anything: chunk k + 2's uploads
    // wait for the event behind chunk k's downloads before they reuse its device arena, the pinned staging of descriptors and scalar
    // results is not reused inside a call, and the host waits once, at the end (with a wait per chunk, the first version, every chunk
    // exposed the pipeline to the host thread's wake-up latency).
    // Chunk size in device memory (option F9_BATCH_CHUNK_MB overrides), measured on B200 / PCIe 5 with config 2's 256 files
    // (tools/e2e_probe.py, ms per call at 64 / 128 / 256 / 512 / 1024 MB): file bytes in and out 37.5 / 35.2 / 33.4 / 33.4 / 33.9,
    // float planes in and out 47.2 / 46.6 / 46.3 / 46.4 / 47.2.
    const size_t chunkMB = (size_t) std::max(1, ctx->diag.get("F9_BATCH_CHUNK_MB", 256));
    const size_t budget = std::min(std::max<size_t>(freeB / 4, 64u << 20), chunkMB << 20);
    // streams: kernels on the context's stream (slot 0's), uploads and downloads on two streams of the context's own
    if (ctx->cur_slot) ctx->swap_slot();
    if (!ctx->alt_stream) {
        F9_TRY_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->alt_stream, cudaStreamNonBlocking));
        ctx->parked.stream = ctx->alt_stream;
    }
    if (!ctx->down_stream) F9_TRY_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    const cudaStream_t sComp = ctx->stream, sUp = ctx->alt_stream, sDown = ctx->down_stream;
    auto event_at = [&](size_t i) -> cudaEvent_t {
        while (ctx->ev_pool.size() <= i) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            ctx->ev_pool.push_back(e);
        }
        return ctx->ev_pool[i];
    };
    {   // the upload and download streams start after whatever the caller already enqueued on the context's stream
        cudaEvent_t ev = event_at(0);
        if (!ev) return ctx->fail(F9_ERR_CUDA, "cudaEventCreate");
        F9_TRY_CUDA(ctx, cudaEventRecord(ev, sComp));
        F9_TRY_CUDA(ctx, cudaStreamWaitEvent(sUp, ev, 0));
        F9_TRY_CUDA(ctx, cudaStreamWaitEvent(sDown, ev, 0));
    }

    struct Chunk { std::vector<int> idx; std::vector<JobPlan> plans; };
    std::vector<Chunk> chunks;
    std::vector<int> idx; std::vector<JobPlan> plans; size_t used = 0;
    auto flush = [&]() {
        if (idx.empty()) return;
        chunks.push_back(Chunk{});
        chunks.back().idx.swap(idx); chunks.back().plans.swap(plans);
        idx.clear(); plans.clear(); used = 0;
    };
    int worst = F9_OK;
    for (int i = 0; i < n_jobs; ++i) {
        const f9_job& J = jobs[i];
        f9_result& R = results[i];
        R = f9_result{}; R.tail_stop_frame = -1;
        JobPlan P;
        if (ext) P.ext = ext[i];
        R.status = validate(J, P.ext);
        if (R.status) { worst = R.status; ctx->err = "invalid job"; continue; }
        // trimLatency arithmetic (Source/MainComponent.cpp:833-845)
        P.latency_frames = J.latency_samples / J.numCh;
        P.start = P.latency_frames;
        P.copied = J.original_length;
        if (P.start + P.copied > J.captured_frames) P.copied = std::max(0, J.captured_frames - P.start);
        if (P.start < 0) P.copied = 0;
        P.convert = (J.fs_in != J.fs_out);
        P.ratio = J.fs_in / J.fs_out;
        P.out_frames = P.convert ? (int) f9_resampled_length(J.original_length, J.fs_in, J.fs_out) : J.original_length;
        if (P.ext.num_out > 0) P.out_frames = (int) P.ext.num_out;       // a time segment of the conversion
        if (P.ext.tail_only) { P.out_frames = 0; P.convert = false; }
        if (J.out && J.out_capacity < P.out_frames) { R.status = F9_ERR_INVALID; worst = R.status; ctx->err = "out_capacity too small"; continue; }
        if (J.flags & F9_JOB_TAIL_SCAN) {
            const long long startFrame = (long long) J.original_length + std::max(P.latency_frames, 0);
            P.polls = (int) std::max<long long>(0, (J.captured_frames - startFrame) / J.tail_hop);
        }
        P.bytes = sizeof(float) * (size_t) J.numCh * (size_t) (pad64(J.captured_frames + 4) + pad64(P.out_frames) + pad64(J.original_length))
                  + (size_t) P.out_frames * J.numCh * 3 + 4096
                  + (J.src_pcm ? (size_t) J.captured_frames * J.src_ch * pcm_bps(J.src_fmt) + 512 : 0);
        R.latency_frames = P.latency_frames; R.trim_start = P.start; R.frames_copied = P.copied;
        R.out_frames = P.out_frames; R.tail_polls = P.polls;
        if (used + P.bytes > budget && !idx.empty()) flush();
        idx.push_back(i); plans.push_back(P); used += P.bytes;
    }
    flush();
    // Reserve per slot: the largest of its chunks in device memory, the sum of its chunks in pinned staging (waits for whatever
    // the slot's stream still has from an earlier asynchronous call, then resets the arenas).
    for (int slot = 0; slot < 2; ++slot) {
        size_t dMax = 0, hSum = 0;
        for (size_t k = (size_t) slot; k < chunks.size(); k += 2) {
            size_t d = 0, h = 0; int mp = 0, nt = 0;
            chunk_bytes(jobs, chunks[k].idx, chunks[k].plans, &d, &h, &mp, &nt);
            dMax = std::max(dMax, d); hSum += h;
        }
        int rc = (chunks.size() > (size_t) slot) ? ctx->arena_reserve(dMax, hSum) : F9_OK;
        if (rc) {
            for (size_t k = 0; k < chunks.size(); ++k) for (int i : chunks[k].idx) results[i].status = rc;
            if (ctx->cur_slot) ctx->swap_slot();
            return rc;
        }
        ctx->swap_slot();
    }
    std::vector<PendingChunk> pending(chunks.size());
    std::vector<char> enqueued(chunks.size(), 0);
    for (size_t k = 0; k < chunks.size(); ++k) {
        if ((int) (k & 1) != ctx->cur_slot) ctx->swap_slot();
        ctx->d_used = 0;                                       // the arena of chunk k - 2: reused once that chunk's download is done
        ChunkStreams CS{sUp, sComp, sDown, event_at(1 + 3 * k), event_at(2 + 3 * k), event_at(3 + 3 * k)};
        int rc = (CS.evUp && CS.evComp && CS.evDown) ? F9_OK : ctx->fail(F9_ERR_CUDA, "cudaEventCreate");
        if (!rc && k >= 2 && enqueued[k - 2] && cudaStreamWaitEvent(sUp, event_at(3 + 3 * (k - 2)), 0) != cudaSuccess) rc = ctx->fail(F9_ERR_CUDA, "cudaStreamWaitEvent");
        if (!rc) rc = run_chunk(ctx, jobs, chunks[k].idx, chunks[k].plans, &pending[k], CS);
        if (rc) {
            // a chunk that failed half-way may have left work on the streams that the next user of its arena must not overtake
            worst = rc; for (int i : chunks[k].idx) results[i].status = rc;
            cudaStreamSynchronize(sUp); cudaStreamSynchronize(sComp); cudaStreamSynchronize(sDown);
        } else enqueued[k] = 1;
    }
    const auto tEnq = std::chrono::steady_clock::now();
    {   // the one wait of the call: the three streams
        cudaError_t e = cudaStreamSynchronize(sUp);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sComp);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sDown);
        if (e != cudaSuccess) { worst = ctx->fail_cuda(e, "cudaStreamSynchronize"); for (auto& c : chunks) for (int i : c.idx) results[i].status = worst; }
        ctx->quiescent = true; ctx->parked.quiescent = true;
    }
    if (ctx->cur_slot) ctx->swap_slot();
    if (ctx->diag.has("F9_BATCH_TIMING")) {                    // development: where a call's time goes (host enqueue against the wait)
        const auto tEnd = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[f9_process_batch] %zu chunks: enqueue %.2f ms, wait %.2f ms\n", chunks.size(),
                     std::chrono::duration<double, std::milli>(tEnq - tCall).count(), std::chrono::duration<double, std::milli>(tEnd - tEnq).count());
    }
    if (worst != F9_ERR_CUDA)
        for (size_t k = 0; k < chunks.size(); ++k)
            if (enqueued[k])
                for (size_t i = 0; i < pending[k].tailJobs.size(); ++i)
                    results[pending[k].idx[(size_t) pending[k].tailJobs[i]]].tail_stop_frame = pending[k].h_stop[i];
    return worst;
}
