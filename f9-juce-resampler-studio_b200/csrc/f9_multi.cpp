// Multi-GPU form of the batch job flow (the reference's loop is one process over AppState.files,
// Source/MainComponent.cpp:581-621, :705-805).  The path shards with no exchange step and no collective:
//   * unit of work = a file, or -- for a file that is large against a GPU's share of the batch -- a group of its channels
//     and / or a time segment of its conversion whose input window carries its own halo (f9_resample_segment_input_range:
//     199 inputs for WindowedSinc, 4 for Lagrange); the file's reverb-tail scan becomes a unit of its own;
//   * units are packed greedily by output-sample count (largest first onto the least loaded GPU);
//   * one host thread + one f9_context (stream set, arenas, table caches) per GPU; results are gathered on the host.
// Host code only: every sample is touched by the kernels behind f9_process_batch.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "f9_internal.cuh"

using namespace f9;

struct f9_multi {
    std::vector<f9_context*> ctx;
    std::string err;
};

namespace {

constexpr long long kSegAlign = 4;            // segment windows start on a multiple of 4 input samples: 16-byte aligned rows for the TMA feed

struct JobShape { int latency_frames, start, copied, out_frames; bool convert; long long cost; };

JobShape shape_of(const f9_job& J) {
    JobShape S{};
    S.latency_frames = J.numCh > 0 ? J.latency_samples / J.numCh : 0;          // Source/MainComponent.cpp:835
    S.start = S.latency_frames;
    S.copied = J.original_length;
    if (S.start + S.copied > J.captured_frames) S.copied = std::max(0, J.captured_frames - S.start);
    if (S.start < 0) S.copied = 0;
    S.convert = J.fs_in != J.fs_out;
    S.out_frames = S.convert ? (int) f9_resampled_length(J.original_length, J.fs_in, J.fs_out) : J.original_length;
    S.cost = (long long) std::max(S.out_frames, 1) * std::max(J.numCh, 1);
    return S;
}

bool job_is_plain(const f9_job& J) {
    return J.numCh > 0 && J.captured_frames >= 0 && J.original_length >= 0 && J.fs_in > 0.0 && J.fs_out > 0.0 && interp_memory(J.interp_kind) != 0 &&
           (J.src_pcm || J.captured);
}

}  // namespace

extern "C" {

// Greedy packing (longest processing time first): unit i goes to the least loaded bin; ties to the lower index.
int f9_shard_units(const long long* costs, int n, int world, int* out_bin) {
    if (n < 0 || world <= 0 || (n > 0 && (!costs || !out_bin))) return F9_ERR_INVALID;
    std::vector<int> order((size_t) n);
    for (int i = 0; i < n; ++i) order[(size_t) i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return costs[a] > costs[b]; });
    std::vector<long long> load((size_t) world, 0);
    for (int i : order) {
        int best = 0;
        for (int r = 1; r < world; ++r) if (load[(size_t) r] < load[(size_t) best]) best = r;
        out_bin[i] = best; load[(size_t) best] += costs[i];
    }
    return F9_OK;
}

// Cut the job list into units for n_devices GPUs and assign them.  Returns the number of units (<= max_units) or an error.
// A job is split only when it is large against a device's share: then by channel groups (float planes in and out only) and by time
// segments of `seg_out` outputs (0 = choose: about 1/4 of a device's share, at least 2^20 outputs).  Jobs that ask for DC removal
// are not split in time (the mean needs the whole channel), jobs that ask for the interleaved 24-bit payload or arrive as
// interleaved file bytes are not split by channel.
int f9_multi_partition(const f9_job* jobs, int n_jobs, int n_devices, long long seg_out, f9_unit* units, int max_units, int* n_units) {
    if (n_jobs < 0 || n_devices <= 0 || max_units < 0 || !n_units || (n_jobs > 0 && !jobs) || (max_units > 0 && !units)) return F9_ERR_INVALID;
    std::vector<f9_unit> U;
    long long total = 0;
    std::vector<JobShape> shapes((size_t) n_jobs);
    for (int i = 0; i < n_jobs; ++i) { shapes[(size_t) i] = job_is_plain(jobs[i]) ? shape_of(jobs[i]) : JobShape{}; total += shapes[(size_t) i].cost; }
    const long long share = std::max<long long>(1, total / n_devices);
    for (int i = 0; i < n_jobs; ++i) {
        const f9_job& J = jobs[i];
        const JobShape& S = shapes[(size_t) i];
        f9_unit u{}; u.job = i; u.ch0 = 0; u.num_ch = J.numCh; u.n0 = 0; u.num_out = 0; u.tail_only = 0; u.cost = std::max<long long>(S.cost, 1);
        const bool big = n_devices > 1 && job_is_plain(J) && S.convert && S.out_frames > 0 && S.cost * 4 > share * 3;
        if (!big) { U.push_back(u); continue; }
        const bool chSplit = !J.src_pcm && !(J.flags & F9_JOB_PCM24) && J.numCh > 1;
        const bool tSplit = !(J.flags & F9_JOB_REMOVE_DC);
        const long long piece = std::max<long long>(1 << 20, seg_out > 0 ? seg_out : share / 4);       // outputs (all channels of the unit) per unit
        int chGroups = 1;
        if (chSplit) chGroups = (int) std::min<long long>(J.numCh, std::max<long long>(1, (S.cost + piece - 1) / piece));
        if (!tSplit && chGroups == 1) { U.push_back(u); continue; }
        for (int g = 0; g < chGroups; ++g) {
            const int c0 = (int) ((long long) J.numCh * g / chGroups), c1 = (int) ((long long) J.numCh * (g + 1) / chGroups);
            const long long perCh = std::max<long long>(1, piece / std::max(1, c1 - c0));
            const long long segLen = tSplit ? (seg_out > 0 ? seg_out : std::max<long long>(1 << 18, perCh)) : S.out_frames;
            for (long long n0 = 0; n0 < S.out_frames; n0 += segLen) {
                f9_unit v = u; v.ch0 = c0; v.num_ch = c1 - c0;
                const long long cnt = std::min<long long>(segLen, S.out_frames - n0);
                if (tSplit) { v.n0 = n0; v.num_out = cnt; }                    // otherwise: the whole conversion of these channels
                v.cost = cnt * v.num_ch;
                U.push_back(v);
            }
        }
        if (J.flags & F9_JOB_TAIL_SCAN) { f9_unit t = u; t.tail_only = 1; t.cost = std::max<long long>(1, (long long) J.numCh * std::max(0, J.captured_frames - J.original_length) / 4); U.push_back(t); }
    }
    if ((int) U.size() > max_units) { *n_units = (int) U.size(); return F9_ERR_NOMEM; }
    std::vector<long long> costs(U.size()); std::vector<int> bins(U.size());
    for (size_t k = 0; k < U.size(); ++k) costs[k] = U[k].cost;
    f9_shard_units(costs.data(), (int) U.size(), n_devices, bins.data());
    for (size_t k = 0; k < U.size(); ++k) { U[k].device = bins[k]; units[k] = U[k]; }
    *n_units = (int) U.size();
    return F9_OK;
}

// Run the units assigned to `device` (every unit when device < 0) on ctx.  unit_results[k] is filled for the units that ran
// (status F9_OK or the error); units of other devices are left untouched.  For a unit that is a whole job the result is the job's;
// for a piece of a job out_frames is the piece's output count and tail_stop_frame is set only by the job's tail unit.
int f9_process_units(f9_context* ctx, const f9_job* jobs, int n_jobs, const f9_unit* units, int n_units, int device, f9_result* unit_results) {
    if (!ctx) return F9_ERR_INVALID;
    if (n_units < 0 || (n_units > 0 && (!units || !jobs || !unit_results))) return ctx->fail(F9_ERR_INVALID, "bad unit array");
    std::vector<f9_job> dj; std::vector<f9_job_ext> dx; std::vector<int> which; std::vector<long long> tailShift;
    std::vector<std::vector<const float*>> inPtrs; std::vector<std::vector<float*>> outPtrs;
    dj.reserve((size_t) n_units); dx.reserve((size_t) n_units); inPtrs.reserve((size_t) n_units); outPtrs.reserve((size_t) n_units);
    for (int k = 0; k < n_units; ++k) {
        const f9_unit& u = units[k];
        if (device >= 0 && u.device != device) continue;
        if (u.job < 0 || u.job >= n_jobs) { unit_results[k] = f9_result{}; unit_results[k].status = F9_ERR_INVALID; continue; }
        const f9_job& J = jobs[u.job];
        f9_job D = J; f9_job_ext X; long long shift = 0;
        const bool whole = u.num_out == 0 && !u.tail_only && u.ch0 == 0 && u.num_ch == J.numCh;
        if (!whole) {
            if (!job_is_plain(J) || u.ch0 < 0 || u.num_ch <= 0 || u.ch0 + u.num_ch > J.numCh) { unit_results[k] = f9_result{}; unit_results[k].status = F9_ERR_INVALID; continue; }
            const JobShape S = shape_of(J);
            inPtrs.emplace_back(); outPtrs.emplace_back();
            std::vector<const float*>& ip = inPtrs.back(); std::vector<float*>& op = outPtrs.back();
            if (u.tail_only) {
                // the capture from one window before the scan's first poll: same polls, same windows, frames shifted by `shift`
                const long long startFrame = (long long) J.original_length + std::max(S.latency_frames, 0);
                shift = std::max<long long>(0, std::min<long long>(startFrame - J.tail_window, J.captured_frames));
                D.flags = F9_JOB_TAIL_SCAN; D.out = nullptr; D.out_pcm24 = nullptr; D.out_capacity = 0;
                D.latency_samples = 0; D.original_length = (int) (startFrame - shift); D.fs_out = D.fs_in;
                D.captured_frames = (int) (J.captured_frames - shift);
                if (J.src_pcm) {
                    const int bps = J.src_fmt == F9_PCM_U8 ? 1 : J.src_fmt == F9_PCM_S16LE ? 2 : J.src_fmt == F9_PCM_S24LE ? 3 : 4;
                    D.src_pcm = (const unsigned char*) J.src_pcm + (size_t) shift * J.src_ch * bps;
                } else { for (int c = 0; c < J.numCh; ++c) ip.push_back(J.captured[c] + shift); D.captured = ip.data(); }
                X.tail_only = 1;
            } else {
                // channels [ch0, ch0 + num_ch), outputs [n0, n0 + num_out) (num_out == 0: the whole conversion)
                D.flags = J.flags & ~F9_JOB_TAIL_SCAN;
                D.numCh = u.num_ch;
                D.latency_samples = S.latency_frames * u.num_ch;             // same latency in frames (:835 divides by the channel count)
                long long first = 0, last = S.copied;
                if (u.num_out > 0) {
                    long long f = 0, l = 0;
                    if (f9_resample_segment_input_range(J.interp_kind, J.fs_in / J.fs_out, u.n0, u.num_out, &f, &l) != F9_OK) { unit_results[k] = f9_result{}; unit_results[k].status = F9_ERR_INVALID; continue; }
                    first = std::max<long long>(0, f) / kSegAlign * kSegAlign; last = std::min<long long>(S.copied, l);
                    if (last < first) last = first;
                    D.flags &= ~F9_JOB_REMOVE_DC;
                    D.latency_samples = 0; D.original_length = (int) (last - first); D.captured_frames = (int) (last - first);
                    X.n0 = u.n0; X.num_out = u.num_out; X.in_offset = first;
                }
                const long long inShift = u.num_out > 0 ? (long long) std::max(S.start, 0) + first : 0;
                if (J.src_pcm) {
                    const int bps = J.src_fmt == F9_PCM_U8 ? 1 : J.src_fmt == F9_PCM_S16LE ? 2 : J.src_fmt == F9_PCM_S24LE ? 3 : 4;
                    D.src_pcm = (const unsigned char*) J.src_pcm + (size_t) inShift * J.src_ch * bps;       // never split by channel
                } else { for (int c = 0; c < u.num_ch; ++c) ip.push_back(J.captured[u.ch0 + c] + inShift); D.captured = ip.data(); }
                if (J.out) { for (int c = 0; c < u.num_ch; ++c) op.push_back(J.out[u.ch0 + c] + u.n0); D.out = op.data(); D.out_capacity = J.out_capacity - (int) u.n0; }
                if (J.out_pcm24 && (J.flags & F9_JOB_PCM24)) D.out_pcm24 = J.out_pcm24 + (size_t) u.n0 * J.numCh * 3;
            }
        }
        dj.push_back(D); dx.push_back(X); which.push_back(k); tailShift.push_back(shift);
    }
    if (dj.empty()) return F9_OK;
    std::vector<f9_result> rr(dj.size());
    const int rc = f9_process_batch_ext(ctx, dj.data(), dx.data(), (int) dj.size(), rr.data());
    for (size_t t = 0; t < dj.size(); ++t) {
        f9_result R = rr[t];
        if (dx[t].tail_only && R.tail_stop_frame >= 0) R.tail_stop_frame += tailShift[t];
        unit_results[which[t]] = R;
    }
    return rc;
}

// Merge unit results into per-job results (status: the worst of the job's units; lengths from the job itself).
int f9_merge_unit_results(const f9_job* jobs, int n_jobs, const f9_unit* units, const f9_result* unit_results, int n_units, f9_result* results) {
    if (n_jobs < 0 || n_units < 0 || (n_jobs > 0 && (!jobs || !results)) || (n_units > 0 && (!units || !unit_results))) return F9_ERR_INVALID;
    int worst = F9_OK;
    for (int i = 0; i < n_jobs; ++i) { results[i] = f9_result{}; results[i].tail_stop_frame = -1; results[i].status = F9_ERR_INVALID; }
    std::vector<char> seen((size_t) n_jobs, 0);
    for (int k = 0; k < n_units; ++k) {
        const f9_unit& u = units[k];
        if (u.job < 0 || u.job >= n_jobs) continue;
        f9_result& R = results[u.job];
        const f9_result& P = unit_results[k];
        const bool whole = u.num_out == 0 && !u.tail_only && u.ch0 == 0 && u.num_ch == jobs[u.job].numCh;
        if (whole) { R = P; seen[(size_t) u.job] = 1; if (P.status) worst = P.status; continue; }
        if (!seen[(size_t) u.job]) {
            const JobShape S = shape_of(jobs[u.job]);
            R.status = F9_OK; R.latency_frames = S.latency_frames; R.trim_start = S.start; R.frames_copied = S.copied; R.out_frames = S.out_frames;
            seen[(size_t) u.job] = 1;
        }
        if (P.status) { R.status = P.status; worst = P.status; }
        if (u.tail_only) { R.tail_stop_frame = P.tail_stop_frame; R.tail_polls = P.tail_polls; }
    }
    return worst;
}

int f9_multi_create(const int* devices, int n_devices, f9_multi** out) {
    if (!out || n_devices <= 0 || !devices) return F9_ERR_INVALID;
    *out = nullptr;
    f9_multi* m = new (std::nothrow) f9_multi();
    if (!m) return F9_ERR_NOMEM;
    for (int i = 0; i < n_devices; ++i) {
        f9_context* c = nullptr;
        const int rc = f9_context_create(devices[i], &c);
        if (rc) { for (f9_context* p : m->ctx) f9_context_destroy(p); delete m; return rc; }
        m->ctx.push_back(c);
    }
    *out = m;
    return F9_OK;
}
void f9_multi_destroy(f9_multi* m) {
    if (!m) return;
    for (f9_context* c : m->ctx) f9_context_destroy(c);
    delete m;
}
int f9_multi_device_count(const f9_multi* m) { return m ? (int) m->ctx.size() : 0; }
f9_context* f9_multi_context(f9_multi* m, int i) { return (m && i >= 0 && i < (int) m->ctx.size()) ? m->ctx[(size_t) i] : nullptr; }
const char* f9_multi_last_error(const f9_multi* m) { return m ? m->err.c_str() : ""; }

int f9_multi_process_batch(f9_multi* m, const f9_job* jobs, int n_jobs, f9_result* results, int* out_device) {
    if (!m || n_jobs < 0 || (n_jobs > 0 && (!jobs || !results))) return F9_ERR_INVALID;
    const int nd = (int) m->ctx.size();
    std::vector<f9_unit> units((size_t) n_jobs + 64);
    int nu = 0;
    int rc = f9_multi_partition(jobs, n_jobs, nd, 0, units.data(), (int) units.size(), &nu);
    if (rc == F9_ERR_NOMEM) { units.resize((size_t) nu); rc = f9_multi_partition(jobs, n_jobs, nd, 0, units.data(), nu, &nu); }
    if (rc) { m->err = "partition failed"; return rc; }
    std::vector<f9_result> ur((size_t) std::max(nu, 1));
    std::vector<int> rcs((size_t) nd, F9_OK);
    std::vector<std::thread> th;
    for (int d = 0; d < nd; ++d)
        th.emplace_back([&, d] { rcs[(size_t) d] = f9_process_units(m->ctx[(size_t) d], jobs, n_jobs, units.data(), nu, d, ur.data()); });
    for (auto& t : th) t.join();
    int worst = f9_merge_unit_results(jobs, n_jobs, units.data(), ur.data(), nu, results);
    for (int d = 0; d < nd; ++d) if (rcs[(size_t) d] && !worst) { worst = rcs[(size_t) d]; m->err = f9_last_error(m->ctx[(size_t) d]); }
    if (out_device) {
        for (int i = 0; i < n_jobs; ++i) out_device[i] = -1;
        for (int k = 0; k < nu; ++k) if (out_device[units[(size_t) k].job] < 0) out_device[units[(size_t) k].job] = units[(size_t) k].device;
    }
    return worst;
}

}  // extern "C"
