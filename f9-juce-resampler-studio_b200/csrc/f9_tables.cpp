// Host-side tables of the resampler: the WindowedSinc lookup table, rational-ratio detection and the
// per-phase (polyphase) tap weights.  Compiled with -ffp-contract=off: the float expressions below
// must round exactly like the scalar interpolator they tabulate.
//
// [JUCE 8.0.10 juce_audio_basics/utilities/juce_WindowedSincInterpolator.cpp, juce_LagrangeInterpolator.cpp,
//  juce_Interpolators.h -- not vendored by the reference (JuceLibraryCode/JuceHeader.h:16 links the module);
//  algorithm as recorded in SURVEY.md Appendix A.]
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "f9_internal.cuh"

namespace f9 {

// Stand-in for WindowedSincTraits::lookupTable[10001] (100 points per zero crossing, 100 crossings).
// JUCE ships a literal table whose window is not documented; f9_sinc_table_set() installs the real one.
void make_default_sinc_table(float* t) {
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < kSincTableSize; ++i) {
        if (i == 0) { t[i] = 1.0f; continue; }
        if (i % 100 == 0) { t[i] = 0.0f; continue; }
        const double x = (double) i / 100.0;
        const double s = std::sin(pi * x) / (pi * x);
        const double w = 0.5 * (1.0 + std::cos(pi * x / 100.0));
        t[i] = (float) (s * w);
    }
}

int interp_memory(int kind) {
    switch (kind) {
        case F9_WINDOWED_SINC: return 200;
        case F9_LAGRANGE: return 5;
        case F9_CATMULL_ROM: return 4;
        case F9_LINEAR: return 2;
        case F9_ZERO_ORDER_HOLD: return 1;
        default: return 0;
    }
}
float interp_latency(int kind) {
    switch (kind) {
        case F9_WINDOWED_SINC: return 100.0f;
        case F9_LAGRANGE: return 2.0f;
        case F9_CATMULL_ROM: return 2.0f;
        case F9_LINEAR: return 1.0f;
        default: return 0.0f;
    }
}

bool find_rational(double ratio, int max_q, long long* p_out, long long* q_out) {
    if (!(ratio > 0.0) || !std::isfinite(ratio)) return false;
    // continued-fraction convergents h/k of ratio
    long long h0 = 0, h1 = 1, k0 = 1, k1 = 0;
    double x = ratio;
    for (int it = 0; it < 64; ++it) {
        double a = std::floor(x);
        if (a > 1e15) break;
        long long ai = (long long) a;
        long long h2 = ai * h1 + h0, k2 = ai * k1 + k0;
        if (k2 > max_q || h2 > (1LL << 40)) break;
        h0 = h1; h1 = h2; k0 = k1; k1 = k2;
        if ((double) h1 / (double) k1 == ratio) { *p_out = h1; *q_out = k1; return true; }
        double frac = x - a;
        if (frac <= 0.0) break;
        x = 1.0 / frac;
    }
    return false;
}

static void sinc_weights(const float* table, float offset, float* w) {
    // Walks the 200 contributing taps exactly like the scalar loop does (i = -100 .. 99; i = 100 never
    // contributes because sincPosition >= 100 there), keeping its incremental index / firstFrac state.
    const int numCrossings = 100;
    const float floatCrossings = (float) numCrossings;
    float firstFrac = 0.0f, lastSincPosition = -1.0f;
    int index = 0, sign = -1;
    for (int i = -numCrossings; i < numCrossings; ++i) {
        const float sincPosition = (1.0f - offset) + (float) i;
        if (i == -numCrossings || (sincPosition >= 0 && lastSincPosition < 0)) {
            const float indexFloat = (sincPosition >= 0.f ? sincPosition : -sincPosition) * 100.0f;
            const float indexFloored = std::floor(indexFloat);
            index = (int) indexFloored;
            firstFrac = indexFloat - indexFloored;
            sign = (sincPosition < 0 ? -1 : 1);
        }
        float v = 0.0f;
        if (sincPosition == 0.0f) v = 1.0f;
        else if (sincPosition < floatCrossings && sincPosition > -floatCrossings) {
            const float v1 = table[index], v2 = table[index + 1];
            v = v1 + (firstFrac * (v2 - v1));
        }
        w[i + numCrossings] = v;
        lastSincPosition = sincPosition;
        index += 100 * sign;
    }
}

static float lagrange_weight(int k, float offset) {
    float a = 1.0f;
    for (int j = 0; j < 5; ++j) {
        if (j == k) continue;
        a *= ((float) (j - 2) - offset) * (1.0f / (float) (j - k));
    }
    return a;
}

static float catmull(float y0, float y1, float y2, float y3, float offset) {
    const float halfY0 = 0.5f * y0, halfY3 = 0.5f * y3;
    return y1 + offset * ((0.5f * y2 - halfY0)
             + (offset * (((y0 + 2.0f * y2) - (halfY3 + 2.5f * y1))
             + (offset * ((halfY3 + 1.5f * y1) - (halfY0 + 1.5f * y2))))));
}

void tap_weights(int kind, const float* sinc_table, float offset, float* w) {
    switch (kind) {
        case F9_WINDOWED_SINC: sinc_weights(sinc_table, offset, w); break;
        case F9_LAGRANGE: for (int k = 0; k < 5; ++k) w[k] = lagrange_weight(k, offset); break;
        case F9_CATMULL_ROM:
            w[0] = catmull(1, 0, 0, 0, offset); w[1] = catmull(0, 1, 0, 0, offset);
            w[2] = catmull(0, 0, 1, 0, offset); w[3] = catmull(0, 0, 0, 1, offset); break;
        case F9_LINEAR: w[0] = 1.0f - offset; w[1] = offset; break;
        default: w[0] = 1.0f; break;
    }
}

void build_poly(int kind, const float* sinc_table, long long p, long long q, PolyHost* out) {
    const int taps = interp_memory(kind);
    out->p = (int) p; out->q = (int) q; out->taps = taps;
    out->qpad = (int) ((q + 31) / 32 * 32);
    out->B.assign((size_t) q, 0);
    out->W.assign((size_t) taps * out->qpad, 0.0f);
    std::vector<float> w((size_t) taps);
    for (long long k = 0; k < q; ++k) {
        const long long t = k * p;
        out->B[(size_t) k] = (int) (t / q);
        const long long phi = t % q;
        const float offset = (float) ((double) phi / (double) q);
        tap_weights(kind, sinc_table, offset, w.data());
        for (int j = 0; j < taps; ++j) out->W[(size_t) j * out->qpad + (size_t) k] = w[(size_t) j];
    }
}

}  // namespace f9

// ---- host scalars that must round exactly like the reference's (compiled without FMA contraction) ---------
namespace f9 {

// Largest float r >= 0 for which  20*log10f(max(r, floorv)) < thrDb  (Source/MainComponent.cpp:868-875 with
// floorv = 1e-10f).  Returns -1 when no r qualifies.  The predicate is monotone in r (log10f is), so a
// bisection over the float bit patterns finds the boundary; the device then compares in the linear domain.
float largest_rms_below(float thrDb, float floorv) {
    auto below = [&](float r) { return 20.0f * std::log10(r > floorv ? r : floorv) < thrDb; };
    if (!below(0.0f)) return -1.0f;
    uint32_t lo = 0, hi = 0x7f800000u;          // lo: true, hi: +inf
    float fhi; std::memcpy(&fhi, &hi, 4);
    if (below(fhi)) return fhi;
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        float fm; std::memcpy(&fm, &mid, 4);
        if (below(fm)) lo = mid; else hi = mid;
    }
    float r; std::memcpy(&r, &lo, 4);
    return r;
}
// Swift peak predicate (AudioProcessingService.swift:723-729): (m > 0 ? 20*log10f(m) : -160) < thrDb.
// *below0 = value at m == 0; returns the largest m > 0 with the predicate true, or -1.
float largest_peak_below(float thrDb, int* below0) {
    *below0 = (-160.0f < thrDb) ? 1 : 0;
    auto below = [&](float m) { return 20.0f * std::log10(m) < thrDb; };
    uint32_t lo = 1, hi = 0x7f800000u;          // smallest denormal .. +inf
    float flo; std::memcpy(&flo, &lo, 4);
    if (!below(flo)) return -1.0f;
    float fhi; std::memcpy(&fhi, &hi, 4);
    if (below(fhi)) return fhi;
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        float fm; std::memcpy(&fm, &mid, 4);
        if (below(fm)) lo = mid; else hi = mid;
    }
    float r; std::memcpy(&r, &lo, 4);
    return r;
}
float noise_floor_db_from_rms(float rms) { return 20.0f * std::log10(rms > 1e-6f ? rms : 1e-6f); }   // MainComponent.cpp:977-981
float nf_threshold_db(int has_nf, float nf_db, float margin_pct) {                                    // AppState.h:252-258
    if (!has_nf) return -80.0f;
    return nf_db + (nf_db * margin_pct / 100.0f);
}
float threshold_linear(float db) { return std::pow(10.0f, db / 20.0f); }                             // AppState.h:246-249

// generateSineWave scalars (Source/MainComponent.cpp:910-911, :929-931), float arithmetic as written there
float sine_phase_increment(float frequency, float sample_rate) {
    const float pi = 3.14159265358979323846f;                 // juce::MathConstants<float>::pi
    return (frequency * 2.0f * pi) / sample_rate;
}
float sine_phase_after_block(float phase, float inc, int num_samples) {
    const float twoPi = 2.0f * 3.14159265358979323846f;
    phase += inc * num_samples;
    if (phase >= twoPi) phase -= twoPi;
    return phase;
}

// GenericInterpolator position recurrence on the host (exact JUCE state): returns inputs consumed.
int run_position_chain(double* pos_io, double ratio, int num_out) {
    double pos = *pos_io;
    int used = 0;
    while (num_out > 0) {
        while (pos >= 1.0) { ++used; pos -= 1.0; }
        pos += ratio;
        --num_out;
    }
    *pos_io = pos;
    return used;
}

// Closed-form position (same arithmetic as the device's pos_generic) for host-side planning.
void position_closed_form(double pos0, double ratio, long long n, long long* c, double* frac_out) {
    const double dn = (double) n;
    const double hi = dn * ratio;
    const double lo = std::fma(dn, ratio, -hi);
    const double s = pos0 + hi;
    const double bb = s - pos0;
    double err = (pos0 - (s - bb)) + (hi - bb);
    err += lo;
    double fl = std::floor(s);
    double frac = (s - fl) + err;
    if (frac < 0.0) { fl -= 1.0; frac += 1.0; }
    else if (frac >= 1.0) { fl += 1.0; frac -= 1.0; }
    *c = (long long) fl;
    if (frac_out) *frac_out = frac;
}

}  // namespace f9

// ---- band-aligned polyphase tables for the register-tiled FIR (see BandedDev in f9_internal.cuh) ----------
namespace f9 {

void build_banded(int kind, const float* sinc_table, long long p, long long q, int TK, int Gpad, BandedHost* out) {
    const int taps = interp_memory(kind);
    const int G = (int) ((q + TK - 1) / TK);
    if (Gpad < G) Gpad = G;
    std::vector<int> B((size_t) q);
    for (long long k = 0; k < q; ++k) B[(size_t) k] = (int) ((k * p) / q);
    int maxShift = 0;
    for (int g = 0; g < G; ++g) {
        const long long k0 = (long long) g * TK, k1 = std::min<long long>(q, k0 + TK) - 1;
        maxShift = std::max(maxShift, B[(size_t) k1] - B[(size_t) k0]);
    }
    const int Tmax = (taps + maxShift + 1) & ~1;
    out->p = (int) p; out->q = (int) q; out->taps = taps; out->TK = TK; out->G = G; out->Gpad = Gpad; out->Tmax = Tmax;
    out->C.assign((size_t) Gpad * Tmax * TK, 0.0f);
    out->wmin.assign((size_t) Gpad, 0);
    std::vector<float> w((size_t) taps);
    for (int g = 0; g < G; ++g) {
        const long long k0 = (long long) g * TK;
        out->wmin[(size_t) g] = B[(size_t) k0] - (taps - 1);
        for (int j = 0; j < TK; ++j) {
            const long long k = k0 + j;
            if (k >= q) break;                                           // padding slots keep zero weights
            const long long phi = (k * p) % q;
            const float offset = (float) ((double) phi / (double) q);    // what (float) subSamplePos is at this phase
            tap_weights(kind, sinc_table, offset, w.data());
            const int shift = B[(size_t) k] - B[(size_t) k0];
            for (int t = 0; t < taps; ++t)
                out->C[((size_t) g * Tmax + (size_t) (t + shift)) * TK + (size_t) j] = w[(size_t) t];
        }
    }
}

}  // namespace f9

// ---- tensor-core FIR tables (see UmmaHost in f9_internal.cuh) ---------------------------------------------
namespace f9 {
namespace {

// float -> IEEE binary16 bits, round to nearest even (normals, subnormals, overflow to infinity).
uint16_t f32_to_f16_bits(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t) (sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u));
    if (x >= 0x477ff000u) return (uint16_t) (sign | 0x7c00u);                   // rounds to >= 65520 -> inf
    if (x < 0x33000001u) return (uint16_t) sign;                               // < 2^-25 (or exactly 2^-25: ties to even 0)
    const int e = (int) (x >> 23) - 127;
    uint32_t m = (x & 0x7fffffu) | 0x800000u;                                  // 24-bit significand
    int shift;                                                                 // bits dropped from m
    uint32_t base;
    if (e >= -14) { shift = 13; base = (uint32_t) (e + 15) << 10; m &= 0x7fffffu; }
    else { shift = 13 + (-14 - e); base = 0; }                                 // subnormal: keep the leading one
    const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    uint32_t r = base + q;
    if (rem > half || (rem == half && (r & 1))) ++r;                           // carries ripple into the exponent correctly
    return (uint16_t) (sign | r);
}
float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
    const int e = (h >> 10) & 31; const uint32_t m = h & 0x3ffu;
    float v;
    if (e == 0) v = std::ldexp((float) m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = std::ldexp((float) (m | 0x400u), e - 25);
    uint32_t b; std::memcpy(&b, &v, 4); b |= sign; std::memcpy(&v, &b, 4);
    return v;
}
inline long long floor16(long long x) { return x >= 0 ? (x / 16) * 16 : -(((-x) + 15) / 16) * 16; }

// Address shift of the plan being built (build_umma's `shift`, 0 .. 3): K origins are congruent to -shift modulo 16, so that rows
// of segments whose first sample sits `shift` floats past a 16-byte boundary still start on 16 bytes (TMA feed, aligned loads).
static thread_local int t_geomShift = 0;
struct GroupGeom { long long t0; int ksteps; long long wend; };   // K origin of the group's first step (multiple of 16 minus the shift), steps, window end
void group_geom(long long p, long long q, int taps, int NB, int g, GroupGeom* out) {
    const long long k0 = (long long) NB * g, k1 = std::min<long long>(q, k0 + NB) - 1;
    const long long wmin = (k0 * p) / q - (taps - 1), wend = (k1 * p) / q + 1;
    out->t0 = floor16(wmin + t_geomShift) - t_geomShift;
    out->ksteps = (int) ((wend - out->t0 + 15) / 16);
    out->wend = wend;
}
int umma_pool_slots(int NB, int GBL, int accCols) {        // accumulator-split pool: power of two, NB columns per slot
    const int spare = accCols - GBL * 2 * NB;              // the columns above accCols hold the operand ring
    int n = 0;
    for (int c = 8; c >= 2; c /= 2) if (c * NB <= spare) { n = c; break; }
    return n;
}

}  // namespace

// Weight image of the Hankel-operand FIR (f9_hankel.cu): A[l = i*L + k, t] = w_k[t - shift - i], shift = 209 - taps, as two
// row-major matrices [128 lanes][16 KS] of fp16 (heads, then tails * 2048): each lane's row goes into TMEM as the MMAs' M-side operand.
bool build_hankel(int kind, const float* sinc_table, int L, std::vector<uint8_t>* image, int* KS_out) {
    const int taps = interp_memory(kind);
    if (taps < 1 || taps > 200 || (L != 2 && L != 4 && L != 8 && L != 16)) return false;
    const int R = 128 / L, KS = (R + 208 + 15) / 16, shift = 209 - taps;
    image->assign((size_t) 2 * KS * 4096, 0);
    std::vector<float> w((size_t) taps);
    for (int k = 0; k < L; ++k) {
        tap_weights(kind, sinc_table, (float) ((double) k / (double) L), w.data());      // phase (k*p) mod q / q with p = 1, q = L
        for (int i = 0; i < R; ++i) {
            const int l = i * L + k;
            for (int j = 0; j < taps; ++j) {
                const int t = i + j + shift;
                const float wv = w[(size_t) j];
                const uint16_t h0 = f32_to_f16_bits(wv);
                const uint16_t h1 = f32_to_f16_bits((wv - f16_bits_to_f32(h0)) * 2048.0f);
                const size_t off = ((size_t) l * (size_t) (KS * 16) + (size_t) t) * 2;
                std::memcpy(image->data() + off, &h0, 2);
                std::memcpy(image->data() + (size_t) KS * 4096 + off, &h1, 2);
            }
        }
    }
    *KS_out = KS;
    return true;
}

size_t umma_smem_bytes(int maxEntries, int NB, int stages, bool tma, bool cta2) {
    const size_t w = (size_t) maxEntries * NB * (cta2 ? 32 : 64);   // tile: 2 K chunks x 2*NB rows x 16 B (half of the rows per CTA of a pair)
    // register loader: converted stages (fp16 head + tail, padded K chunks); TMA feed: raw fp32 boxes of 128 rows x 128 B,
    // 1024-byte aligned for the 128-byte swizzle
    const size_t ring = tma ? (size_t) stages * 16384 + 1024 : (size_t) stages * 8 * (128 * 16 + 32);
    const size_t epi = 128 * 20 * 4 + (size_t) maxEntries * 16;                // transpose buffer + the issue lists (UmmaOp)
    const size_t bars = (size_t) (2 * stages + 2 * kUmmaMaxGroups + 8) * 8 + 16;
    return w + ring + epi + bars + 128;                    // + alignment slack
}

// Relative cost (SM cycles per output) of running ratio p/q with groups of NB slots, GBL groups per block.  Tensor pipe:
// tcgen05.cp 64 clk per 4 KB operand tile (4 per stage of 32 samples), three MMAs per active (K step, group) at
// max(16, NB/2) clk each (measured: an MMA that fetches a fresh B tile costs >= 16 clk).  Issue: ~45 clk per MMA.
// Load/store unit: one wavefront per 128 bytes loaded, stored to shared memory, and three per 128 bytes of output.
double umma_cost_per_output(int taps, long long p, long long q, int NB, int GBL, size_t* smem2) {
    const int G = (int) ((q + NB - 1) / NB);
    const int nGB = (G + GBL - 1) / GBL;
    double tensor = 0.0, issue = 0.0, lsu = 0.0; int maxEntries = 0;
    for (int b = 0; b < nGB; ++b) {
        long long lo = (1LL << 60), hi = -(1LL << 60); int entries = 0;
        for (int g = b * GBL; g < std::min(G, (b + 1) * GBL); ++g) {
            GroupGeom gg; group_geom(p, q, taps, NB, g, &gg);
            lo = std::min(lo, gg.t0); hi = std::max(hi, gg.t0 + 16LL * gg.ksteps); entries += gg.ksteps;
        }
        const int nK = (int) ((hi - lo) / 16);
        tensor += 128.0 * nK + 3.0 * std::max(16, NB / 2) * entries;
        issue += 135.0 * entries + 150.0 * nK;
        lsu += 2.0 * nK * 16 * 128 * 4 / 128.0;
        maxEntries = std::max(maxEntries, entries);
    }
    lsu += 3.0 * (double) q * 128 * 4 / 128.0;
    if (smem2) *smem2 = umma_smem_bytes(maxEntries, NB, 2);
    return std::max(std::max(tensor, issue), lsu) / (128.0 * (double) q);
}

// Tensor-core plan for ratio p/q: scale p/q by m so that a period has 64..224 slots; choose the scaling, the group width
// (16 or 32 slots) and the block size with the lowest modelled cost whose tables fit shared memory with two staging buffers.
// *m = 0 when no plan fits.
// Accumulator-split plan of (p, q, NB, GBL): pool slots (0 = no split), split step, operand-ring depth.  See build_umma.
static int umma_pool_plan(long long p, long long q, int taps, int NB, int GBL, int* split_out, int* aSlots_out) {
    const int G = (int) ((q + NB - 1) / NB);
    int split = 0;
    auto pool_for = [&](int accCols) {
        int poolN = umma_pool_slots(NB, GBL, accCols); split = 0;
        if (taps < 64 || poolN == 0) return 0;
        for (int g = 0; g < G; ++g) {
            GroupGeom gg; group_geom(p, q, taps, NB, g, &gg);
            const long long k1 = std::min<long long>(q, (long long) NB * g + NB) - 1;
            const long long centre = (k1 * p) / q - (taps - 1) - gg.t0 + taps / 2 + 2;     // K offset of the last slot's centre tap
            split = std::max(split, (int) ((centre + 8 + 15) / 16));
        }
        for (int g = 0; g < G && poolN > 0; ++g) {
            GroupGeom gg; group_geom(p, q, taps, NB, g, &gg);
            if (gg.ksteps <= split) poolN = 0;                                              // a group without a second part
            const int gl = g % GBL;
            if (gl >= poolN && poolN > 0) {                                                 // slot reuse inside a tile: the previous
                GroupGeom gp; group_geom(p, q, taps, NB, g - poolN, &gp);                   // user must have finished before
                if (gg.t0 / 16 + split < gp.t0 / 16 + gp.ksteps) poolN = 0;                 // this group crosses its split
            }
        }
        return poolN;
    };
    int aSlots = 2, poolN = pool_for(448);
    if (GBL * 2 * NB <= 384) {
        const int splitWide = split, poolNarrow = pool_for(384);
        if ((poolN == 0) == (poolNarrow == 0)) { aSlots = 4; poolN = poolNarrow; }
        else split = splitWide;
    }
    if (poolN == 0) split = 0;
    *split_out = split; *aSlots_out = aSlots;
    return poolN;
}

void umma_choose_plan(int taps, long long p, long long q, long long* m_out, int* NB_out, int* GBL_out, int forceNB) {
    double best = 1e30; *m_out = 0; *NB_out = 0; *GBL_out = 0;
    for (long long m = 1; m * q <= 16LL * kUmmaMaxGroups * kUmmaMaxBlocks && m * p + taps + 48 <= 16 * kUmmaMaxNK; ++m) {
        const long long ps = p * m, qs = q * m;
        if (qs < 48 && (m + 1) * q <= 224) continue;                           // too few slots per period: keep scaling
        for (int NB : {32, 16}) {
            if (forceNB && forceNB != NB) continue;
            const int G = (int) ((qs + NB - 1) / NB), maxG = std::min(kUmmaMaxGroups, 448 / (2 * NB));
            // a period scaled for alignment (m > 1 of an unaligned p) splits best into blocks of whole original periods: every
            // block then has the window of one period, as the unscaled plan's tiles have
            int nGB0 = (G + maxG - 1) / maxG;
            if (m > 1 && (p & 3) != 0)
                for (int d = (int) m; d > nGB0; --d)
                    if (m % d == 0 && G % d == 0 && G / d <= maxG) { nGB0 = d; break; }
            for (int nGB = nGB0; nGB <= kUmmaMaxBlocks && nGB <= G; ++nGB) {
                const int GBL = (G + nGB - 1) / nGB;
                size_t smem2 = 0;
                double c = umma_cost_per_output(taps, ps, qs, NB, GBL, &smem2);
                if (smem2 > 227 * 1024) continue;
                if (ps & 3) c *= 1.4;                           // rows not 16-byte aligned: register loader instead of the TMA feed
                // Upsampling with a long window: the groups of a block all overlap in time, and without the accumulator split the
                // truncation error reaches the tolerance (44.1 -> 96 k measured 1.125 x 2^-20 unsplit).  Take more, smaller blocks
                // until every group of a block has its own pool slot; the extra staging is the price.
                if (ps < qs && taps >= 64) {
                    int sp = 0, as = 0;
                    if (umma_pool_plan(ps, qs, taps, NB, GBL, &sp, &as) == 0) {
                        if (nGB < kUmmaMaxBlocks && nGB < G) continue;
                        c *= 8.0;                               // no block count gives the split at this scale: last resort only
                    }
                }
                if (c < 0.97 * best) { best = c; *m_out = m; *GBL_out = GBL; *NB_out = NB; }    // ties go to the smaller plan
                break;                                                          // more blocks only cost more
            }
        }
        // rows not yet 16-byte aligned: a multiple of the period may be, which buys the TMA feed (measured at 147/160: 1.6x for
        // the 200-tap kernel, 1.33x for Lagrange once the blocks are whole original periods)
        if (qs >= 224 && ((ps & 3) == 0 || m >= 4)) break;
    }
}

bool build_umma(int kind, const float* sinc_table, long long p, long long q, int NB, int GBL, UmmaHost* out, int shift) {
    struct ShiftGuard { ShiftGuard(int v) { t_geomShift = v; } ~ShiftGuard() { t_geomShift = 0; } } guard(shift & 3);
    const int taps = interp_memory(kind);
    if (NB != 16 && NB != 32) return false;
    const int G = (int) ((q + NB - 1) / NB);
    if (GBL < 1 || GBL > kUmmaMaxGroups || GBL * 2 * NB > 448) return false;
    const int nGB = (G + GBL - 1) / GBL;
    if (nGB > kUmmaMaxBlocks) return false;
    *out = UmmaHost();
    out->p = (int) p; out->q = (int) q; out->taps = taps; out->NB = NB; out->G = G; out->GBL = GBL; out->nGB = nGB;
    const int tileBytes = NB * 64, chunkBytes = NB * 32;
    // accumulation split: only for long windows; the split step is the same for every group (first step past every slot's
    // centre tap), which keeps "past the split" a bottom-end range of the active groups.
    // TMEM budget: 512 columns = accumulators (2*NB per group) + pool + operand ring (32 columns per stage).  A ring of four
    // stages (accumulators + pool <= 384 columns) decouples the converters from the MMAs; plans whose pool needs the room
    // (groups that overlap a lot in time, e.g. upsampling) keep the two-stage ring.
    int split = 0, aSlots = 2;
    const int poolN = umma_pool_plan(p, q, taps, NB, GBL, &split, &aSlots);
    out->aSlots = aSlots;
    out->poolN = poolN; out->split = split;
    std::vector<float> w((size_t) taps);
    for (int b = 0; b < nGB; ++b) {
        UmmaBlockInfo& BI = out->blk[b];
        const int g0 = b * GBL, g1 = std::min(G, g0 + GBL);
        std::vector<GroupGeom> geo((size_t) (g1 - g0));
        long long lo = (1LL << 60), hi = -(1LL << 60), wendMax = -(1LL << 60);
        for (int g = g0; g < g1; ++g) {
            group_geom(p, q, taps, NB, g, &geo[(size_t) (g - g0)]);
            lo = std::min(lo, geo[(size_t) (g - g0)].t0);
            hi = std::max(hi, geo[(size_t) (g - g0)].t0 + 16LL * geo[(size_t) (g - g0)].ksteps);
            wendMax = std::max(wendMax, geo[(size_t) (g - g0)].wend);
        }
        BI.U0 = (int) lo; BI.nK = (int) ((hi - lo) / 16); BI.nGroups = g1 - g0; BI.slot0 = g0 * NB;
        BI.nStages = std::max((BI.nK + 1) / 2, (int) ((wendMax - lo + 3 + 31) / 32));
        BI.wOff = (int) out->W.size();
        if (BI.nK > kUmmaMaxNK) return false;
        // per-slot taps of every group of the block, then the tiles in schedule order (K step major, group minor)
        std::vector<std::vector<float>> gw((size_t) (g1 - g0));
        std::vector<std::vector<int>> gshift((size_t) (g1 - g0));
        for (int g = g0; g < g1; ++g) {
            auto& W = gw[(size_t) (g - g0)]; auto& S = gshift[(size_t) (g - g0)];
            W.assign((size_t) NB * taps, 0.0f); S.assign((size_t) NB, -1);
            for (int s = 0; s < NB; ++s) {
                const long long k = (long long) NB * g + s;
                if (k >= q) break;
                const long long phi = (k * p) % q;
                const float offset = (float) ((double) phi / (double) q);
                tap_weights(kind, sinc_table, offset, w.data());
                std::memcpy(&W[(size_t) s * taps], w.data(), sizeof(float) * (size_t) taps);
                S[(size_t) s] = (int) ((k * p) / q - (taps - 1) - geo[(size_t) (g - g0)].t0);   // K offset of tap 0 inside the group window
            }
        }
        int entries = 0;
        for (int g = g0; g < g1; ++g) {                                        // tiles group-major, K step minor
            const GroupGeom& gg = geo[(size_t) (g - g0)];
            const int gl = g - g0;
            out->gStart[b][gl] = (uint8_t) ((gg.t0 - lo) / 16);
            out->gSteps[b][gl] = (uint8_t) gg.ksteps;
            out->gTile[b][gl] = (uint16_t) entries;
            if (gl > 0 && (out->gStart[b][gl] < out->gStart[b][gl - 1] ||
                           out->gStart[b][gl] + gg.ksteps < out->gStart[b][gl - 1] + out->gSteps[b][gl - 1])) return false;   // windows move monotonically
            for (int j = 0; j < gg.ksteps; ++j) {
                const size_t base = out->W.size();
                out->W.resize(base + (size_t) tileBytes, 0);
                for (int s = 0; s < NB; ++s) {
                    const int sh = gshift[(size_t) gl][(size_t) s];
                    if (sh < 0) continue;
                    for (int kk = 0; kk < 16; ++kk) {
                        const int tap = 16 * j + kk - sh;
                        if (tap < 0 || tap >= taps) continue;
                        const float wv = gw[(size_t) gl][(size_t) s * taps + (size_t) tap];
                        const uint16_t h0 = f32_to_f16_bits(wv);
                        const uint16_t h1 = f32_to_f16_bits((wv - f16_bits_to_f32(h0)) * 2048.0f);
                        const size_t off0 = base + (size_t) (kk / 8) * chunkBytes + (size_t) s * 16 + (size_t) (kk % 8) * 2;
                        const size_t off1 = off0 + (size_t) NB * 16;
                        std::memcpy(&out->W[off0], &h0, 2); std::memcpy(&out->W[off1], &h1, 2);
                    }
                }
                ++entries;
            }
        }
        BI.nEntries = entries;
        out->maxEntries = std::max(out->maxEntries, entries);
        out->maxNK = std::max(out->maxNK, BI.nK);
    }
    // issue lists: per block, per issuing warp, the (stage, group, K step) entries in the order the warp walks them
    auto build_ops = [&](int b, int nIss, int* opOff, int* opStart) {
        UmmaBlockInfo& BI = out->blk[b];
        *opOff = (int) out->W.size();
        int count = 0;
        for (int w = 0; w < nIss; ++w) {
            opStart[w] = count;
            for (int st = 0; st < BI.nStages; ++st)
                for (int gl = w; gl < BI.nGroups; gl += nIss)
                    for (int h = 0; h < 2; ++h) {
                        const int g0 = out->gStart[b][gl], gn = out->gSteps[b][gl];
                        const int j = 2 * st + h - g0;
                        if (j < 0 || j >= gn) continue;
                        UmmaOp op; std::memset(&op, 0, sizeof op);
                        op.d1Col = (uint16_t) ((2 * gl + 1) * NB);
                        op.bOff = (uint32_t) (out->gTile[b][gl] + j) * 4u * (uint32_t) NB;
                        op.stage = (uint8_t) st; op.h = (uint8_t) h; op.gl = (uint8_t) gl;
                        uint8_t f = 0;
                        if (j > 0) f |= kOpAcc;
                        if (j == 0) f |= kOpWaitDrain;
                        if (j == gn - 1) f |= kOpLast;
                        if (poolN == 0 || j < split) f |= kOpMerged;
                        else {
                            op.poolCol = (uint16_t) (GBL * 2 * NB + (gl & (poolN - 1)) * NB);
                            if (j > split) f |= kOpPoolAcc;
                            if (j == split) {
                                f |= kOpWaitPool;
                                if (gl >= poolN) op.waitGl = (uint8_t) (gl - poolN);                  // same tile
                                else { int lu = gl; while (lu + poolN < BI.nGroups) lu += poolN; op.waitGl = (uint8_t) lu; f |= kOpPoolPrevTile; }
                            }
                        }
                        op.flags = f;
                        const size_t at = out->W.size();
                        out->W.resize(at + sizeof op);
                        std::memcpy(&out->W[at], &op, sizeof op);
                        ++count;
                    }
        }
        opStart[nIss] = count;
        return count == BI.nEntries;
    };
    for (int b = 0; b < nGB; ++b) {
        UmmaBlockInfo& BI = out->blk[b];
        if (!build_ops(b, kUmmaIssuers, &BI.opOff, BI.opStart) || !build_ops(b, kUmmaIssuersTma, &BI.opOffT, BI.opStartT)) return false;
    }
    // CTA-pair kernel: the weight tiles again, split by slot halves (cluster rank r holds slots 16 r .. 16 r + 15 of every group)
    for (int b = 0; b < nGB; ++b) out->blk[b].w2Off[0] = out->blk[b].w2Off[1] = -1;
    if (NB == 32) {
        for (int b = 0; b < nGB; ++b) {
            UmmaBlockInfo& BI = out->blk[b];
            for (int r = 0; r < 2; ++r) {
                BI.w2Off[r] = (int) out->W.size();
                out->W.resize(out->W.size() + (size_t) BI.nEntries * 1024, 0);
                for (int e = 0; e < BI.nEntries; ++e) {
                    const uint8_t* src = out->W.data() + BI.wOff + (size_t) e * tileBytes;
                    uint8_t* dst = out->W.data() + BI.w2Off[r] + (size_t) e * 1024;
                    for (int c = 0; c < 2; ++c)
                        for (int i = 0; i < 16; ++i) {
                            std::memcpy(dst + c * 512 + i * 16, src + (size_t) c * chunkBytes + (size_t) (16 * r + i) * 16, 16);
                            std::memcpy(dst + c * 512 + 256 + i * 16, src + (size_t) c * chunkBytes + (size_t) NB * 16 + (size_t) (16 * r + i) * 16, 16);
                        }
                }
            }
        }
    }
    return true;
}

}  // namespace f9

// ---- diagnostics (host only) -------------------------------------------------------------------------------
// Builds the tensor-core plan of ratio p/q and checks its tables against the fp32 polyphase weights: every (slot, tap)
// appears in exactly one weight tile position, everything else is zero, and head + tail/2048 reproduces the weight.
// Diagnostic for the Hankel-operand FIR's weight image (CPU only): rebuilds it for (kind, 1:L) and checks that lane l = i*L + k
// holds w_k[j] at t = i + j + shift and zeros elsewhere.  Returns the largest |w - (w0 + w1/2048)|, -1 bad arguments, -3 defect.
extern "C" double f9_hankel_selfcheck(int kind, int L, int* info /* 4 ints or NULL: KS, elements per tile, buffer bytes, shared-memory bytes */) {
    using namespace f9;
    const int taps = interp_memory(kind);
    if (taps == 0) return -1.0;
    std::vector<float> table((size_t) kSincTableSize + 1, 0.0f);
    make_default_sinc_table(table.data());
    std::vector<uint8_t> image; int KS = 0;
    if (!build_hankel(kind, table.data(), L, &image, &KS)) return -1.0;
    const int R = 128 / L, shift = 209 - taps, K = KS * 16;
    if (R + 208 > K || image.size() != (size_t) 2 * KS * 4096) return -3.0;
    std::vector<float> w((size_t) taps);
    double maxErr = 0.0;
    for (int l = 0; l < 128; ++l) {
        const int i = l / L, k = l % L;
        tap_weights(kind, table.data(), (float) ((double) k / (double) L), w.data());
        for (int t = 0; t < K; ++t) {
            uint16_t h0, h1;
            std::memcpy(&h0, image.data() + ((size_t) l * K + t) * 2, 2);
            std::memcpy(&h1, image.data() + (size_t) KS * 4096 + ((size_t) l * K + t) * 2, 2);
            const int j = t - shift - i;
            const double want = (j >= 0 && j < taps) ? (double) w[(size_t) j] : 0.0;
            const double got = (double) f16_bits_to_f32(h0) + (double) f16_bits_to_f32(h1) / 2048.0;
            if (want == 0.0 && (h0 & 0x7fff || h1 & 0x7fff)) return -3.0;
            maxErr = std::max(maxErr, std::fabs(want - got));
        }
    }
    if (info) {
        HankelDev D; D.L = L; D.KS = KS; D.elems = hankel_tile_elems(L, KS); D.bufBytes = (D.elems * 2 + 1023) / 1024 * 1024;
        info[0] = KS; info[1] = D.elems; info[2] = D.bufBytes; info[3] = (int) hankel_smem_bytes(D);
    }
    return maxErr;
}

extern "C" double f9_umma_selfcheck(int kind, long long p, long long q, int* info /* 8 ints or NULL */) {
    using namespace f9;
    const int taps = interp_memory(kind);
    if (taps == 0 || p <= 0 || q <= 0) return -1.0;
    std::vector<float> table((size_t) kSincTableSize + 1, 0.0f);
    make_default_sinc_table(table.data());
    long long m = 0; int NB = 0, GBL = 0;
    umma_choose_plan(taps, p, q, &m, &NB, &GBL);
    if (m == 0) return -2.0;
    UmmaHost H;
    if (!build_umma(kind, table.data(), p * m, q * m, NB, GBL, &H)) return -3.0;
    const long long ps = p * m, qs = q * m;
    double maxErr = 0.0;
    std::vector<float> w((size_t) taps);
    for (int b = 0; b < H.nGB; ++b) {
        const UmmaBlockInfo& BI = H.blk[b];
        for (int gl = 0; gl < BI.nGroups; ++gl) {
            const int g = b * H.GBL + gl;
            for (int s = 0; s < NB; ++s) {
                const long long k = (long long) NB * g + s;
                std::vector<double> rec((size_t) taps, 0.0); std::vector<int> seen((size_t) taps, 0);
                const long long tap0 = k < qs ? (k * ps) / qs - (taps - 1) : 0;             // input offset of tap 0 (period relative)
                for (int j = 0; j < H.gSteps[b][gl]; ++j) {
                    const uint8_t* tile = H.W.data() + BI.wOff + (size_t) (H.gTile[b][gl] + j) * NB * 64;
                    for (int kk = 0; kk < 16; ++kk) {
                        uint16_t h0, h1;
                        std::memcpy(&h0, tile + (size_t) (kk / 8) * NB * 32 + (size_t) s * 16 + (size_t) (kk % 8) * 2, 2);
                        std::memcpy(&h1, tile + (size_t) (kk / 8) * NB * 32 + (size_t) (NB + s) * 16 + (size_t) (kk % 8) * 2, 2);
                        const double v = (double) f16_bits_to_f32(h0) + (double) f16_bits_to_f32(h1) / 2048.0;
                        const long long in = BI.U0 + 16LL * (H.gStart[b][gl] + j) + kk;     // input offset this K index reads
                        const long long tap = in - tap0;
                        if (k < qs && tap >= 0 && tap < taps) { rec[(size_t) tap] += v; ++seen[(size_t) tap]; }
                        else if (v != 0.0) return -4.0;                                     // weight outside the slot's window
                    }
                }
                if (k >= qs) continue;
                tap_weights(kind, table.data(), (float) ((double) ((k * ps) % qs) / (double) qs), w.data());
                for (int t = 0; t < taps; ++t) {
                    if (seen[(size_t) t] != 1) return -5.0;                                 // a tap is missing or duplicated
                    maxErr = std::max(maxErr, std::fabs(rec[(size_t) t] - (double) w[(size_t) t]));
                }
            }
        }
    }
    if (info) {
        info[0] = (int) m; info[1] = NB; info[2] = H.G; info[3] = H.GBL; info[4] = H.nGB; info[5] = H.poolN; info[6] = H.split;
        info[7] = (int) umma_smem_bytes(H.maxEntries, NB, 2);
        info[0] |= H.aSlots << 16;                          // operand ring depth in the high half of info[0]
    }
    return maxErr;
}
