// Host-side tables of the resampler: the WindowedSinc lookup table, rational-ratio detection and the
// per-phase (polyphase) tap weights.  Compiled with -ffp-contract=off: the float expressions below
// must round exactly like the scalar interpolator they tabulate.
//
// [JUCE 8.0.10 juce_audio_basics/utilities/juce_WindowedSincInterpolator.cpp, juce_LagrangeInterpolator.cpp,
//  juce_Interpolators.h -- not vendored by the reference (JuceLibraryCode/JuceHeader.h:16 links the module);
//  algorithm as recorded in SURVEY.md Appendix A.]
#include <algorithm>
#include <cmath>
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
