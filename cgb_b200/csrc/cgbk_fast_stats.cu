// MODE 1 of the tiled scan (cgbk_fast.cuh): soft-max histogram + moments (+ score plane,
// + hit candidates when a threshold is set), and the parameters of that epilogue.
#include <math.h>
#include <stdlib.h>

#include "cgbk_fast.cuh"

// Code variant of the stats modes for a lane segment of L bases: the peeled copies waste no
// epilogue on the previous lane's windows but are twice the code, which a warp that runs a
// single small tile pays for in cold instruction fetch.  CGBK_STATS_VARIANT=0|2 overrides.
int stats_variant(int L) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("CGBK_STATS_VARIANT");
        forced = e ? atoi(e) : -1;
    }
    if (forced == CGBK_VAR_UNIFIED || forced == CGBK_VAR_PEEL) return forced;
    return (L == 32 || L == 128) ? CGBK_VAR_PEEL : CGBK_VAR_UNIFIED;
}

static cudaError_t dispatch_stats(int G, const uint32_t* gtab, GenomeView g, SeqTable seqs, FastWork w, int m,
                                  int L, int grid, StatsParams sp, int device, cudaStream_t st) {
    const bool peel = stats_variant(L) == CGBK_VAR_PEEL;
#define CGBK_CASE(GG)                                                                                              \
    case GG:                                                                                                       \
        return peel ? launch_one<GG, 1, CGBK_VAR_PEEL>(gtab, g, seqs, w, m, L, grid, sp, device, st)               \
                    : launch_one<GG, 1, CGBK_VAR_UNIFIED>(gtab, g, seqs, w, m, L, grid, sp, device, st);
    switch (G) {
        CGBK_CASE(1) CGBK_CASE(2) CGBK_CASE(3) CGBK_CASE(4)
        CGBK_CASE(5) CGBK_CASE(6) CGBK_CASE(7) CGBK_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef CGBK_CASE
}

bool cgbk_hist_range_ok(const cgbk_model* mdl, double hist_lo, double hist_hi) {
    const double scale = ldexp(1.0, mdl->kexp), lim = 2147483648.0 - 65536.0;
    if (!(hist_hi > hist_lo) || (hist_hi - hist_lo) * scale >= lim) return false;
    const double lo_rel = hist_lo * scale - (double)mdl->tab_bias;            // lower edge as the kernel sees it
    const double s_min = mdl->min_score * scale - (double)mdl->tab_bias - 65536.0;
    const double s_max = (mdl->max_score + 1.0) * scale - (double)mdl->tab_bias + 65536.0;
    return fabs(lo_rel) < lim && s_max - lo_rel < lim && s_min - lo_rel > -lim;
}

StatsParams make_stats_params(const cgbk_model* mdl, double hist_lo, double hist_hi, int n_bins,
                              unsigned long long* d_hist, int* plane) {
    StatsParams sp;
    const double scale = ldexp(1.0, mdl->kexp);
    sp.neg_inv_scale = (float)(-1.0 / scale);
    sp.scale = (float)scale;
    sp.inv_scale_d = 1.0 / scale;
    {
        const int j = mdl->kexp < 23 ? 23 - mdl->kexp : 0;
        sp.magic = (float)ldexp(1.0, j);
        sp.magic_bits = (127 + j) << 23;
        sp.magic_mul = mdl->kexp > 23 ? (1 << (mdl->kexp - 23)) : 1;
    }
    // histogram lower edge as the table path sees it (its scores are off by -tab_bias)
    sp.lo_fix = (int)(llrint(hist_lo * scale) - (long long)mdl->tab_bias);
    const double range_fix = (hist_hi - hist_lo) * scale;
    // bin = floor(u * n_bins / range_fix) as mulhi(u, M) >> k (signed: a u a hair below zero gives
    // -1) with M = ceil(2^(32+k) n_bins / range_fix) scaled up to just below 2^31.  M overstates
    // the ratio by less than a relative 2^-30, i.e. by less than 2^-22 of a bin over 256 bins,
    // while two distinct integers u lie n_bins / range_fix >= 2^-19 of a bin apart for every
    // admissible range: rounded UP, every integer u gets exactly floor(u * n_bins / range_fix)
    // (rounded down, a u that sits exactly on a bin edge would fall into the bin below).
    double mul = 4294967296.0 * (double)n_bins / range_fix;
    int k = 0;
    while (k < 24 && mul * 2.0 < 2147483647.0) { mul *= 2.0; ++k; }
    mul = ceil(mul);
    if (mul > 2147483647.0) mul = 2147483647.0;
    sp.bin_mul = (unsigned int)mul;
    sp.bin_shift = k;
    sp.bin_c = -(long long)sp.lo_fix * (long long)sp.bin_mul;
    // the lean kernel drops the clamps of the bin index: allowed when the range covers every score
    // a window can take (soft-max <= max + 1) with a margin for the fixed-point rounding
    sp.lean_ok = (d_hist == nullptr) ||
                 (hist_lo <= mdl->min_score - 1e-3 && hist_hi >= mdl->max_score + 1.0 + 1e-3) ? 1 : 0;
    sp.n_bins = n_bins;
    sp.hist_words = n_bins <= 256 ? 32 : 1;
    sp.exact_tab = mdl->d_exact + CGBK_EXACT_TAB_OFFSET;
    sp.hist_lo_d = hist_lo;
    sp.inv_binw_d = (double)n_bins / (hist_hi - hist_lo);
    sp.d_stats = nullptr;
    sp.finalize = 0;
    sp.host_stats = nullptr;
    sp.host_hist = nullptr;
    // moment centre / unit / table bias: fixed per model (build_model_host in cgbk_api.cu)
    sp.c_fix = mdl->c_fix;
    sp.c = (double)sp.c_fix / scale;
    sp.sh = mdl->sh;
    sp.half = sp.sh > 0 ? (1 << (sp.sh - 1)) : 0;
    sp.tab_bias = mdl->tab_bias;
    sp.q = ldexp(1.0, -(mdl->kexp - sp.sh));
    sp.inv_q = ldexp(1.0, mdl->kexp - sp.sh);
    sp.d_hist = d_hist;
    sp.plane = plane;
    sp.thr_fix = CGBK_NO_THRESHOLD;
    sp.thr_m1 = 1 << 30;                 // above every score: no flag is ever set
    sp.thr = 0.0;
    sp.revseq = 0;
    return sp;
}

// Fused scan: candidate threshold in fixed point.  A hit is (double)(float)score >= thr with the
// exact float64 score; the fixed-point strand sum differs from score * 2^kexp by at most m
// units (m table entries rounded to half a unit each, plus the whole-unit bias shift of at
// most m / 2), and the float32 rounding can lift a score by half an ulp of float32 (< 2^-18
// for |score| < 64, relative 2^-24 beyond).  Everything at or above the bound is a candidate
// and is decided by the exact re-score: no false negatives.
void set_stats_threshold(StatsParams* sp, const cgbk_model* mdl, double thr, int revseq) {
    const double scale = ldexp(1.0, mdl->kexp);
    const double slack = 1e-5 + fabs(thr) * 1.2e-7;
    double t = floor((thr - slack) * scale) - (double)(2 * mdl->m + 2);
    // every window sum is below 2^30 - 4000 in magnitude (cgbk_api.cu picks kexp that way), so a
    // threshold clamped to that range decides the same and thr - 1 - score stays inside int32
    const double lim = 1073741824.0 - 4000.0;
    sp->thr = thr;
    sp->revseq = revseq ? 1 : 0;
    if (t > lim) {                       // above every score: no candidates
        sp->thr_fix = CGBK_NO_THRESHOLD;
        sp->thr_m1 = 1 << 30;
        return;
    }
    if (!(t > -lim)) t = -lim;           // at or below every score: everything is a candidate
    sp->thr_fix = (int)t;
    // compared against table-path sums, which are off by -tab_bias (|t - bias| < 2^31: t is within
    // +-2^30, the bias within half the score range)
    sp->thr_m1 = (int)((long long)t - 1 - (long long)mdl->tab_bias);
}

cudaError_t launch_fast_stats(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, int L,
                              int grid, const StatsParams& sp, cudaStream_t st) {
    if (!sp.plane && sp.lean_ok) return launch_fast_stats_lean(mdl, g, seqs, w, L, grid, sp, st);
    return dispatch_stats(mdl->G, mdl->d_precise, g, seqs, w, mdl->m, L, grid, sp, mdl->device, st);
}

// Reduction of the integer partials outside the scan (a chunked scan whose last range was empty,
// or a scan without tiles): one block.
__global__ void __launch_bounds__(256) bg_finalize_int_kernel(const unsigned long long* __restrict__ partials,
                                                              int n_partials, double centre, double q, double* stats,
                                                              int accumulate) {
    __shared__ unsigned long long scratch[128];
    unsigned long long tn = 0ull, tsd = 0ull, tlo = 0ull, thi = 0ull;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) {
        const unsigned long long* p = partials + 4ll * i;
        const unsigned long long l = p[2];
        tn += p[0]; tsd += p[1];
        tlo += l; thi += p[3] + (tlo < l ? 1ull : 0ull);
    }
    block_reduce_moments(tn, tsd, tlo, thi, scratch);
    if (threadIdx.x == 0) {
        double cn, s1, s2;
        moments_from_integers(tn, tsd, tlo, thi, centre, q, cn, s1, s2);
        if (accumulate) { cn += stats[CGBK_BG_N]; s1 += stats[CGBK_BG_SUM]; s2 += stats[CGBK_BG_SUMSQ]; }
        stats[CGBK_BG_N] = cn; stats[CGBK_BG_SUM] = s1; stats[CGBK_BG_SUMSQ] = s2;
        const double mean = cn > 0 ? s1 / cn : nan("");
        double var = cn > 0 ? s2 / cn - mean * mean : nan("");
        if (var < 0) var = 0;
        stats[CGBK_BG_MEAN] = mean;
        stats[CGBK_BG_STD] = sqrt(var);
    }
}

cudaError_t launch_bg_finalize_int(const unsigned long long* partials, int n_partials, double centre, double q,
                                   double* d_stats, int accumulate, cudaStream_t st) {
    bg_finalize_int_kernel<<<1, 256, 0, st>>>(partials, n_partials, centre, q, d_stats, accumulate);
    return cudaGetLastError();
}
