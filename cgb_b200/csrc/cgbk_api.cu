// C ABI of libcgbk.so: handles, host float64 motif math, table construction and
// the launch sequences of the kernels in cgbk_exact.cu / cgbk_fast.cu.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cgbk_internal.cuh"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local long long g_launch_info[3] = {0, 0, 0};

int cgbk_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cgbk_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return CGBK_OK;
    return cgbk_set_error(CGBK_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CUDA_TRY(expr)                                         \
    do {                                                       \
        int _rc = cgbk_check_cuda((expr), #expr);              \
        if (_rc != CGBK_OK) return _rc;                        \
    } while (0)

// optional event timing of the tiled kernel (bench.py roofline)
static thread_local int g_prof_on = 0;
static thread_local cudaEvent_t g_prof_ev[2] = {nullptr, nullptr};
static thread_local int g_prof_valid = 0;

extern "C" int cgbk_profile_enable(int on) {
    if (on && !g_prof_ev[0]) {
        CUDA_TRY(cudaEventCreate(&g_prof_ev[0]));
        CUDA_TRY(cudaEventCreate(&g_prof_ev[1]));
    }
    g_prof_on = on;
    g_prof_valid = 0;
    return CGBK_OK;
}

extern "C" int cgbk_profile_last_ms(float* ms) {
    if (!ms) return cgbk_set_error(CGBK_ERR_INVALID, "ms is NULL");
    if (!g_prof_valid) return cgbk_set_error(CGBK_ERR_INVALID, "no profiled scan recorded");
    CUDA_TRY(cudaEventSynchronize(g_prof_ev[1]));
    CUDA_TRY(cudaEventElapsedTime(ms, g_prof_ev[0], g_prof_ev[1]));
    return CGBK_OK;
}

// per-launch timing of a batched scan (cgbk_profile_enable(2)): one event pair per model's scan
// kernel, read back with cgbk_profile_batch_ms
static thread_local std::vector<cudaEvent_t>* g_prof_pool = nullptr;
static thread_local int g_prof_pairs = 0;

static cudaError_t prof_pair_event(int pair, int which, cudaStream_t st) {
    if (!g_prof_pool) g_prof_pool = new std::vector<cudaEvent_t>();
    const size_t idx = 2 * (size_t)pair + which;
    while (g_prof_pool->size() <= idx) {
        cudaEvent_t e;
        cudaError_t rc = cudaEventCreate(&e);
        if (rc != cudaSuccess) return rc;
        g_prof_pool->push_back(e);
    }
    return cudaEventRecord((*g_prof_pool)[idx], st);
}

extern "C" int cgbk_profile_batch_ms(float* ms, int n) {
    if (!ms || n < 0) return cgbk_set_error(CGBK_ERR_INVALID, "bad argument");
    if (n > g_prof_pairs) return cgbk_set_error(CGBK_ERR_INVALID, "only %d launches were timed", g_prof_pairs);
    for (int k = 0; k < n; ++k) {
        CUDA_TRY(cudaEventSynchronize((*g_prof_pool)[2 * k + 1]));
        CUDA_TRY(cudaEventElapsedTime(ms + k, (*g_prof_pool)[2 * k], (*g_prof_pool)[2 * k + 1]));
    }
    return CGBK_OK;
}

#define PROF_START(st) do { if (g_prof_on) CUDA_TRY(cudaEventRecord(g_prof_ev[0], (st))); } while (0)
#define PROF_STOP(st) do { if (g_prof_on) { CUDA_TRY(cudaEventRecord(g_prof_ev[1], (st))); g_prof_valid = 1; } } while (0)

// Multiprocessor count per device, queried once per device (lock-free: a race only repeats the
// query).  Launch sizes follow the device the handle lives on, not a literal.
int cgbk_device_sms(int device) {
    static std::atomic<int> cache[64];
    if (device < 0 || device >= 64) return 148;
    int v = cache[device].load(std::memory_order_relaxed);
    if (v <= 0) {
        int q = 0;
        if (cudaDeviceGetAttribute(&q, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || q <= 0) {
            cudaGetLastError();
            q = 148;
        }
        cache[device].store(q, std::memory_order_relaxed);
        v = q;
    }
    return v;
}

int cgbk_current_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
    return cgbk_device_sms(dev);
}

extern "C" int cgbk_version(void) { return CGBK_VERSION; }
extern "C" const char* cgbk_last_error(void) { return g_err; }

extern "C" int cgbk_last_launch_info(int64_t* out3) {
    if (!out3) return cgbk_set_error(CGBK_ERR_INVALID, "out3 is NULL");
    for (int i = 0; i < 3; ++i) out3[i] = g_launch_info[i];
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// host float64 motif math
//   Biopython 1.76 PositionSpecificScoringMatrix.mean/min/max and
//   thresholds.ScoreDistribution (pssm= branch) + threshold_patser, reached from
//   cgb/pssm_model.py:57-70.  Uniform background, as PSSMModel passes none.
// ---------------------------------------------------------------------------
static double py_floordiv(double vx, double wx) {   // CPython float_floor_div
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0) {
        if ((wx < 0) != (mod < 0)) { mod += wx; div -= 1.0; }
    }
    double floordiv;
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, vx / wx);
    }
    return floordiv;
}

static void pssm_stats(const double* lo, int m, double* ic, double* mn, double* mx) {
    double sx = 0.0, smin = 0.0, smax = 0.0;
    for (int i = 0; i < m; ++i) {
        double cmin = lo[i * 4], cmax = lo[i * 4];
        for (int k = 0; k < 4; ++k) {
            const double v = lo[i * 4 + k];
            if (v < cmin) cmin = v;
            if (v > cmax) cmax = v;
            if (isnan(v)) continue;
            if (isinf(v) && v < 0) continue;
            const double p = 0.25 * pow(2.0, v);
            sx += p * v;
        }
        smin += cmin; smax += cmax;
    }
    *ic = sx; *mn = smin; *mx = smax;
}

extern "C" int cgbk_patser_threshold(const double* lo, int m, int precision, double* threshold,
                                     double* ic_out, double* min_out, double* max_out) {
    if (!lo || m <= 0 || precision <= 0) return cgbk_set_error(CGBK_ERR_INVALID, "bad motif or precision");
    for (int i = 0; i < 4 * m; ++i)
        if (!isfinite(lo[i]))
            return cgbk_set_error(CGBK_ERR_INVALID, "non-finite log-odds at position %d", i / 4);
    double ic, pmin, pmax;
    pssm_stats(lo, m, &ic, &pmin, &pmax);
    if (ic_out) *ic_out = ic;
    if (min_out) *min_out = pmin;
    if (max_out) *max_out = pmax;
    if (!threshold) return CGBK_OK;
    const double min_score = pmin < 0.0 ? pmin : 0.0;
    const double interval = (pmax > 0.0 ? pmax : 0.0) - min_score;
    const long long n_points = (long long)precision * m;
    if (n_points < 2) return cgbk_set_error(CGBK_ERR_INVALID, "precision*m must be >= 2");
    const double step = interval / (double)(n_points - 1);
    auto index_diff = [&](double x) -> long long { return (long long)py_floordiv(x + 0.5 * step, step); };
    std::vector<double> dens(n_points, 0.0), next(n_points, 0.0);
    long long start = -index_diff(min_score);
    if (start < 0) start += n_points;            // Python negative index
    if (start < 0 || start >= n_points) return cgbk_set_error(CGBK_ERR_INVALID, "degenerate score range");
    dens[start] = 1.0;
    // Same additions in the same order as the reference's loops (letters outer, i ascending,
    // target index clamped to the grid), restricted to the support of the density -- adding the
    // +0.0 of an empty cell changes nothing -- and with the clamped ends peeled off so that the
    // interior is a plain shifted accumulate.
    long long s_lo = start, s_hi = start;
    for (int pos = 0; pos < m; ++pos) {
        std::fill(next.begin(), next.end(), 0.0);
        long long n_lo = n_points - 1, n_hi = 0;
        for (int k = 0; k < 4; ++k) {             // letters in alphabet order
            const long long d = index_diff(lo[pos * 4 + k]);
            const double* src = dens.data();
            double* dst = next.data();
            long long i = s_lo;
            const long long below_end = std::min(s_hi, -d - 1);            // i + d < 0  -> cell 0
            for (; i <= below_end; ++i) dst[0] += src[i] * 0.25;
            const long long in_end = std::min(s_hi, n_points - 1 - d);     // 0 <= i + d <= n - 1
            for (; i <= in_end; ++i) dst[i + d] += src[i] * 0.25;
            for (; i <= s_hi; ++i) dst[n_points - 1] += src[i] * 0.25;     // i + d > n - 1 -> last cell
            long long a = s_lo + d, b = s_hi + d;
            a = a < 0 ? 0 : (a > n_points - 1 ? n_points - 1 : a);
            b = b < 0 ? 0 : (b > n_points - 1 ? n_points - 1 : b);
            if (a < n_lo) n_lo = a;
            if (b > n_hi) n_hi = b;
        }
        dens.swap(next);
        s_lo = n_lo; s_hi = n_hi;
    }
    const double fpr = pow(2.0, -ic);
    long long i = n_points;
    double prob = 0.0;
    while (prob < fpr) {
        i -= 1;
        if (i < 0) return cgbk_set_error(CGBK_ERR_INVALID, "threshold search ran off the distribution");
        prob += dens[i];
    }
    *threshold = min_score + (double)i * step;
    return CGBK_OK;
}

// Motif batches (JASPAR-scale sets, BASELINE configs[4]): the same DP for n motifs, spread over
// host threads.  Each motif's arithmetic is exactly cgbk_patser_threshold's (one thread per
// motif, no shared accumulation), so the results are identical to n single calls.
extern "C" int cgbk_patser_threshold_batch(const double* logodds, const int* m, int n, int precision,
                                           int n_threads, double* threshold, double* ic,
                                           double* min_score, double* max_score, int* status) {
    if (!logodds || !m || n < 0) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument");
    std::vector<long long> off(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        if (m[i] <= 0) return cgbk_set_error(CGBK_ERR_INVALID, "motif %d has length %d", i, m[i]);
        off[i + 1] = off[i] + 4ll * m[i];
    }
    if (n_threads < 1) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads < 1) n_threads = 1;
    }
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    std::vector<int> rc(n, CGBK_OK);
    std::vector<std::string> msg(n);
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            double t = 0, a = 0, b = 0, c = 0;
            rc[i] = cgbk_patser_threshold(logodds + off[i], m[i], precision, threshold ? &t : nullptr, &a, &b, &c);
            if (rc[i] != CGBK_OK) msg[i] = g_err;               // thread-local text of this worker
            if (threshold) threshold[i] = t;
            if (ic) ic[i] = a;
            if (min_score) min_score[i] = b;
            if (max_score) max_score[i] = c;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    int first_bad = -1;
    for (int i = 0; i < n; ++i) {
        if (status) status[i] = rc[i];
        if (rc[i] != CGBK_OK && first_bad < 0) first_bad = i;
    }
    if (first_bad >= 0)
        return cgbk_set_error(rc[first_bad], "motif %d: %s", first_bad, msg[first_bad].c_str());
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------
extern "C" int64_t cgbk_bases2_words(int64_t cap_bases) { return cap_bases / 16; }

extern "C" int cgbk_pack_ascii(const uint8_t* d_ascii, int64_t n, int64_t dst_off, uint32_t* d_bases2,
                               uint32_t* d_amb, uint32_t* d_lower, int64_t cap_bases, void* stream) {
    if (n < 0 || dst_off < 0 || dst_off % CGBK_SEQ_ALIGN != 0)
        return cgbk_set_error(CGBK_ERR_INVALID, "dst_off must be a non-negative multiple of %d", CGBK_SEQ_ALIGN);
    int64_t padded = ((n + CGBK_SEQ_ALIGN - 1) / CGBK_SEQ_ALIGN) * CGBK_SEQ_ALIGN;
    if (padded == 0) padded = CGBK_SEQ_ALIGN;
    if (dst_off + padded > cap_bases)
        return cgbk_set_error(CGBK_ERR_INVALID, "sequence of %lld bases at offset %lld exceeds capacity %lld",
                              (long long)n, (long long)dst_off, (long long)cap_bases);
    if ((n > 0 && !d_ascii) || !d_bases2 || !d_amb || !d_lower)
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL device pointer");
    ON_DEVICE(cgbk_device_of(d_bases2));        // no handle: run where the planes live
    CUDA_TRY(launch_pack(d_ascii, n, dst_off, d_bases2, d_amb, d_lower, (cudaStream_t)stream));
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
// Row-major [256][NW] table -> the chunk-major layout fill_tables() (cgbk_fast.cu) reads with
// vector loads: all 16-byte chunks [c][e][4], then the 8-byte chunk [e][2], then the 4-byte chunk [e].
static void chunked_layout(const uint32_t* rows, int NW, uint32_t* out) {
    const int n16 = NW / 4, has8 = (NW % 4) >= 2 ? 1 : 0, has4 = NW % 2;
    for (int e = 0; e < 256; ++e) {
        for (int c = 0; c < n16; ++c)
            for (int k = 0; k < 4; ++k) out[(c * 256 + e) * 4 + k] = rows[e * NW + 4 * c + k];
        if (has8)
            for (int k = 0; k < 2; ++k) out[n16 * 1024 + e * 2 + k] = rows[e * NW + 4 * n16 + k];
        if (has4) out[n16 * 1024 + has8 * 512 + e] = rows[e * NW + NW - 1];
    }
}

static void group_partial(const double* tab, int m, int g, int e, double* out) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) {
        const int j = 4 * g + k;
        if (j < m) s += tab[j * 4 + ((e >> (2 * k)) & 3)];
    }
    *out = s;
}

// Sizes of the three parts of a model's device block: exact tables | int32 group tables |
// 16-bit filter tables (the last written per threshold by ensure_coarse)
static size_t model_exact_bytes() { return (sizeof(double) * 2 * CGBK_EXACT_DOUBLES + 255) / 256 * 256; }
static size_t model_precise_bytes(int G) { return (sizeof(uint32_t) * 256 * 2 * G + 255) / 256 * 256; }
static size_t model_coarse_bytes(int G) { return (sizeof(uint32_t) * 256 * G + 255) / 256 * 256; }
static size_t model_block_bytes(int G) { return model_exact_bytes() + model_precise_bytes(G) + model_coarse_bytes(G); }

static int validate_logodds(const double* lo, int m) {
    if (m < 1) return cgbk_set_error(CGBK_ERR_INVALID, "motif length must be >= 1");
    if (m > CGBK_MAX_MOTIF)
        return cgbk_set_error(CGBK_ERR_UNSUPPORTED, "motif length %d exceeds CGBK_MAX_MOTIF=%d", m, CGBK_MAX_MOTIF);
    for (int i = 0; i < 4 * m; ++i)
        if (!isfinite(lo[i]))
            return cgbk_set_error(CGBK_ERR_INVALID, "non-finite log-odds at position %d (zero probability "
                                  "without pseudocounts is not supported)", i / 4);
    return CGBK_OK;
}

// Host side of a model: float64 tables of both strands, column means, fixed-point scale; and the
// image of its device block (exact tables + int32 group tables; the filter part stays unwritten).
static void build_model_host(const double* lo, int m, int device, cgbk_model* M, unsigned char* image) {
    M->m = m; M->G = (m + 3) / 4; M->device = device; M->sms = cgbk_device_sms(device);
    double ic;
    pssm_stats(lo, m, &ic, &M->min_score, &M->max_score);
    double R = 0.0;
    for (int j = 0; j < m; ++j) {
        double colmax = 0.0;
        for (int k = 0; k < 4; ++k) {
            M->fwd[j * 4 + k] = lo[j * 4 + k];
            M->rc[j * 4 + k] = lo[(m - 1 - j) * 4 + (3 - k)];
            colmax = fmax(colmax, fabs(lo[j * 4 + k]));
        }
        R += colmax;
    }
    for (int j = 0; j < m; ++j) {
        // sum(pssm[l][pos] for l in 'ACGT') / 4, Python's left-to-right sum from int 0
        M->mean_f[j] = (((0 + M->fwd[j * 4]) + M->fwd[j * 4 + 1]) + M->fwd[j * 4 + 2] + M->fwd[j * 4 + 3]) / 4;
        M->mean_r[j] = (((0 + M->rc[j * 4]) + M->rc[j * 4 + 1]) + M->rc[j * 4 + 2] + M->rc[j * 4 + 3]) / 4;
    }
    // exact tables
    std::vector<double> ex(CGBK_EXACT_DOUBLES, 0.0);
    memcpy(ex.data(), M->fwd, sizeof(M->fwd));
    memcpy(ex.data() + CGBK_MAXM * 4, M->rc, sizeof(M->rc));
    memcpy(ex.data() + CGBK_MAXM * 8, M->mean_f, sizeof(M->mean_f));
    memcpy(ex.data() + CGBK_MAXM * 8 + CGBK_MAXM, M->mean_r, sizeof(M->mean_r));
    // int32 fixed-point tables: scale 2^kexp (at most 2^28) with |any window sum| < 2^30
    // (a little below 2^30, so that threshold - score of the fused scan cannot overflow either)
    int kexp = 28;
    if (R > 0) {
        const int k2 = (int)floor(log2((1073741824.0 - 8192.0) / R));
        if (k2 < kexp) kexp = k2;
    }
    // ... and with the whole score range (soft-max <= max + 1, default histogram edges at the
    // integers around it) below 2^31 units, so that (score - histogram edge) fits an int32
    while (kexp > 0 && (M->max_score - M->min_score + 4.0) * ldexp(1.0, kexp) >= 2147483648.0 - 1048576.0) --kexp;
    M->kexp = kexp;
    const double scale = ldexp(1.0, kexp);
    std::vector<long long> qf(4 * m), qr(4 * m);
    for (int i = 0; i < 4 * m; ++i) { qf[i] = llrint(M->fwd[i] * scale); qr[i] = llrint(M->rc[i] * scale); }
    // Each entry is off by up to half a unit.  Over a window the m errors add up, and when many
    // entries share one value (a motif built from a few sites: most log-odds are the
    // pseudocount's) they all lean the same way -- 12 x 0.45 units = 3e-7 on EVERY window of a
    // 12-bp motif at 2^-24, straight into the background mean (tools/fuzz_parity.py, seed 56).
    // The expected sum of the errors under a uniform base composition is therefore taken out
    // again, as a whole number of units, at position 0 of each strand.
    {
        double bias_f = 0.0, bias_r = 0.0;
        for (int i = 0; i < 4 * m; ++i) {
            bias_f += ((double)qf[i] - M->fwd[i] * scale) * 0.25;
            bias_r += ((double)qr[i] - M->rc[i] * scale) * 0.25;
        }
        const long long shift_f = llrint(bias_f), shift_r = llrint(bias_r);
        for (int b = 0; b < 4; ++b) { qf[b] -= shift_f; qr[b] -= shift_r; }
    }
    // Moment centre and unit (see make_stats_params): the centre is the middle of the score range,
    // the unit 2^-(kexp - sh) the finest the integer accumulators allow.  A body adds 32 re-centred
    // scores in an int32 (|d| < 2^26) and a tile up to 128 squares in an int64 (|d| < 2^28).  For
    // LexA (score range 70 bits, kexp 24) that is 2^-20; a motif with a narrow range gets an exact
    // sum.  (With a fixed 2^-18 a degenerate 2-bp motif -- 14 distinct scores, so the rounding does
    // not average out -- was off by 9e-7 in the mean, which its 3 000-window posteriors amplify
    // past 1e-6; found by tools/fuzz_parity.py.)
    {
        const double centre = 0.5 * (M->min_score + M->max_score);
        M->c_fix = (int)llrint(centre * scale);
        const double half_range_fix = (0.5 * (M->max_score - M->min_score) + 1.0) * scale + 1.0;  // soft-max <= max + 1
        int need = 0;
        while (need < 30 && ldexp(1.0, 26 + need) <= half_range_fix) ++need;
        M->sh = need;
        // The tables carry -(c_fix - half) at position 0 of each strand: the scan's strand sums are
        // then centred on c_fix and already hold the rounding half-unit, so a window's moment
        // term is one shift.  |sum - bias| <= range / 2 + half: further inside int32 than before.
        M->tab_bias = M->c_fix - (need > 0 ? (1 << (need - 1)) : 0);
        for (int b = 0; b < 4; ++b) { qf[b] -= M->tab_bias; qr[b] -= M->tab_bias; }
    }
    const int G = M->G;
    std::vector<uint32_t> precise((size_t)256 * 2 * G);
    for (int e = 0; e < 256; ++e)
        for (int g = 0; g < G; ++g) {
            long long sf = 0, sr = 0;
            for (int k = 0; k < 4; ++k) {
                const int j = 4 * g + k;
                if (j < m) { sf += qf[j * 4 + ((e >> (2 * k)) & 3)]; sr += qr[j * 4 + ((e >> (2 * k)) & 3)]; }
            }
            precise[(size_t)e * 2 * G + 2 * g] = (uint32_t)(int32_t)sf;
            precise[(size_t)e * 2 * G + 2 * g + 1] = (uint32_t)(int32_t)sr;
        }
    // ... followed by the same tables in ExactTables order: fr[j][b] = (fwd, rc), mean[j] = (f, r)
    ex.resize(2 * CGBK_EXACT_DOUBLES);
    for (int i = 0; i < CGBK_MAXM * 4; ++i) {
        ex[CGBK_EXACT_TAB_OFFSET + 2 * i] = ex[i];
        ex[CGBK_EXACT_TAB_OFFSET + 2 * i + 1] = ex[CGBK_MAXM * 4 + i];
    }
    for (int i = 0; i < CGBK_MAXM; ++i) {
        ex[CGBK_EXACT_TAB_OFFSET + CGBK_MAXM * 8 + 2 * i] = ex[CGBK_MAXM * 8 + i];
        ex[CGBK_EXACT_TAB_OFFSET + CGBK_MAXM * 8 + 2 * i + 1] = ex[CGBK_MAXM * 9 + i];
    }
    memcpy(image, ex.data(), sizeof(double) * 2 * CGBK_EXACT_DOUBLES);
    chunked_layout(precise.data(), 2 * G, (uint32_t*)(image + model_exact_bytes()));
    M->coarse_valid = 0;
}

static void bind_model_block(cgbk_model* M, unsigned char* block) {
    M->d_exact = (double*)block;
    M->d_precise = (uint32_t*)(block + model_exact_bytes());
    M->d_coarse = (uint32_t*)(block + model_exact_bytes() + model_precise_bytes(M->G));
}

extern "C" int cgbk_model_create(const double* lo, int m, int device, cgbk_model** out) {
    if (!lo || !out) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument");
    int rc = validate_logodds(lo, m);
    if (rc) return rc;
    ON_DEVICE(device);
    cgbk_model* M = (cgbk_model*)calloc(1, sizeof(cgbk_model));
    const int G = (m + 3) / 4;
    std::vector<unsigned char> image(model_block_bytes(G), 0);
    build_model_host(lo, m, device, M, image.data());
    unsigned char* block = nullptr;
    if (cudaMalloc((void**)&block, model_block_bytes(G)) != cudaSuccess) {
        free(M);
        return cgbk_set_error(CGBK_ERR_CUDA, "cudaMalloc of model tables failed");
    }
    M->owns_block = 1;
    bind_model_block(M, block);
    if (cudaMemcpy(block, image.data(), model_exact_bytes() + model_precise_bytes(G), cudaMemcpyHostToDevice) !=
        cudaSuccess) {
        cgbk_model_destroy(M);
        return cgbk_set_error(CGBK_ERR_CUDA, "upload of model tables failed");
    }
    *out = M;
    return CGBK_OK;
}

extern "C" void cgbk_model_destroy(cgbk_model* M) {
    if (!M) return;
    if (M->owns_block && M->d_exact) cudaFree(M->d_exact);   // one block: d_precise and d_coarse live inside it
    if (M->owns_block) free(M);                               // batch members are freed with their batch
}

// K models of one device: ONE device allocation and ONE upload (ordered on `stream`) instead of
// K synchronous allocate + copy pairs.  The reference builds one PSSMModel per genome in its
// go() loop (cgb/__init__.py:673-686, pssm_model.py:31-35); a motif set or a genome batch builds
// them all here.
struct cgbk_model_batch {
    int n;
    int device;
    cgbk_model* models;          // [n]
    unsigned char* d_block;      // all device tables
    unsigned char* h_image;      // pinned staging of the upload (kept until destroy)
};

extern "C" int cgbk_model_batch_create(const double* logodds, const int* m, int n, int device, void* stream,
                                       cgbk_model_batch** out) {
    if (!logodds || !m || !out || n < 1) return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    std::vector<size_t> lo_off(n + 1, 0), blk_off(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        int rc = validate_logodds(logodds + lo_off[i], m[i]);
        if (rc) {
            std::string msg = g_err;
            return cgbk_set_error(rc, "motif %d: %s", i, msg.c_str());
        }
        lo_off[i + 1] = lo_off[i] + 4 * (size_t)m[i];
        blk_off[i + 1] = blk_off[i] + model_block_bytes((m[i] + 3) / 4);
    }
    ON_DEVICE(device);
    cgbk_model_batch* B = (cgbk_model_batch*)calloc(1, sizeof(cgbk_model_batch));
    B->n = n; B->device = device;
    B->models = (cgbk_model*)calloc(n, sizeof(cgbk_model));
    if (cudaMallocHost((void**)&B->h_image, blk_off[n]) != cudaSuccess ||
        cudaMalloc((void**)&B->d_block, blk_off[n]) != cudaSuccess) {
        cgbk_model_batch_destroy(B);
        return cgbk_set_error(CGBK_ERR_CUDA, "allocation of %zu bytes of model tables failed", blk_off[n]);
    }
    // table construction is independent per model: spread over host threads when there are many
    auto build_range = [&](int lo_i, int hi_i) {
        for (int i = lo_i; i < hi_i; ++i) {
            build_model_host(logodds + lo_off[i], m[i], device, &B->models[i], B->h_image + blk_off[i]);
            bind_model_block(&B->models[i], B->d_block + blk_off[i]);
        }
    };
    int n_thr = n >= 64 ? (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (n_thr > 1) {
        cgbk_device_sms(device);                    // warm the cache before the threads read it
        std::vector<std::thread> pool;
        const int per = (n + n_thr - 1) / n_thr;
        for (int t = 0; t < n_thr; ++t)
            if (t * per < n) pool.emplace_back(build_range, t * per, std::min(n, (t + 1) * per));
        for (auto& th : pool) th.join();
    } else {
        build_range(0, n);
    }
    if (cudaMemcpyAsync(B->d_block, B->h_image, blk_off[n], cudaMemcpyHostToDevice, (cudaStream_t)stream) !=
        cudaSuccess) {
        cgbk_model_batch_destroy(B);
        return cgbk_set_error(CGBK_ERR_CUDA, "upload of model tables failed");
    }
    *out = B;
    return CGBK_OK;
}

extern "C" cgbk_model* cgbk_model_batch_get(cgbk_model_batch* B, int k) {
    if (!B || k < 0 || k >= B->n) return nullptr;
    return &B->models[k];
}

extern "C" int cgbk_model_batch_size(const cgbk_model_batch* B) { return B ? B->n : 0; }

extern "C" void cgbk_model_batch_destroy(cgbk_model_batch* B) {
    if (!B) return;
    if (B->d_block) cudaFree(B->d_block);
    if (B->h_image) cudaFreeHost(B->h_image);
    free(B->models);
    free(B);
}

// PSSMModel.score_seq for sequences of exactly the motif's length, on the host, in the
// reference's arithmetic (pssm_model.py:103-133 with Biopython's _pwm.c underneath): double
// accumulate in position order, float32 store, non-ACGT windows repaired by _calculate
// (lower-case letters are unknown there, pssm_model.py:96-100; a repaired single-window score
// stays float64, :117-127), soft-max log2(2**f + 2**r).  This is what score_self() needs (the
// ~10^2 sites of the collections, pssm_model.py:84-86): a few thousand table look-ups that
// are not worth a kernel launch.  `sites`: n_sites * m bytes back to back.
extern "C" int cgbk_score_sites_host(const double* lo, int m, const char* sites, int n_sites, double* fwd_out,
                                     double* rc_out, double* both_out) {
    if (!lo || (n_sites > 0 && !sites) || n_sites < 0) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument");
    int rc0 = validate_logodds(lo, m);
    if (rc0) return rc0;
    auto code = [](unsigned char ch, bool* lower) -> int {
        *lower = false;
        switch (ch) {
            case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
            case 'a': *lower = true; return 0; case 'c': *lower = true; return 1;
            case 'g': *lower = true; return 2; case 't': *lower = true; return 3;
            default: return -1;
        }
    };
    for (int s = 0; s < n_sites; ++s) {
        const unsigned char* w = (const unsigned char*)sites + (size_t)s * m;
        bool amb = false;
        for (int j = 0; j < m; ++j) { bool l; if (code(w[j], &l) < 0) amb = true; }
        double f = 0.0, r = 0.0;
        for (int j = 0; j < m; ++j) {
            bool lower;
            const int b = code(w[j], &lower);
            if (!amb) {
                f += lo[j * 4 + b];
                r += lo[(m - 1 - j) * 4 + (3 - b)];
            } else {
                // _calculate: anything that is not an upper-case ACGT key adds the column mean
                const bool unk = b < 0 || lower;
                const double mf = (((0 + lo[j * 4]) + lo[j * 4 + 1]) + lo[j * 4 + 2] + lo[j * 4 + 3]) / 4;
                const double* rcj = lo + (m - 1 - j) * 4;      // rc[j][b] = lo[m-1-j][3-b]
                const double mr = (((0 + rcj[3]) + rcj[2]) + rcj[1] + rcj[0]) / 4;
                f += unk ? mf : lo[j * 4 + b];
                r += unk ? mr : rcj[3 - b];
            }
        }
        const double fs = amb ? f : (double)(float)f, rs = amb ? r : (double)(float)r;
        if (fwd_out) fwd_out[s] = fs;
        if (rc_out) rc_out[s] = rs;
        if (both_out) both_out[s] = log(pow(2.0, fs) + pow(2.0, rs)) / log(2.0);
    }
    return CGBK_OK;
}

// 16-bit filter tables for one threshold (see cgbk_fast.cu header).
//   per strand: deficit d_g[e] = max_g - x_g[e]; a window can reach thr' = thr - margin
//   only if sum_g d_g <= Dd = M - thr'.  Quantised deficits floor(d*S), saturated at
//   Dq + 1, are stored as surpluses u = Qcap - dq, with a constant K folded into group 0
//   so that "sum >= 0x8000" <=> "sum of quantised deficits <= Dq".  No false negatives.
static const double kFilterMargin = 1e-4;   // covers the float32 rounding of the exact score

static void build_coarse_strand(const double* tab, int m, int G, double thr, uint32_t* halves /*[256][G]*/) {
    std::vector<double> x((size_t)256 * G), gmax(G, -INFINITY);
    for (int g = 0; g < G; ++g)
        for (int e = 0; e < 256; ++e) {
            group_partial(tab, m, g, e, &x[(size_t)e * G + g]);
            gmax[g] = fmax(gmax[g], x[(size_t)e * G + g]);
        }
    double Mx = 0.0;
    for (int g = 0; g < G; ++g) Mx += gmax[g];
    double Dd = Mx - (thr - kFilterMargin);
    if (!(Dd >= 0.0)) {   // threshold above the maximum score (or NaN): this strand can never hit
        for (size_t i = 0; i < (size_t)256 * G; ++i) halves[i] = 0;
        return;
    }
    if (Dd < 1e-9) Dd = 1e-9;
    long long target = (G > 1) ? (32768 - G) / (G - 1) : 32767;
    if (target > 32767) target = 32767;
    double S = (double)target / Dd;
    long long Dq = (long long)floor(Dd * S);
    while (Dq > target) { S = nextafter(S, 0.0); Dq = (long long)floor(Dd * S); }
    const long long Qcap = Dq + 1;
    const long long U = (long long)G * Qcap - Dq;
    const long long K = 0x8000 - U;     // >= 0 by the choice of target
    for (int e = 0; e < 256; ++e)
        for (int g = 0; g < G; ++g) {
            const double d = gmax[g] - x[(size_t)e * G + g];
            long long dq = (long long)floor(d * S);
            if (dq > Qcap || !(d * S < 4e9)) dq = Qcap;
            long long u = Qcap - dq + (g == 0 ? K : 0);
            halves[(size_t)e * G + g] = (uint32_t)u;
        }
}

static int ensure_coarse(cgbk_model* M, double thr, cudaStream_t st) {
    if (M->coarse_valid && M->coarse_thr == thr) return CGBK_OK;
    const int G = M->G;
    std::vector<uint32_t> hf((size_t)256 * G), hr((size_t)256 * G), packed((size_t)256 * G);
    build_coarse_strand(M->fwd, M->m, G, thr, hf.data());
    build_coarse_strand(M->rc, M->m, G, thr, hr.data());
    for (size_t i = 0; i < packed.size(); ++i) packed[i] = (hf[i] << 16) | (hr[i] & 0xffffu);
    std::vector<uint32_t> chunked(packed.size());
    chunked_layout(packed.data(), G, chunked.data());
    // stream-ordered: a previous scan on this stream may still be reading the table (pageable
    // source: the runtime stages the bytes before returning)
    CUDA_TRY(cudaMemcpyAsync(M->d_coarse, chunked.data(), sizeof(uint32_t) * chunked.size(),
                             cudaMemcpyHostToDevice, st));
    M->coarse_valid = 1;
    M->coarse_thr = thr;
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// sequence sets and tiling
// ---------------------------------------------------------------------------
extern "C" int64_t cgbk_seqset_table_ints(int n_seq) { return (2 + CGBK_N_TILINGS) * (int64_t)n_seq + CGBK_N_TILINGS; }

extern "C" int cgbk_seqset_create(const int64_t* seq_off, const int64_t* seq_len, int n_seq, int device,
                                  int64_t* d_table, void* stream, cgbk_seqset** out) {
    if (!seq_off || !seq_len || !out || n_seq < 1) return cgbk_set_error(CGBK_ERR_INVALID, "bad sequence table");
    if (!d_table) return cgbk_set_error(CGBK_ERR_INVALID, "d_table (device int64[cgbk_seqset_table_ints(n_seq)]) is NULL");
    for (int i = 0; i < n_seq; ++i) {
        if (seq_off[i] < 0 || seq_off[i] % CGBK_SEQ_ALIGN != 0 || seq_len[i] < 0)
            return cgbk_set_error(CGBK_ERR_INVALID, "sequence %d: offset must be a multiple of %d", i, CGBK_SEQ_ALIGN);
        if (i > 0 && seq_off[i] < seq_off[i - 1] + seq_len[i - 1])
            return cgbk_set_error(CGBK_ERR_INVALID, "sequence %d overlaps its predecessor", i);
    }
    ON_DEVICE(device);
    cgbk_seqset* S = (cgbk_seqset*)calloc(1, sizeof(cgbk_seqset));
    S->n_seq = n_seq; S->device = device;
    S->h_off = (int64_t*)malloc(sizeof(int64_t) * n_seq);
    S->h_len = (int64_t*)malloc(sizeof(int64_t) * n_seq);
    memcpy(S->h_off, seq_off, sizeof(int64_t) * n_seq);
    memcpy(S->h_len, seq_len, sizeof(int64_t) * n_seq);
    // caller-owned device block: off[n] | len[n] | CGBK_N_TILINGS x tile_prefix[n + 1]; no cudaMalloc / cudaFree
    // (both synchronise the device) on the per-genome path
    int64_t* block = d_table;
    S->d_off = block; S->d_len = block + n_seq;
    for (int k = 0; k < CGBK_N_TILINGS; ++k) {
        S->tiling[k].h_tile_prefix = (int64_t*)calloc(n_seq + 1, sizeof(int64_t));
        S->tiling[k].d_tile_prefix = block + 2 * n_seq + k * (n_seq + 1);
        S->tiling[k].cached_m = -1; S->tiling[k].cached_L = -1; S->tiling[k].cached_wcap = -1;
        S->tiling[k].n_tiles = 0;
    }
    cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
    if (n_seq > CGBK_INLINE_SEQS) {     // small tables travel in the kernel parameters instead
        e1 = cudaMemcpyAsync(S->d_off, S->h_off, sizeof(int64_t) * n_seq, cudaMemcpyHostToDevice,
                             (cudaStream_t)stream);
        e2 = cudaMemcpyAsync(S->d_len, S->h_len, sizeof(int64_t) * n_seq, cudaMemcpyHostToDevice,
                             (cudaStream_t)stream);
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        cgbk_seqset_destroy(S);
        return cgbk_set_error(CGBK_ERR_CUDA, "upload of sequence table failed");
    }
    *out = S;
    return CGBK_OK;
}

extern "C" void cgbk_seqset_destroy(cgbk_seqset* S) {
    if (!S) return;
    free(S->h_off); free(S->h_len);                           // the device table is the caller's
    for (int k = 0; k < CGBK_N_TILINGS; ++k) free(S->tiling[k].h_tile_prefix);
    free(S);
}

extern "C" int64_t cgbk_seqset_total_windows(const cgbk_seqset* S, int m) {
    int64_t t = 0;
    for (int i = 0; i < S->n_seq; ++i)
        if (S->h_len[i] - m + 1 > 0) t += S->h_len[i] - m + 1;
    return t;
}

// windows of sequence i that a scan with window cap `wcap` scores (wcap < 0: all of them).  The
// cap is how a chunk of a longer sequence is scanned with motifs of several lengths: the chunk
// carries the halo of the longest motif, and every motif stops at the chunk's own windows.
static inline int64_t windows_of(const cgbk_seqset* S, int i, int m, int64_t wcap) {
    int64_t nwin = S->h_len[i] - m + 1;
    if (nwin < 0) nwin = 0;
    if (wcap >= 0 && nwin > wcap) nwin = wcap;
    return nwin;
}

static int64_t count_tiles(const cgbk_seqset* S, int m, int L, int64_t wcap = -1) {
    const int64_t stride = 32ll * L - 32;
    int64_t t = 0;
    for (int i = 0; i < S->n_seq; ++i) {
        const int64_t nwin = windows_of(S, i, m, wcap);
        if (nwin > 0) t += (nwin + stride - 1) / stride;
    }
    return t;
}

static int ensure_tiles(cgbk_seqset* S, int slot, int m, int L, cudaStream_t st, int64_t wcap = -1) {
    cgbk_seqset::Tiling& T = S->tiling[slot];
    if (T.cached_m == m && T.cached_L == L && T.cached_wcap == wcap) return CGBK_OK;
    if (wcap >= 0 && S->n_seq > CGBK_INLINE_SEQS)
        return cgbk_set_error(CGBK_ERR_INVALID, "a window cap needs a set of at most %d sequences", CGBK_INLINE_SEQS);
    const int64_t stride = 32ll * L - 32;
    int64_t* prefix = T.h_tile_prefix;
    prefix[0] = 0;
    for (int i = 0; i < S->n_seq; ++i) {
        const int64_t nwin = windows_of(S, i, m, wcap);
        prefix[i + 1] = prefix[i] + (nwin > 0 ? (nwin + stride - 1) / stride : 0);
    }
    // stream-ordered after any earlier scan of this set on the same stream; small tables
    // travel in the kernel parameters instead
    if (S->n_seq > CGBK_INLINE_SEQS)
        CUDA_TRY(cudaMemcpyAsync(T.d_tile_prefix, prefix, sizeof(int64_t) * (S->n_seq + 1),
                                 cudaMemcpyHostToDevice, st));
    T.n_tiles = prefix[S->n_seq];
    T.cached_m = m; T.cached_L = L; T.cached_wcap = wcap;
    return CGBK_OK;
}

static int device_sms(int device) { return cgbk_device_sms(device); }

// Lane segment length L (32, 64, 96 or 128 bases; a warp tile owns 32 L - 32 windows).
// Tiles are dealt to the resident warps in rounds and a round costs about the same however
// full it is, so the tiling is picked to minimise rounds x (time of one round), with the round
// times measured on B200 under full load (tools/scan_scale.py over 0.1 Mbp - 1 Gbp,
// profiles/r1_scan_sizes.md): large inputs want the longest segment (least overhead per
// window), small ones the tiling whose last round is not mostly empty -- 5 Mbp is ONE round at
// L = 96 in the stats scan but three at L = 32.
static int choose_L(const cgbk_seqset* S, int m, int warps_total, const double (&round_us)[4], int64_t wcap = -1) {
    int best_L = 128;
    double best = 0.0;
    for (int L = 128; L >= 32; L -= 32) {
        const int64_t tiles = count_tiles(S, m, L, wcap);
        const int64_t rounds = (tiles + warps_total - 1) / warps_total;
        const double cost = (double)rounds * round_us[L / 32 - 1];
        if (L == 128 || cost < best * 0.999) { best = cost; best_L = L; }
    }
    return best_L;
}
// stats scan: 12 warps per SM; hit filter: 20 warps per SM.  Microseconds per round at L = 32, 64, 96, 128.
static const double kStatsRoundUs[4] = {4.6, 9.8, 12.6, 14.0};
static const double kHitsRoundUs[4] = {3.8, 6.3, 8.6, 10.7};
static int choose_L_stats(const cgbk_seqset* S, int m, int sms, int64_t wcap = -1) {
    return choose_L(S, m, sms * (fast_threads((m + 3) / 4, 1) / 32), kStatsRoundUs, wcap);
}
static int choose_L_hits(const cgbk_seqset* S, int m, int sms) { return choose_L(S, m, sms * 20, kHitsRoundUs); }

#define PARTIAL_CAP 1024
#define WORK_HEADER (64 + 4 * 8 * PARTIAL_CAP)

extern "C" int64_t cgbk_scan_workspace_bytes(const cgbk_seqset* S, int m, int64_t cand_cap) {
    if (!S) return 0;
    int64_t b = WORK_HEADER + 4 * count_tiles(S, m, 32) + 64 + 8 * (cand_cap > 0 ? cand_cap : 0);
    return (b + 255) / 256 * 256;
}

static int carve(const cgbk_seqset* S, int m, void* d_work, int64_t work_bytes, int64_t cand_cap, FastWork* w) {
    const int64_t need = cgbk_scan_workspace_bytes(S, m, cand_cap);
    if (!d_work || work_bytes < need)
        return cgbk_set_error(CGBK_ERR_INVALID, "workspace of %lld bytes is smaller than the %lld required",
                              (long long)work_bytes, (long long)need);
    unsigned char* p = (unsigned char*)d_work;
    w->counters = (long long*)p;
    w->partials = (unsigned long long*)(p + 64);
    w->partial_cap = PARTIAL_CAP;
    const int64_t dirty_bytes = (4 * count_tiles(S, m, 32) + 63) / 64 * 64;
    w->dirty_tiles = (unsigned int*)(p + WORK_HEADER);
    w->cands = (unsigned long long*)(p + WORK_HEADER + dirty_bytes);
    w->cand_cap = cand_cap;
    w->tile_begin = 0; w->tile_end = 0;          // set by the caller once the tiling is known
    w->ctr_slot = CGBK_CTR_INTERNAL_TILE; w->partial_base = 0;
    w->hits8 = nullptr; w->hit_cap = 0; w->hit_cursor = nullptr; w->fplane = nullptr;
    return CGBK_OK;
}

static GenomeView view_of(const cgbk_genome* g) {
    GenomeView v;
    v.bases2 = g->d_bases2; v.amb = g->d_amb; v.lower = g->d_lower; v.cap_bases = g->cap_bases;
    return v;
}

static int check_genome(const cgbk_genome* g) {
    if (!g || !g->d_bases2 || !g->d_amb || !g->d_lower || g->cap_bases <= 0 || g->cap_bases % CGBK_SEQ_ALIGN)
        return cgbk_set_error(CGBK_ERR_INVALID, "bad cgbk_genome (NULL plane or capacity not a multiple of %d)",
                              CGBK_SEQ_ALIGN);
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// exact entry points
// ---------------------------------------------------------------------------
extern "C" int cgbk_score_windows(const cgbk_model* M, const cgbk_genome* genome, const int64_t* d_start,
                                  const int64_t* d_out_off, int n_regions, int64_t total, int flags,
                                  float* d_f32, float* d_r32, double* d_both, double* d_f64, double* d_r64,
                                  void* stream) {
    if (!M) return cgbk_set_error(CGBK_ERR_INVALID, "NULL model");
    int rc = check_genome(genome);
    if (rc) return rc;
    if (n_regions < 0 || total < 0 || (total > 0 && (!d_start || !d_out_off)))
        return cgbk_set_error(CGBK_ERR_INVALID, "bad region table");
    ON_DEVICE(M->device);
    CUDA_TRY(launch_score_windows(M, view_of(genome), (const long long*)d_start, (const long long*)d_out_off,
                                  n_regions, total, flags, d_f32, d_r32, d_both, d_f64, d_r64,
                                  (cudaStream_t)stream));
    return CGBK_OK;
}

static SeqTable table_of(const cgbk_seqset* S, int slot);

extern "C" int cgbk_posteriors(const cgbk_model* M, const cgbk_genome* genome, const cgbk_seqset* set,
                               const int32_t* d_plane, const int64_t* d_start, const int32_t* d_nwin,
                               int n_regions, const double* d_ms_m, const double* d_ms_bg, double p_motif,
                               double alpha, double* d_post, void* stream) {
    if (!M) return cgbk_set_error(CGBK_ERR_INVALID, "NULL model");
    int rc = check_genome(genome);
    if (rc) return rc;
    if (n_regions < 0 || (n_regions > 0 && (!d_start || !d_nwin || !d_post || !d_ms_m || !d_ms_bg)))
        return cgbk_set_error(CGBK_ERR_INVALID, "bad promoter table");
    ON_DEVICE(M->device);
    PlaneView pv;
    pv.plane = nullptr; pv.L = 32; pv.D = 4 * (M->G - 1); pv.inv_scale = ldexp(1.0, -M->kexp);
    SeqTable seqs = {};
    if (d_plane) {
        if (!set) return cgbk_set_error(CGBK_ERR_INVALID, "a score plane needs the sequence set it was written for");
        if (set->tiling[CGBK_TILING_STATS].cached_m != M->m)
            return cgbk_set_error(CGBK_ERR_INVALID, "the score plane was not written for this motif length "
                                  "(run cgbk_background_stats with d_plane first)");
        pv.plane = d_plane; pv.L = set->tiling[CGBK_TILING_STATS].cached_L;
        seqs = table_of(set, CGBK_TILING_STATS);
    }
    CUDA_TRY(launch_posteriors(M, view_of(genome), seqs, pv, (const long long*)d_start, d_nwin, n_regions,
                               d_ms_m, d_ms_bg, p_motif, alpha, d_post, (cudaStream_t)stream));
    g_launch_info[0] = 1;
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// tiled scans
// ---------------------------------------------------------------------------
static SeqTable table_of(const cgbk_seqset* S, int slot) {
    const cgbk_seqset::Tiling& T = S->tiling[slot];
    const int64_t wcap = T.cached_wcap;
    SeqTable t;
    t.off = (const long long*)S->d_off; t.len = (const long long*)S->d_len;
    t.tile_prefix = (const long long*)T.d_tile_prefix; t.n_seq = S->n_seq; t.n_tiles = T.n_tiles;
    for (int i = 0; i < CGBK_INLINE_SEQS; ++i) {
        const bool have = i < S->n_seq && S->n_seq <= CGBK_INLINE_SEQS;
        t.off_v[i] = have ? S->h_off[i] : 0;
        // with a window cap the kernel sees the sequence end where the capped windows end
        t.len_v[i] = have ? (wcap >= 0 ? std::min<int64_t>(S->h_len[i], wcap + T.cached_m - 1) : S->h_len[i]) : 0;
        t.prefix_v[i] = have ? T.h_tile_prefix[i] : 0;
    }
    return t;
}

extern "C" int cgbk_scan_hits(cgbk_model* M, const cgbk_genome* genome, const cgbk_seqset* set,
                              double threshold, int flags, cgbk_hit* d_hits, int64_t hit_cap,
                              int64_t* d_counters, void* d_work, int64_t work_bytes, int64_t cand_cap,
                              void* stream) {
    if (!M || !set || !d_counters) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument");
    if (isnan(threshold)) return cgbk_set_error(CGBK_ERR_INVALID, "threshold is NaN");
    int rc = check_genome(genome);
    if (rc) return rc;
    if (hit_cap < 0 || cand_cap < 1 || (hit_cap > 0 && !d_hits))
        return cgbk_set_error(CGBK_ERR_INVALID, "bad output capacity");
    ON_DEVICE(M->device);
    cudaStream_t st = (cudaStream_t)stream;
    cgbk_seqset* S = const_cast<cgbk_seqset*>(set);
    FastWork w;
    rc = carve(S, M->m, d_work, work_bytes, cand_cap, &w);
    if (rc) return rc;
    const int threads = fast_threads(M->G, 0);
    const int sms = device_sms(M->device);
    // the tiling depends only on (set, m), not on the kernel variant: a score plane written by
    // the stats scan stays addressable after a hit scan of the same set
    const int L = choose_L_hits(S, M->m, sms);
    rc = ensure_tiles(S, CGBK_TILING_HITS, M->m, L, st);
    if (rc) return rc;
    rc = ensure_coarse(M, threshold, st);
    if (rc) return rc;
    const int64_t n_tiles = S->tiling[CGBK_TILING_HITS].n_tiles;
    w.tile_end = n_tiles;
    CUDA_TRY(cudaMemsetAsync(w.counters, 0, 64, st));
    int launches = 0;
    if (n_tiles > 0) {
        const int wpb = threads / 32;
        long long grid = (n_tiles + wpb - 1) / wpb;
        if (grid > sms) grid = sms;
        PROF_START(st);
        CUDA_TRY(launch_fast_filter(M, view_of(genome), table_of(S, CGBK_TILING_HITS), w, L, (int)grid, st));
        PROF_STOP(st);
        CUDA_TRY(launch_rescore(M, view_of(genome), table_of(S, CGBK_TILING_HITS), w, threshold, flags, d_hits, hit_cap, L, st));
        launches = 3;
    }
    CUDA_TRY(cudaMemcpyAsync(d_counters, w.counters, 64, cudaMemcpyDeviceToDevice, st));
    g_launch_info[0] = launches; g_launch_info[1] = n_tiles; g_launch_info[2] = L;
    return CGBK_OK;
}

// ints the score plane needs for any tiling of this set (the L = 32 tiling is the largest)
extern "C" int64_t cgbk_score_plane_ints(const cgbk_seqset* S, int m) {
    if (!S) return 0;
    int64_t best = 0;
    for (int L = 32; L <= 128; L += 32) {
        const int64_t need = count_tiles(S, m, L) * (int64_t)((L / 32 + 1) * 1024);
        if (need > best) best = need;
    }
    return best;
}

int cgbk_bg_scan_begin(cgbk_model* M, const cgbk_genome* genome, const cgbk_seqset* set, double hist_lo,
                       double hist_hi, int n_bins, unsigned long long* d_hist, double* d_stats,
                       int accumulate, int32_t* d_plane, int64_t plane_ints, void* d_work,
                       int64_t work_bytes, cudaStream_t st, BgScan* run) {
    if (!M || !set || !d_stats) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument");
    int rc = check_genome(genome);
    if (rc) return rc;
    if (d_hist) {
        if (n_bins < 1 || n_bins > CGBK_MAX_HIST_BINS || !(hist_hi > hist_lo))
            return cgbk_set_error(CGBK_ERR_INVALID, "histogram needs 1..%d bins and hi > lo", CGBK_MAX_HIST_BINS);
        if (!cgbk_hist_range_ok(M, hist_lo, hist_hi))
            return cgbk_set_error(CGBK_ERR_INVALID, "histogram range too wide for the fixed-point scores");
    } else {
        n_bins = 1; hist_lo = M->min_score - 1.0; hist_hi = M->max_score + 2.0;
    }
    run->moments_only = d_hist ? 0 : 1;      // no histogram wanted: the scan without one (MODE 2)
    cgbk_seqset* S = const_cast<cgbk_seqset*>(set);
    rc = carve(S, M->m, d_work, work_bytes, 0, &run->w);
    if (rc) return rc;
    run->M = M; run->g = view_of(genome); run->S = S; run->d_stats = d_stats;
    run->threads = fast_threads(M->G, 1);
    run->sms = device_sms(M->device);
    // the tiling depends only on (set, m), not on the kernel variant: a score plane written by
    // the stats scan stays addressable after a hit scan of the same set
    run->L = choose_L_stats(S, M->m, run->sms);
    rc = ensure_tiles(S, CGBK_TILING_STATS, M->m, run->L, st);
    if (rc) return rc;
    if (d_plane && plane_ints < S->tiling[CGBK_TILING_STATS].n_tiles * (int64_t)((run->L / 32 + 1) * 1024))
        return cgbk_set_error(CGBK_ERR_INVALID, "score plane of %lld ints is smaller than the %lld required",
                              (long long)plane_ints, (long long)cgbk_score_plane_ints(S, M->m));
    // Clearing the accumulators: ONE memset when the caller laid out
    // stats record | histogram | workspace (the workspace begins with the scan's counters), as
    // the pipeline and the Python wrapper do; separate ones otherwise.
    {
        unsigned char* rec = (unsigned char*)d_stats;
        const size_t rec_bytes = sizeof(double) * CGBK_BG_COUNT;
        const size_t hist_bytes = d_hist ? sizeof(unsigned long long) * n_bins : 0;
        const bool hist_adj = d_hist && (unsigned char*)d_hist == rec + rec_bytes;
        const bool ctr_adj = hist_adj && (unsigned char*)run->w.counters == rec + rec_bytes + hist_bytes;
        if (!accumulate && ctr_adj) {
            CUDA_TRY(cudaMemsetAsync(rec, 0, rec_bytes + hist_bytes + 64, st));
        } else {
            CUDA_TRY(cudaMemsetAsync(run->w.counters, 0, 64, st));
            if (!accumulate && hist_adj) {
                CUDA_TRY(cudaMemsetAsync(rec, 0, rec_bytes + hist_bytes, st));
            } else if (!accumulate) {
                if (d_hist) CUDA_TRY(cudaMemsetAsync(d_hist, 0, hist_bytes, st));
                CUDA_TRY(cudaMemsetAsync(d_stats, 0, rec_bytes, st));
            }
        }
    }
    run->sp = make_stats_params(M, hist_lo, hist_hi, n_bins, d_hist, d_plane);
    run->sp.d_stats = d_stats;
    run->next_tile = 0; run->n_ranges = 0; run->partials_used = 0; run->finalized = 0;
    g_launch_info[0] = 0; g_launch_info[1] = S->tiling[CGBK_TILING_STATS].n_tiles; g_launch_info[2] = run->L;
    return CGBK_OK;
}

int cgbk_bg_scan_range(BgScan* run, long long bases_ready, int last, cudaStream_t st) {
    cgbk_seqset* S = run->S;
    long long t_end = S->tiling[CGBK_TILING_STATS].n_tiles;
    if (!last) {
        // only single-sequence sets are fed in chunks: tile t reads bases [t * stride, t * stride + 32 L)
        if (S->n_seq != 1) return cgbk_set_error(CGBK_ERR_INVALID, "chunked scans take one sequence");
        const long long span = 32ll * run->L, stride = span - 32;
        const long long ready = bases_ready - S->h_off[0];
        t_end = ready < span ? 0 : (ready - span) / stride + 1;
        if (t_end > S->tiling[CGBK_TILING_STATS].n_tiles) t_end = S->tiling[CGBK_TILING_STATS].n_tiles;
    }
    if (t_end <= run->next_tile) {
        // nothing new to scan; a last call still owes the reduction of the earlier ranges
        if (last && run->n_ranges > 0 && !run->finalized) {
            CUDA_TRY(launch_bg_finalize_int(run->w.partials, run->partials_used, run->sp.c, run->sp.q, run->d_stats, 1, st));
            run->finalized = 1;
            g_launch_info[0] += 1;
        }
        return CGBK_OK;
    }
    if (run->n_ranges >= CGBK_MAX_TILE_RANGES)
        return cgbk_set_error(CGBK_ERR_INVALID, "more than %d tile ranges", CGBK_MAX_TILE_RANGES);
    const int wpb = run->threads / 32;
    long long grid = (t_end - run->next_tile + wpb - 1) / wpb;
    if (grid > run->sms) grid = run->sms;
    FastWork w = run->w;
    w.tile_begin = run->next_tile; w.tile_end = t_end;
    w.ctr_slot = CGBK_CTR_INTERNAL_TILE + run->n_ranges;
    w.partial_base = run->partials_used;
    StatsParams sp = run->sp;
    sp.finalize = last ? 1 : 0;      // this launch's last block reduces all partial moments
    PROF_START(st);
    if (run->moments_only)
        CUDA_TRY(launch_fast_moments(run->M, run->g, table_of(S, CGBK_TILING_STATS), w, run->L, (int)grid, sp, st));
    else
        CUDA_TRY(launch_fast_stats(run->M, run->g, table_of(S, CGBK_TILING_STATS), w, run->L, (int)grid, sp, st));
    PROF_STOP(st);
    if (last) run->finalized = 1;
    run->next_tile = t_end;
    run->n_ranges += 1;
    run->partials_used += (int)grid;
    g_launch_info[0] += 1;
    return CGBK_OK;
}

int cgbk_bg_scan_end(BgScan* run, cudaStream_t st) {
    if (!run->finalized) {
        // an empty scan (no tiles at all), or the caller never flagged a last range
        CUDA_TRY(launch_bg_finalize_int(run->w.partials, run->partials_used, run->sp.c, run->sp.q, run->d_stats, 1, st));
        run->finalized = 1;
        g_launch_info[0] += 1;
    }
    return CGBK_OK;
}

extern "C" int cgbk_background_stats(cgbk_model* M, const cgbk_genome* genome, const cgbk_seqset* set,
                                     double hist_lo, double hist_hi, int n_bins, unsigned long long* d_hist,
                                     double* d_stats, int accumulate, int32_t* d_plane, int64_t plane_ints,
                                     void* d_work, int64_t work_bytes, void* stream) {
    if (!M) return cgbk_set_error(CGBK_ERR_INVALID, "NULL model");
    ON_DEVICE(M->device);
    cudaStream_t st = (cudaStream_t)stream;
    BgScan run;
    int rc = cgbk_bg_scan_begin(M, genome, set, hist_lo, hist_hi, n_bins, d_hist, d_stats, accumulate, d_plane,
                                plane_ints, d_work, work_bytes, st, &run);
    if (rc) return rc;
    rc = cgbk_bg_scan_range(&run, 0, 1, st);
    if (rc) return rc;
    return cgbk_bg_scan_end(&run, st);
}

// Background over a region table (the reference's population: cgb/genome.py:215-218)
extern "C" int64_t cgbk_region_stats_workspace_bytes(void) { return 3 * 8 * PARTIAL_CAP * 2; }

extern "C" int cgbk_background_stats_regions(const cgbk_model* M, const cgbk_genome* genome, const int64_t* d_start,
                                             const int32_t* d_nwin, int n_regions, double hist_lo, double hist_hi,
                                             int n_bins, unsigned long long* d_hist, double* d_stats,
                                             int accumulate, void* d_work, int64_t work_bytes, void* stream) {
    if (!M || !d_stats) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument");
    int rc = check_genome(genome);
    if (rc) return rc;
    if (n_regions < 0 || (n_regions > 0 && (!d_start || !d_nwin)))
        return cgbk_set_error(CGBK_ERR_INVALID, "bad region table");
    if (d_hist && (n_bins < 1 || n_bins > CGBK_MAX_HIST_BINS || !(hist_hi > hist_lo)))
        return cgbk_set_error(CGBK_ERR_INVALID, "histogram needs 1..%d bins and hi > lo", CGBK_MAX_HIST_BINS);
    if (!d_work || work_bytes < cgbk_region_stats_workspace_bytes())
        return cgbk_set_error(CGBK_ERR_INVALID, "workspace smaller than cgbk_region_stats_workspace_bytes()");
    ON_DEVICE(M->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) {
        CUDA_TRY(cudaMemsetAsync(d_stats, 0, sizeof(double) * CGBK_BG_COUNT, st));
        if (d_hist) CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * n_bins, st));
    }
    if (!d_hist) { n_bins = 1; hist_lo = 0.0; hist_hi = 1.0; }
    CUDA_TRY(launch_region_stats(M, view_of(genome), (const long long*)d_start, d_nwin, n_regions, hist_lo, hist_hi,
                                 n_bins, d_hist, (double*)d_work, PARTIAL_CAP * 2, d_stats, 1, st));
    g_launch_info[0] = 2;
    return CGBK_OK;
}

extern "C" int cgbk_background_finalize(double* d_stats, void* stream) {
    if (!d_stats) return cgbk_set_error(CGBK_ERR_INVALID, "NULL stats");
    ON_DEVICE(cgbk_device_of(d_stats));
    CUDA_TRY(launch_bg_finalize(nullptr, 0, d_stats, 1, (cudaStream_t)stream));
    return CGBK_OK;
}

extern "C" int cgbk_background_finalize_batch(double* d_stats, int n_records, void* stream) {
    if (!d_stats || n_records < 0) return cgbk_set_error(CGBK_ERR_INVALID, "bad argument");
    ON_DEVICE(cgbk_device_of(d_stats));
    if (n_records > 0) CUDA_TRY(launch_bg_finalize_batch(d_stats, n_records, (cudaStream_t)stream));
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// motif batches: K models over one packed buffer in one call
// ---------------------------------------------------------------------------
// Workspace of a batched scan: moment partials | 4 per-tile int arrays | prefix-sum scratch |
// flag plane (one bit per window) | candidate record pairs.  Sized for any motif length and tiling
// of the set (the L = 32 tiling has the most tiles, the largest rows-per-window ratio).
struct BatchLayout {
    int64_t max_tiles, off_tiles, off_scan, scan_bytes, off_flags, flag_bytes, off_pairs, total;
};

static BatchLayout batch_layout(const cgbk_seqset* S, int64_t wcap, int with_hits, int64_t cand_cap) {
    BatchLayout b = {};
    int64_t o = 4 * 8 * PARTIAL_CAP;
    if (with_hits) {
        b.max_tiles = count_tiles(S, 1, 32, wcap);
        for (int L = 32; L <= 128; L += 32) {
            const int64_t need = count_tiles(S, 1, L, wcap) * (int64_t)((L / 32 + 1) * 128);
            if (need > b.flag_bytes) b.flag_bytes = need;
        }
        b.off_tiles = o;  o += (4 * 4 * (b.max_tiles + 1) + 255) / 256 * 256;
        b.scan_bytes = cgbk_scan_tmp_bytes(b.max_tiles + 1);
        b.off_scan = o;   o += (b.scan_bytes + 255) / 256 * 256;
        b.off_flags = o;  o += (b.flag_bytes + 255) / 256 * 256;
        b.off_pairs = o;  o += 16 * (cand_cap > 0 ? cand_cap : 0);
    }
    b.total = (o + 255) / 256 * 256;
    return b;
}

extern "C" int64_t cgbk_scan_batch_workspace_bytes(const cgbk_seqset* set, int64_t win_cap, int64_t cand_cap) {
    if (!set) return 0;
    return batch_layout(set, win_cap, cand_cap > 0 ? 1 : 0, cand_cap).total;
}

extern "C" int cgbk_scan_batch(cgbk_model* const* models, int n_models, const cgbk_genome* genome,
                               const cgbk_seqset* set, int64_t win_cap, const double* thresholds, int flags,
                               const double* hist_lo, const double* hist_hi, int n_bins,
                               unsigned long long* d_hist, double* d_stats, uint64_t* d_hits8, int64_t hit_cap,
                               int64_t* d_counters, void* d_work, int64_t work_bytes, int64_t cand_cap,
                               void* stream) {
    if (!models || n_models < 1 || !set || !d_stats || !d_counters)
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    int rc = check_genome(genome);
    if (rc) return rc;
    for (int k = 0; k < n_models; ++k) {
        if (!models[k]) return cgbk_set_error(CGBK_ERR_INVALID, "model %d is NULL", k);
        if (models[k]->device != models[0]->device)
            return cgbk_set_error(CGBK_ERR_INVALID, "model %d lives on another device", k);
        if (thresholds && isnan(thresholds[k])) return cgbk_set_error(CGBK_ERR_INVALID, "threshold %d is NaN", k);
    }
    if (!thresholds) cand_cap = 0;
    if (thresholds) {
        if (hit_cap < 0 || cand_cap < 1 || cand_cap >= (1ll << 30) || (hit_cap > 0 && !d_hits8))
            return cgbk_set_error(CGBK_ERR_INVALID, "bad hit / candidate capacity");
        if (genome->cap_bases >= (1ll << 31))
            return cgbk_set_error(CGBK_ERR_UNSUPPORTED, "compact hit records address buffers below 2^31 bases "
                                  "(shard the sequence, or use cgbk_scan_hits)");
    }
    if (d_hist && (n_bins < 1 || n_bins > CGBK_MAX_HIST_BINS))
        return cgbk_set_error(CGBK_ERR_INVALID, "histogram needs 1..%d bins", CGBK_MAX_HIST_BINS);
    const BatchLayout lay = batch_layout(set, win_cap, thresholds ? 1 : 0, cand_cap);
    const int64_t work_need = lay.total;
    if (!d_work || work_bytes < work_need)
        return cgbk_set_error(CGBK_ERR_INVALID, "workspace of %lld bytes is smaller than the %lld required",
                              (long long)work_bytes, (long long)work_need);
    ON_DEVICE(models[0]->device);
    cudaStream_t st = (cudaStream_t)stream;
    cgbk_seqset* S = const_cast<cgbk_seqset*>(set);
    const GenomeView gv = view_of(genome);
    // Clearing: ONE memset when the caller laid out stats | histograms | counters back to back
    // (as the Python wrapper does), separate ones otherwise.
    {
        unsigned char* a = (unsigned char*)d_stats;
        const size_t stats_bytes = sizeof(double) * CGBK_BG_COUNT * (size_t)n_models;
        const size_t hist_bytes = d_hist ? sizeof(unsigned long long) * (size_t)n_bins * n_models : 0;
        const size_t ctr_bytes = 64 * (size_t)(n_models + 1);
        const bool hist_adj = !d_hist || (unsigned char*)d_hist == a + stats_bytes;
        const bool ctr_adj = hist_adj && (unsigned char*)d_counters == a + stats_bytes + hist_bytes;
        if (ctr_adj) {
            CUDA_TRY(cudaMemsetAsync(a, 0, stats_bytes + hist_bytes + ctr_bytes, st));
        } else {
            CUDA_TRY(cudaMemsetAsync(d_stats, 0, stats_bytes, st));
            if (d_hist) CUDA_TRY(cudaMemsetAsync(d_hist, 0, hist_bytes, st));
            CUDA_TRY(cudaMemsetAsync(d_counters, 0, ctr_bytes, st));
        }
    }
    long long launches = 0, tiles_total = 0;
    PROF_START(st);
    for (int k = 0; k < n_models; ++k) {
        cgbk_model* M = models[k];
        const int L = choose_L_stats(S, M->m, M->sms, win_cap);
        rc = ensure_tiles(S, CGBK_TILING_BATCH, M->m, L, st, win_cap);
        if (rc) return rc;
        const int64_t n_tiles = S->tiling[CGBK_TILING_BATCH].n_tiles;
        double lo = hist_lo ? hist_lo[k] : floor(M->min_score), hi = hist_hi ? hist_hi[k] : ceil(M->max_score) + 1.0;
        unsigned long long* hist_k = d_hist ? d_hist + (size_t)k * n_bins : nullptr;
        if (hist_k) {
            if (!(hi > lo) || !cgbk_hist_range_ok(M, lo, hi))
                return cgbk_set_error(CGBK_ERR_INVALID, "model %d: bad histogram range", k);
        } else {
            lo = M->min_score - 1.0; hi = M->max_score + 2.0;
        }
        StatsParams sp = make_stats_params(M, lo, hi, hist_k ? n_bins : 1, hist_k, nullptr);
        sp.d_stats = d_stats + (size_t)k * CGBK_BG_COUNT;
        sp.finalize = 1;
        if (thresholds) set_stats_threshold(&sp, M, thresholds[k], (flags & CGBK_RC_REVSEQ_ORDER) ? 1 : 0);
        FastWork w;
        w.counters = (long long*)d_counters + (size_t)k * CGBK_CTR_COUNT;
        w.partials = (unsigned long long*)d_work;
        w.partial_cap = PARTIAL_CAP;
        w.dirty_tiles = nullptr;
        w.cands = nullptr;
        w.cand_cap = 0;
        w.fplane = thresholds ? (unsigned int*)((unsigned char*)d_work + lay.off_flags) : nullptr;
        int* tile_ints = (int*)((unsigned char*)d_work + lay.off_tiles);
        w.tile_cand = thresholds ? tile_ints : nullptr;
        w.tile_begin = 0; w.tile_end = n_tiles;
        w.ctr_slot = CGBK_CTR_INTERNAL_TILE; w.partial_base = 0;
        w.hits8 = thresholds ? (unsigned long long*)d_hits8 : nullptr;
        w.hit_cap = hit_cap;
        w.hit_cursor = (long long*)d_counters + (size_t)n_models * CGBK_CTR_COUNT;
        if (n_tiles > 0) {
            const int wpb = fast_threads(M->G, 1) / 32;
            long long grid = (n_tiles + wpb - 1) / wpb;
            if (grid > M->sms) grid = M->sms;
            const SeqTable tab = table_of(S, CGBK_TILING_BATCH);
            if (g_prof_on == 2) CUDA_TRY(prof_pair_event(k, 0, st));
            if (hist_k) CUDA_TRY(launch_fast_stats(M, gv, tab, w, L, (int)grid, sp, st));
            else CUDA_TRY(launch_fast_moments(M, gv, tab, w, L, (int)grid, sp, st));
            if (g_prof_on == 2) CUDA_TRY(prof_pair_event(k, 1, st));
            ++launches;
            if (thresholds) {
                SitePass sp2;
                sp2.tile_cand = tile_ints;
                sp2.tile_cand_off = tile_ints + (lay.max_tiles + 1);
                sp2.tile_hits = tile_ints + 2 * (lay.max_tiles + 1);
                sp2.tile_hit_off = tile_ints + 3 * (lay.max_tiles + 1);
                sp2.pairs = (unsigned long long*)((unsigned char*)d_work + lay.off_pairs);
                sp2.cand_cap = cand_cap;
                sp2.scan_tmp = (unsigned char*)d_work + lay.off_scan;
                sp2.scan_tmp_bytes = lay.scan_bytes;
                // (profiling: pair n_models + k brackets model k's site pass)
                if (g_prof_on == 2) CUDA_TRY(prof_pair_event(n_models + k, 0, st));
                CUDA_TRY(launch_site_pass(M, gv, tab, w, sp2, L, thresholds[k], flags, st));
                if (g_prof_on == 2) CUDA_TRY(prof_pair_event(n_models + k, 1, st));
                launches += 3;          // score, emit, finish (+ two library prefix sums)
            }
        } else {
            if (g_prof_on == 2) {
                CUDA_TRY(prof_pair_event(k, 0, st)); CUDA_TRY(prof_pair_event(k, 1, st));
                if (thresholds) { CUDA_TRY(prof_pair_event(n_models + k, 0, st)); CUDA_TRY(prof_pair_event(n_models + k, 1, st)); }
            }
            CUDA_TRY(launch_bg_finalize(nullptr, 0, sp.d_stats, 1, st));    // n = 0, mean = std = NaN
            ++launches;
        }
        tiles_total += n_tiles;
    }
    g_prof_pairs = g_prof_on == 2 ? (thresholds ? 2 * n_models : n_models) : 0;
    PROF_STOP(st);
    g_launch_info[0] = launches; g_launch_info[1] = tiles_total; g_launch_info[2] = 0;
    return CGBK_OK;
}

// ---------------------------------------------------------------------------
// split site records (see split_sites_kernel, cgbk_exact.cu)
// ---------------------------------------------------------------------------
extern "C" int64_t cgbk_split_sites_blocks(int64_t cap_bases) { return (cap_bases + 4095) / 4096; }

extern "C" int cgbk_split_sites(const uint64_t* d_hits8, int64_t n, int64_t cap_bases, uint16_t* d_local,
                                float* d_score, int32_t* d_block_start, void* stream) {
    if (n < 0 || cap_bases < 0 || !d_block_start || (n > 0 && (!d_hits8 || !d_local || !d_score)))
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or negative argument");
    if (n >= (1ll << 31)) return cgbk_set_error(CGBK_ERR_UNSUPPORTED, "more than 2^31 - 1 sites in one list");
    const int dev = cgbk_device_of(d_block_start);
    ON_DEVICE(dev);
    CUDA_TRY(launch_split_sites((const unsigned long long*)d_hits8, nullptr, 1, n, d_local, d_score, d_block_start,
                                cgbk_split_sites_blocks(cap_bases), cgbk_device_sms(dev), (cudaStream_t)stream));
    return CGBK_OK;
}

// the lists of n_models models back to back in ONE launch: d_cum (device, int64 [n_models + 1]) holds
// where each list starts; block_start is [n_models][blocks + 1], relative to the start of each list
extern "C" int cgbk_split_sites_batch(const uint64_t* d_hits8, const int64_t* d_cum, int n_models, int64_t n_total,
                                      int64_t cap_bases, uint16_t* d_local, float* d_score, int32_t* d_block_start,
                                      void* stream) {
    if (n_models < 1 || n_total < 0 || cap_bases < 0 || !d_cum || !d_block_start ||
        (n_total > 0 && (!d_hits8 || !d_local || !d_score)))
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or negative argument");
    const int dev = cgbk_device_of(d_block_start);
    ON_DEVICE(dev);
    CUDA_TRY(launch_split_sites((const unsigned long long*)d_hits8, (const long long*)d_cum, n_models, n_total, d_local,
                                d_score, d_block_start, cgbk_split_sites_blocks(cap_bases), cgbk_device_sms(dev),
                                (cudaStream_t)stream));
    return CGBK_OK;
}
