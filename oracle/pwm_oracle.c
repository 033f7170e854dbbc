/* C restatement of the reference CPU hot path (TEST ORACLE, not product).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (cgb_b200) never
 * links or calls it.
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   cgbo_calculate            Biopython 1.76 Bio/motifs/_pwm.c calculate(), reached
 *                             from cgb/pssm_model.py:114-115 (dependency not vendored;
 *                             restated from its published algorithm: PARITY UNPINNED).
 *   cgbo_score_seq            cgb/pssm_model.py:103-133 (score_seq) + :88-101 (_calculate)
 *   cgbo_scan_hits            cgb/genome.py:433-445 (one region of identify_sites)
 *   cgbo_binding_probability  cgb/binding_model.py:94-113
 *
 * Numeric conventions: see oracle/np_oracle.py header ("legacy64" soft-max of the
 * pinned numpy 1.x; float32 score vs float64 threshold).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int code_of(unsigned char c) {        /* _pwm.c switch */
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}
static inline int code_upper(unsigned char c) {     /* PSSM dict keys (pssm_model.py:97) */
    switch (c) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        default: return -1;
    }
}

/* _pwm.c calculate(): n windows, double accumulate j ascending, float32 store,
 * NaN when any byte of the window is outside ACGTacgt. */
void cgbo_calculate(const unsigned char* seq, int64_t len, const double* table, int m,
                    float* scores) {
    int64_t n = len - m + 1;
    for (int64_t i = 0; i < n; i++) {
        double score = 0.0;
        int ok = 1;
        for (int j = 0; j < m; j++) {
            int k = code_of(seq[i + j]);
            if (k >= 0) score += table[j * 4 + k];
            else ok = 0;
        }
        scores[i] = ok ? (float)score : NAN;
    }
}

/* pssm_model.py:88-101 */
static double calculate_ambiguous(const double* table, const unsigned char* site, int m) {
    double score = 0.0;
    for (int pos = 0; pos < m; pos++) {
        int k = code_upper(site[pos]);
        if (k >= 0) score += table[pos * 4 + k];
        else score += (0 + table[pos * 4 + 0] + table[pos * 4 + 1] + table[pos * 4 + 2]
                       + table[pos * 4 + 3]) / 4;
    }
    return score;
}

/* pssm_model.py:103-133.  f32/r32: per-strand float32 scores after the NaN repair
 * (what both=False returns); both64: the soft-max list (may be NULL).
 * When len == m the reference keeps repaired values as Python floats (not rounded
 * to float32); f64_single/r64_single receive those (may be NULL). */
void cgbo_score_seq(const unsigned char* seq, int64_t len, const double* fwd, const double* rc,
                    int m, float* f32, float* r32, double* both64,
                    double* f64_single, double* r64_single) {
    int64_t n = len - m + 1;
    cgbo_calculate(seq, len, fwd, m, f32);
    cgbo_calculate(seq, len, rc, m, r32);
    for (int64_t i = 0; i < n; i++) {
        double fs = (double)f32[i], rs = (double)r32[i];
        if (isnan(f32[i])) {
            fs = calculate_ambiguous(fwd, seq + i, m);
            rs = calculate_ambiguous(rc, seq + i, m);
            f32[i] = (float)fs; r32[i] = (float)rs;
            if (n != 1) { fs = (double)f32[i]; rs = (double)r32[i]; }
        }
        if (n == 1) {
            if (f64_single) *f64_single = fs;
            if (r64_single) *r64_single = rs;
        }
        if (both64) both64[i] = log(pow(2.0, fs) + pow(2.0, rs)) / log(2.0);
    }
}

static void revcomp(const unsigned char* s, int64_t n, unsigned char* out) {
    for (int64_t i = 0; i < n; i++) {
        unsigned char c = s[n - 1 - i], o = c;
        switch (c) {
            case 'A': o = 'T'; break; case 'C': o = 'G'; break;
            case 'G': o = 'C'; break; case 'T': o = 'A'; break;
            case 'a': o = 't'; break; case 'c': o = 'g'; break;
            case 'g': o = 'c'; break; case 't': o = 'a'; break;
            default: break;
        }
        out[i] = o;
    }
}

/* genome.py:433-445 for one region [0,len): returns the number of hits found;
 * writes at most cap of them (pos, strand, score) in the reference's emission
 * order (all + strand hits by position, then all - strand hits by position). */
int64_t cgbo_scan_hits(const unsigned char* seq, int64_t len, const double* fwd, const double* rc,
                       int m, double threshold, int64_t* pos, int32_t* strand, float* score,
                       int64_t cap) {
    int64_t n = len - m + 1, k = 0;
    if (n <= 0) return 0;
    float* f = (float*)malloc(sizeof(float) * n);
    float* r = (float*)malloc(sizeof(float) * n);
    cgbo_score_seq(seq, len, fwd, rc, m, f, r, NULL, NULL, NULL);
    for (int64_t i = 0; i < n; i++)
        if ((double)f[i] >= threshold) {
            if (k < cap) { pos[k] = i; strand[k] = 1; score[k] = f[i]; }
            k++;
        }
    unsigned char* rcs = (unsigned char*)malloc(len);
    revcomp(seq, len, rcs);
    cgbo_score_seq(rcs, len, fwd, rc, m, f, r, NULL, NULL, NULL);
    for (int64_t i = 0; i < n; i++) {            /* enumerate(reversed(rc_scores)) */
        float s = f[n - 1 - i];
        if ((double)s >= threshold) {
            if (k < cap) { pos[k] = i; strand[k] = -1; score[k] = s; }
            k++;
        }
    }
    free(f); free(r); free(rcs);
    return k;
}

/* scipy.stats.norm(loc, scale).pdf(x) = exp(-y*y/2)/sqrt(2*pi)/scale */
static inline double norm_pdf(double x, double loc, double scale) {
    double y = (x - loc) / scale;
    return exp(-(y * y) / 2.0) / sqrt(2 * M_PI) / scale;
}

/* binding_model.py:94-113 on precomputed soft-max scores */
double cgbo_posterior_from_scores(const double* s, int64_t n, double mu_m, double std_m,
                                  double mu_bg, double std_bg, double p_motif, double alpha) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double pm = norm_pdf(s[i], mu_m, std_m), pb = norm_pdf(s[i], mu_bg, std_bg);
        double lh_m = alpha * pm + (1 - alpha) * pb;
        acc += log(pb) - log(lh_m);
    }
    double ratio = exp(acc);
    return 1 / (1 + ratio * (1 - p_motif) / p_motif);
}

double cgbo_binding_probability(const unsigned char* seq, int64_t len, const double* fwd,
                                const double* rc, int m, double mu_m, double std_m,
                                double mu_bg, double std_bg, double p_motif, double alpha) {
    int64_t n = len - m + 1;
    if (n <= 0) return NAN;
    float* f = (float*)malloc(sizeof(float) * n);
    float* r = (float*)malloc(sizeof(float) * n);
    double* b = (double*)malloc(sizeof(double) * n);
    cgbo_score_seq(seq, len, fwd, rc, m, f, r, b, NULL, NULL);
    double p = cgbo_posterior_from_scores(b, n, mu_m, std_m, mu_bg, std_bg, p_motif, alpha);
    free(f); free(r); free(b);
    return p;
}

/* genome.py:215-226 run over ALL windows of one sequence (the GPU workload):
 * soft-max score of every window, accumulated as n, sum, sum of squares. */
void cgbo_background_moments(const unsigned char* seq, int64_t len, const double* fwd,
                             const double* rc, int m, double* out3) {
    int64_t n = len - m + 1;
    out3[0] = out3[1] = out3[2] = 0.0;
    if (n <= 0) return;
    float* f = (float*)malloc(sizeof(float) * n);
    float* r = (float*)malloc(sizeof(float) * n);
    double* b = (double*)malloc(sizeof(double) * n);
    cgbo_score_seq(seq, len, fwd, rc, m, f, r, b, NULL, NULL);
    double s1 = 0, s2 = 0;
    for (int64_t i = 0; i < n; i++) { s1 += b[i]; s2 += b[i] * b[i]; }
    out3[0] = (double)n; out3[1] = s1; out3[2] = s2;
    free(f); free(r); free(b);
}
