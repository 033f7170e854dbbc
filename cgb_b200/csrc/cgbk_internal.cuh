// Internal declarations shared by the cgbk translation units (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "cgbk.h"

#define CGBK_MAXM CGBK_MAX_MOTIF
#define CGBK_MAXG 8  // groups of 4 positions

// ---------------------------------------------------------------------------
// host-side handles
// ---------------------------------------------------------------------------
struct cgbk_model {
    int m;
    int G;       // ceil(m / 4): number of 4-mer groups
    int device;
    // host float64 tables, [pos][ACGT]
    double fwd[CGBK_MAXM * 4];
    double rc[CGBK_MAXM * 4];     // rc[j][b] = fwd[m-1-j][3-b]  (pssm_model.py:52-55)
    double mean_f[CGBK_MAXM];     // sum(pssm[l][pos] for l in 'ACGT') / 4  (pssm_model.py:100)
    double mean_r[CGBK_MAXM];
    double min_score, max_score;  // PSSM.min / PSSM.max
    // device copies
    double* d_exact;              // fwd[128] rc[128] mean_f[32] mean_r[32]
    uint32_t* d_precise;          // [256][2G] int32 fixed point (f, r interleaved per group)
    uint32_t* d_coarse;           // [256][G]  packed (f16 << 16 | r16), threshold dependent
    int kexp;                     // precise scale = 2^kexp
    // coarse-table cache key
    int coarse_valid;
    double coarse_thr;
};

struct cgbk_seqset {
    int n_seq;
    int device;
    int64_t* h_off;
    int64_t* h_len;
    int64_t* d_off;       // [n_seq]
    int64_t* d_len;       // [n_seq]
    // tile prefix tables are built per (m, lane segment) on demand and cached
    int cached_m, cached_L;
    int64_t* d_tile_prefix;  // [n_seq + 1]
    int64_t n_tiles;
};

#define CGBK_EXACT_DOUBLES (CGBK_MAXM * 4 * 2 + CGBK_MAXM * 2)

struct GenomeView {
    const uint32_t* bases2;
    const uint32_t* amb;
    const uint32_t* lower;
    long long cap_bases;
};

struct SeqTable {
    const long long* off;
    const long long* len;
    const long long* tile_prefix;
    int n_seq;
    long long n_tiles;
};

// per-thread error text + status helpers (cgbk_api.cu)
int cgbk_set_error(int code, const char* fmt, ...);
int cgbk_check_cuda(cudaError_t e, const char* what);

// ---------------------------------------------------------------------------
// launchers implemented in the kernel translation units
// ---------------------------------------------------------------------------
struct FastWork {      // carved out of the caller's workspace
    long long* counters;        // CGBK_CTR_COUNT
    unsigned int* dirty_tiles;  // n_tiles entries
    unsigned long long* cands;  // cand_cap entries
    long long cand_cap;
    double* partials;           // per-block (sum, sumsq, n) triples
    int partial_cap;            // triples available
};

cudaError_t launch_pack(const uint8_t* d_ascii, long long n, long long dst_off, uint32_t* b2,
                        uint32_t* amb, uint32_t* lower, cudaStream_t st);

cudaError_t launch_score_windows(const cgbk_model* mdl, GenomeView g, const long long* d_start,
                                 const long long* d_out_off, int n_regions, long long total,
                                 int flags, float* f32, float* r32, double* both, double* f64,
                                 double* r64, cudaStream_t st);

cudaError_t launch_posteriors(const cgbk_model* mdl, GenomeView g, const long long* d_start,
                              const int* d_nwin, int n_regions, const double* d_ms_m,
                              const double* d_ms_bg, double p_motif, double alpha, double* d_post,
                              cudaStream_t st);

cudaError_t launch_rescore(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                           double thr, int flags, cgbk_hit* hits, long long hit_cap, int L,
                           cudaStream_t st);

cudaError_t launch_dirty_stats(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                               int L, double hist_lo, double inv_binw, int n_bins,
                               unsigned long long* d_hist, int partial_base, cudaStream_t st);

cudaError_t launch_bg_finalize(const double* partials, int n_partials, double* d_stats,
                               int accumulate, cudaStream_t st);

cudaError_t launch_fast_filter(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                               int L, int grid, cudaStream_t st);

cudaError_t launch_fast_stats(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                              int L, int grid, double hist_lo, double inv_binw, int n_bins,
                              unsigned long long* d_hist, cudaStream_t st);

int fast_threads(int G, int mode);
