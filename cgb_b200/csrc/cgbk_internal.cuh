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
    int sms;     // multiprocessors of that device
    int owns_block;   // d_exact was allocated for this model alone (not part of a model batch)
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
    // moment centre (fixed point), moment rounding shift, and the constant tab_bias = c_fix - half
    // that the int32 group tables carry: every fixed-point strand sum comes out of the scan already
    // re-centred (and pre-rounded) for the moment sums, one addition per window less
    int c_fix, sh, tab_bias;
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
    // Tile prefix tables are built per (m, lane segment) on demand and cached, one slot per scan
    // kind: the stats scan (whose tiling also addresses the score plane) and the hit scan have
    // different numbers of resident warps and therefore different best tilings.
    struct Tiling {
        int cached_m, cached_L;
        int64_t cached_wcap;     // window cap the tiling was built with (-1: none)
        int64_t* d_tile_prefix;  // [n_seq + 1]
        int64_t* h_tile_prefix;  // [n_seq + 1] (staging for the async upload; inline copy source)
        int64_t n_tiles;
    } tiling[3];
};
#define CGBK_N_TILINGS 3
#define CGBK_TILING_STATS 0
#define CGBK_TILING_HITS 1
#define CGBK_TILING_BATCH 2   // cgbk_scan_batch: many motif lengths in a row; leaves the score plane's tiling alone

#define CGBK_EXACT_DOUBLES (CGBK_MAXM * 4 * 2 + CGBK_MAXM * 2)
// d_exact holds the plain tables (CGBK_EXACT_DOUBLES doubles) followed by the same numbers in
// ExactTables order (strands interleaved), which kernels without a staging step read in place
#define CGBK_EXACT_TAB_OFFSET CGBK_EXACT_DOUBLES

struct GenomeView {
    const uint32_t* bases2;
    const uint32_t* amb;
    const uint32_t* lower;
    long long cap_bases;
};

#define CGBK_CTR_INTERNAL_DONE 3   // blocks of the stats scan that have written their partial moments
#define CGBK_CTR_INTERNAL_TILE 4   // dynamic tile scheduler of the tiled scan (slots 4..7: one per
                                   // tile range of a chunked background scan)
#define CGBK_MAX_TILE_RANGES 4
#define CGBK_INLINE_SEQS 4   // sequence tables this small travel by value in the kernel parameters

struct SeqTable {
    const long long* off;
    const long long* len;
    const long long* tile_prefix;
    int n_seq;
    long long n_tiles;
    // copies for n_seq <= CGBK_INLINE_SEQS (constant-bank reads instead of global loads)
    long long off_v[CGBK_INLINE_SEQS];
    long long len_v[CGBK_INLINE_SEQS];
    long long prefix_v[CGBK_INLINE_SEQS];
};

// per-thread error text + status helpers (cgbk_api.cu)
int cgbk_set_error(int code, const char* fmt, ...);
int cgbk_check_cuda(cudaError_t e, const char* what);
// multiprocessor count of a device / of the current device (cached per device)
int cgbk_device_sms(int device);
int cgbk_current_sms();

// Every entry point that takes a handle runs on the handle's device: if another device is
// current it is switched for the duration of the call and restored on return.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            changed = err == cudaSuccess;
        }
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};
// device that owns a device pointer (entry points without a handle run where their buffers live);
// the current device when the pointer is unknown to the runtime
static inline int cgbk_device_of(const void* d_ptr) {
    cudaPointerAttributes at;
    int cur = 0;
    cudaGetDevice(&cur);
    if (!d_ptr || cudaPointerGetAttributes(&at, d_ptr) != cudaSuccess) { cudaGetLastError(); return cur; }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : cur;
}
#define ON_DEVICE(dev)                                                    \
    DeviceGuard _guard(dev);                                              \
    if (_guard.err != cudaSuccess) return cgbk_check_cuda(_guard.err, "selecting the handle's device")


// ---------------------------------------------------------------------------
// launchers implemented in the kernel translation units
// ---------------------------------------------------------------------------
struct StatsParams;

struct FastWork {      // carved out of the caller's workspace
    long long* counters;        // CGBK_CTR_COUNT
    unsigned int* dirty_tiles;  // n_tiles entries
    unsigned long long* cands;  // cand_cap entries
    long long cand_cap;
    unsigned long long* partials;   // per-block integer moments (n, sum d, sum d^2 low, sum d^2 high)
    int partial_cap;            // quadruples available
    // one launch of the tiled scan covers tiles [tile_begin, tile_end) and hands them out
    // through counters[ctr_slot]; its blocks write moment partials from partial_base on
    long long tile_begin, tile_end;
    int ctr_slot;
    int partial_base;
    // fused stats scan: compact hit list shared by the models of a batch.  Tiles with
    // non-ACGT bytes are scored exactly inside the scan and append their hits right there.
    unsigned long long* hits8;  // cgbk_hit8 records
    long long hit_cap;
    long long* hit_cursor;      // append position (one counter for the whole batch)
    // candidate flags of the fused scan: [tile][B + 1 rows][32 lanes] words, one bit per completed
    // window (row b < B: bit 31 - t is the window completed at body position t, i.e. tile-local
    // window lane * L + 32 b + t - D; row B: bit D - 1 - j is window lane * L + L - D + j)
    unsigned int* fplane;
    int* tile_cand;             // [n_tiles] flagged windows per tile (written with the flag plane)
};

// compact hit record (include/cgbk.h cgbk_hit8): gpos << 33 | minus << 32 | float32 score bits
__host__ __device__ inline unsigned long long cgbk_pack_hit8(long long gpos, int minus, unsigned int score_bits) {
    return ((unsigned long long)gpos << 33) | ((unsigned long long)(minus ? 1u : 0u) << 32) | score_bits;
}

#define CGBK_NO_THRESHOLD 0x7fffffff   // StatsParams.thr_fix of a scan without hit search

cudaError_t launch_pack(const uint8_t* d_ascii, long long n, long long dst_off, uint32_t* b2,
                        uint32_t* amb, uint32_t* lower, cudaStream_t st);

cudaError_t launch_score_windows(const cgbk_model* mdl, GenomeView g, const long long* d_start,
                                 const long long* d_out_off, int n_regions, long long total,
                                 int flags, float* f32, float* r32, double* both, double* f64,
                                 double* r64, cudaStream_t st);

// score plane as the posterior kernel reads it (plane == nullptr: re-score every window)
struct PlaneView {
    const int* plane;
    int L;              // lane segment of the tiling the plane was written with
    int D;              // 4 * (G - 1)
    double inv_scale;   // 2^-kexp
};

cudaError_t launch_posteriors(const cgbk_model* mdl, GenomeView g, SeqTable seqs, PlaneView pv,
                              const long long* d_start, const int* d_nwin, int n_regions,
                              const double* d_ms_m, const double* d_ms_bg, double p_motif,
                              double alpha, double* d_post, cudaStream_t st);

cudaError_t launch_rescore(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                           double thr, int flags, cgbk_hit* hits, long long hit_cap, int L,
                           cudaStream_t st);

cudaError_t launch_bg_finalize(const double* partials, int n_partials, double* d_stats,
                               int accumulate, cudaStream_t st);

// the same for the tiled scan's integer partials (n, sum d, sum d^2 as 128 bits; score = centre + d * q)
cudaError_t launch_bg_finalize_int(const unsigned long long* partials, int n_partials, double centre, double q,
                                   double* d_stats, int accumulate, cudaStream_t st);

cudaError_t launch_bg_finalize_batch(double* d_stats, int n_records, cudaStream_t st);

cudaError_t launch_split_sites(const unsigned long long* hits8, const long long* d_cum, int n_models, long long n,
                               unsigned short* local16, float* score32, int* block_start, long long n_blocks,
                               int sms, cudaStream_t st);

cudaError_t launch_region_stats(const cgbk_model* mdl, GenomeView g, const long long* d_start, const int* d_nwin,
                                int n_regions, double hist_lo, double hist_hi, int n_bins,
                                unsigned long long* d_hist, double* partials, int partial_cap, double* d_stats,
                                int accumulate, cudaStream_t st);

cudaError_t launch_fast_filter(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                               int L, int grid, cudaStream_t st);

// parameters of the soft-max / histogram / moment epilogue of the stats scan
struct StatsParams {
    float neg_inv_scale;   // -2^-kexp
    float scale;           // 2^kexp
    // soft-max correction to fixed point through a float add: magic = 2^max(0, 23 - kexp), its bit
    // pattern, and 2^max(0, kexp - 23) (see the stats epilogue of cgbk_fast.cuh)
    float magic;
    int magic_bits, magic_mul;
    double inv_scale_d;    // 2^-kexp
    int lo_fix;            // histogram lower edge, fixed point
    unsigned int bin_mul;  // ceil(2^(32 + bin_shift) * n_bins / range_fix), below 2^31 (signed high multiply)
    int bin_shift;
    long long bin_c;       // -lo_fix * bin_mul: bin = ((score * bin_mul + bin_c) >> 32) >> bin_shift
    int lean_ok;           // the histogram range covers every possible score: no clamping needed
    int n_bins;
    int hist_words;        // shared-memory words per histogram bin: 32 (one per lane, <= 256 bins) or 1
    int tab_bias;          // the strand sums of the table path are (true sum - tab_bias): c_fix - half
    int c_fix;             // moment centre, fixed point
    int sh;                // kexp - 18: moment rounding shift
    int half;              // 1 << (sh - 1) (0 when sh == 0)
    double c;              // centre in score units
    double q;              // moment unit in score units: 2^-(kexp - sh)
    double inv_q;          // 2^(kexp - sh)
    unsigned long long* d_hist;
    int* plane;            // optional score plane [tile][(B + 1) * 32 rows][32 lanes]
    // tiles with non-ACGT bytes are scored exactly inside the scan (device ExactTables)
    const void* exact_tab;
    double hist_lo_d;      // histogram lower edge and 1 / bin width in score units
    double inv_binw_d;
    // the block that finishes last in the launch flagged `finalize` reduces every per-block
    // partial moment written so far into d_stats (n, sum, sumsq accumulate; mean, std derived)
    double* d_stats;
    int finalize;
    // optional device-visible aliases of pinned host buffers: the finalizing block also
    // stores the finished record (and histogram) there, sparing the caller a D2H copy
    double* host_stats;
    unsigned long long* host_hist;
    // fused scan: windows whose fixed-point strand score reaches thr_fix are hit candidates
    // (CGBK_NO_THRESHOLD: none; thr_m1 = thr_fix - 1 clamped so that thr_m1 - score cannot
    // overflow); thr / revseq serve the exactly scored tiles
    int thr_fix;
    int thr_m1;
    int revseq;
    double thr;
};

// the histogram edges are usable with this model's fixed-point scores: (score - lower edge) of
// every score the table path can produce fits an int32, and so does the range
bool cgbk_hist_range_ok(const cgbk_model* mdl, double hist_lo, double hist_hi);

StatsParams make_stats_params(const cgbk_model* mdl, double hist_lo, double hist_hi, int n_bins,
                              unsigned long long* d_hist, int* plane);

void set_stats_threshold(StatsParams* sp, const cgbk_model* mdl, double thr, int revseq);

cudaError_t launch_fast_stats(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                              int L, int grid, const StatsParams& sp, cudaStream_t st);
// MODE 3 / 4: no score plane, no bin clamps (chosen by the two launchers above when they apply)
cudaError_t launch_fast_stats_lean(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                                   int L, int grid, const StatsParams& sp, cudaStream_t st);
cudaError_t launch_fast_moments_lean(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                                     int L, int grid, const StatsParams& sp, cudaStream_t st);
// the same without the histogram (MODE 2)
cudaError_t launch_fast_moments(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                                int L, int grid, const StatsParams& sp, cudaStream_t st);

// Fused scan, site pass: the flagged windows of the flag plane are re-scored exactly and written
// to the compact hit list in (position, + strand first) order.  Scratch arrays of n_tiles ints
// and cand_cap record pairs come carved out of the caller's workspace.
struct SitePass {
    const int* tile_cand;       // [n_tiles] from the scan
    int* tile_cand_off;         // [n_tiles] exclusive prefix of tile_cand
    int* tile_hits;             // [n_tiles] sites per tile
    int* tile_hit_off;          // [n_tiles] exclusive prefix of tile_hits
    unsigned long long* pairs;  // [2 * cand_cap] per flagged window: + strand record, - strand record (or ~0)
    long long cand_cap;
    void* scan_tmp;             // prefix-sum scratch
    long long scan_tmp_bytes;
};
cudaError_t launch_site_pass(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, const SitePass& sp,
                             int L, double thr, int flags, cudaStream_t st);
// exclusive prefix sum of n ints (cgbk_sort.cu, CUB)
long long cgbk_scan_tmp_bytes(long long n);
cudaError_t cgbk_exclusive_sum(const int* in, int* out, long long n, void* tmp, long long tmp_bytes, cudaStream_t st);

// score plane addressing shared by the writers (cgbk_fast.cu, dirty tiles) and the reader
// (posteriors): window wt of a tile with lane segment L = 32 * B and overlap D = 4 * (G - 1)
__host__ __device__ inline long long plane_index(long long tile, int wt, int L, int D) {
    const int B = L >> 5;
    const int lane = wt / L, within = wt - lane * L;
    const int row = within < L - D ? within + D : B * 32 + (within - (L - D));
    return tile * (long long)((B + 1) * 1024) + row * 32 + lane;
}

int fast_threads(int G, int mode);

// ---------------------------------------------------------------------------
// background scan in tile ranges (cgbk_api.cu), for callers that feed the genome in chunks
// (cgbk_pipeline.cu overlaps the H2D of chunk k + 1 with the scan of chunk k):
//   begin -> validates, clears the accumulators, fixes the tiling
//   range -> scans the not yet scanned tiles that lie entirely below bases_ready
//            (all remaining tiles when last != 0); at most CGBK_MAX_TILE_RANGES launches
//   end   -> makes sure the partial moments were reduced into d_stats (the last range's
//            last block does it; an empty scan needs the tiny finalize kernel)
// ---------------------------------------------------------------------------
struct BgScan {
    cgbk_model* M;
    GenomeView g;
    cgbk_seqset* S;
    FastWork w;
    StatsParams sp;
    int L, sms, threads;
    int moments_only;      // no histogram asked for: the scan runs without one (MODE 2)
    double* d_stats;
    long long next_tile;
    int n_ranges, partials_used, finalized;
};
int cgbk_bg_scan_begin(cgbk_model* M, const cgbk_genome* genome, const cgbk_seqset* set, double hist_lo,
                       double hist_hi, int n_bins, unsigned long long* d_hist, double* d_stats,
                       int accumulate, int32_t* d_plane, int64_t plane_ints, void* d_work,
                       int64_t work_bytes, cudaStream_t st, BgScan* out);
int cgbk_bg_scan_range(BgScan* run, long long bases_ready, int last, cudaStream_t st);
int cgbk_bg_scan_end(BgScan* run, cudaStream_t st);
