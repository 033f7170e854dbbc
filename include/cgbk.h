/* cgbk.h — C ABI of the B200-native cgb hot path (libcgbk.so).
 *
 * Scope: PSWM log-odds scanning of both DNA strands, the genome-wide background
 * score histogram/moments, thresholded hit lists and the per-promoter Bayesian
 * posterior of ErillLab/cgb.  Plain pointers and sizes only: no
 * torch types.  All d_* pointers are DEVICE pointers owned by the caller (e.g.
 * torch.Tensor.data_ptr()); `stream` is a cudaStream_t (NULL = default stream).
 * Every call is asynchronous on `stream` unless stated otherwise and returns a
 * status code; cgbk_last_error() returns the message of the calling thread's last
 * failure.  The library never frees caller memory.  State outside the handles: the
 * calling thread's error text and launch statistics, a per-device cache of the
 * multiprocessor count and of the kernels' shared-memory attribute (set once per device
 * under a lock), and the regulation pipeline's per-thread copy stream.  Entry points that
 * take a handle run on the handle's device: if another device is current it is selected for
 * the duration of the call and restored on return.
 *
 * Each entry point names the reference interface (file:line under
 * /root/reference) it replaces.
 */
#ifndef CGBK_H
#define CGBK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGBK_VERSION 211

/* status codes; the Python layer maps them to the reference's exception types */
#define CGBK_OK 0
#define CGBK_ERR_INVALID 1      /* bad argument            -> ValueError   */
#define CGBK_ERR_CUDA 2         /* CUDA runtime failure    -> RuntimeError */
#define CGBK_ERR_UNSUPPORTED 3  /* e.g. motif longer than CGBK_MAX_MOTIF   */

#define CGBK_MAX_MOTIF 32       /* longest PSWM the kernels accept (bases) */
#define CGBK_SEQ_ALIGN 512      /* sequence starts are multiples of this (bases) */
#define CGBK_MAX_HIST_BINS 4096

typedef struct cgbk_model cgbk_model;    /* opaque: one PSSM (log-odds) on one device */
typedef struct cgbk_model_batch cgbk_model_batch;  /* opaque: K models in one device block */
typedef struct cgbk_seqset cgbk_seqset;  /* opaque: sequence table of a packed buffer */

/* Packed genome buffer ("Chromid.sequence", cgb/chromid.py:57-60, as HBM planes).
 * Coordinates are base offsets into one concatenated, padded buffer; sequence k
 * occupies [seq_off[k], seq_off[k]+seq_len[k]) with seq_off[k] % CGBK_SEQ_ALIGN == 0. */
typedef struct {
    const uint32_t* d_bases2;  /* 2 bits/base, 16 bases/word, base p at bits 2*(p%16); A0 C1 G2 T3 */
    const uint32_t* d_amb;     /* 1 bit/base, 32 bases/word: byte outside ACGTacgt (Biopython _pwm.c default:) */
    const uint32_t* d_lower;   /* 1 bit/base: lower-case acgt (matters only to pssm_model.py:96-100) */
    int64_t cap_bases;         /* capacity of the three planes in bases (multiple of CGBK_SEQ_ALIGN) */
} cgbk_genome;

/* One thresholded site: cgb/genome.py:22 `Site` minus the Python objects. */
typedef struct {
    int64_t gpos;    /* window start, offset into the packed buffer */
    int32_t strand;  /* +1 / -1 (genome.py:438,445) */
    float score;     /* float32 PSSM score of that strand, as Biopython's calculate() stores it */
} cgbk_hit;

/* Compact site record of the batched scan: (gpos << 33) | (minus strand ? 1 << 32 : 0) |
 * float32 score bits.  Sorting the integers sorts by (position, + strand first).  Addresses
 * packed buffers below 2^31 bases (a rank's shard of a longer sequence). */
typedef uint64_t cgbk_hit8;
#define CGBK_HIT8_GPOS(h) ((int64_t)((h) >> 33))
#define CGBK_HIT8_STRAND(h) ((((h) >> 32) & 1u) ? -1 : 1)
#define CGBK_HIT8_SCORE_BITS(h) ((uint32_t)((h) & 0xffffffffu))

/* Counters written by cgbk_scan_hits (device int64[8]). */
#define CGBK_CTR_CANDIDATES 0   /* windows passed by the 16-bit filter */
#define CGBK_CTR_HITS 1         /* hits found (may exceed hit_cap: re-run with a larger buffer) */
#define CGBK_CTR_DIRTY_TILES 2  /* tiles holding non-ACGT bytes, scored by the exact kernel */
#define CGBK_CTR_COUNT 8

/* Layout of the device double[] written by cgbk_background_stats. */
#define CGBK_BG_N 0        /* number of windows */
#define CGBK_BG_SUM 1      /* sum of soft-max scores */
#define CGBK_BG_SUMSQ 2    /* sum of squares */
#define CGBK_BG_MEAN 3     /* np.mean(bg_scores)          (binding_model.py:87) */
#define CGBK_BG_STD 4      /* np.std(bg_scores), ddof = 0 (binding_model.py:87) */
#define CGBK_BG_COUNT 8

/* flags */
#define CGBK_RC_MATRIX_ORDER 0  /* minus strand = rev_comp_pssm.calculate(seq)   (pssm_model.py:115) */
#define CGBK_RC_REVSEQ_ORDER 1  /* minus strand = pssm.calculate(revcomp(seq))   (genome.py:440-445) */
#define CGBK_SINGLE_WINDOW_F64 2 /* len(seq)==m quirk: repaired scores stay float64 (pssm_model.py:117-127) */

int cgbk_version(void);
const char* cgbk_last_error(void);

/* ---- host float64 motif math (no GPU) ------------------------------------- */

/* PSSMModel.patser_threshold (cgb/pssm_model.py:62-70): Biopython
 * pssm.distribution(precision).threshold_patser() on a [m][4] (A,C,G,T) log-odds
 * table with the uniform background; also returns pssm.mean() (IC, :57-60),
 * pssm.min and pssm.max.  Any output pointer may be NULL. */
int cgbk_patser_threshold(const double* logodds, int m, int precision,
                          double* threshold, double* ic, double* min_score, double* max_score);

/* The same for a batch of n motifs (PSSMModel.IC / .patser_threshold of every model of a
 * JASPAR-scale set, cgb/pssm_model.py:57-70): logodds holds the [m_i][4] tables back to back,
 * outputs are arrays of n (any may be NULL).  Motifs are spread over n_threads host threads
 * (<= 0: all hardware threads); each motif's arithmetic is that of the single call, so the
 * numbers are identical.  status[i] (optional) receives the per-motif status; the call returns
 * the first failure, if any. */
int cgbk_patser_threshold_batch(const double* logodds, const int* m, int n, int precision, int n_threads,
                                double* threshold, double* ic, double* min_score, double* max_score,
                                int* status);

/* ---- packing ("str(record.seq)" -> HBM planes) ----------------------------- */

/* Words (uint32) the bases2 plane needs for cap_bases bases; amb/lower need half. */
int64_t cgbk_bases2_words(int64_t cap_bases);

/* ASCII bytes d_ascii[0..n) -> planes at base offset dst_off (multiple of
 * CGBK_SEQ_ALIGN); the tail up to the next multiple of CGBK_SEQ_ALIGN is zeroed.
 * Replaces Chromid.sequence / subsequence (cgb/chromid.py:57-60,92-97) and the
 * per-call Seq(...) of pssm_model.py:113. */
int cgbk_pack_ascii(const uint8_t* d_ascii, int64_t n, int64_t dst_off,
                    uint32_t* d_bases2, uint32_t* d_amb, uint32_t* d_lower,
                    int64_t cap_bases, void* stream);

/* The same packing on the HOST (AVX2 when the CPU has it, host threads for large inputs;
 * n_threads 0 = as many as the machine has, at most 16): ascii[0..n_bases) -> three host planes
 * of cgbk_packed_capacity(n_bases) bases (n_bases rounded up to CGBK_SEQ_ALIGN, at least one
 * block; the tail is zero), bit for bit what cgbk_pack_ascii writes on the device.  n_amb /
 * n_lower (optional) receive the number of set bits of the two flag planes: a replicon with
 * none of either can travel as its bases2 plane alone (0.25 byte per base).  A sequence kept or
 * stored in this form (the "packed genome file" of PackedGenome.save) is what
 * cgbk_regulation_pipeline_packed takes.  No GPU is involved. */
int64_t cgbk_packed_capacity(int64_t n_bases);
int cgbk_pack_host(const uint8_t* ascii, int64_t n_bases, uint32_t* bases2, uint32_t* amb,
                   uint32_t* lower, int64_t cap_bases, int n_threads, int64_t* n_amb,
                   int64_t* n_lower);

/* ---- GenBank feature table (host only) ---------------------------------------- */

/* Structural index of a GenBank feature table (the text between the FEATURES header line and
 * ORIGIN): what Biopython's SeqIO.read hands the reference as record.features
 * (cgb/chromid.py:32-35,99-122), found in one C pass so that the host never walks the table line
 * by line.  Offsets are relative to `block`.
 *   feat[k][10] = key_off, key_len, loc_off, loc_len, q_first, q_count, start, end, strand, compound
 *                 (start/end 0-based half-open, fuzzy ends dropped; compound = join/order/bond or
 *                 unparsable: not one contiguous span)
 *   qual[j][5]  = key_off, key_len, val_off, val_len, simple  (raw value, right-stripped:
 *                 continuation lines and quotes still inside; val_len = -1 for a flag qualifier
 *                 such as /pseudo; simple = 1 when the value is one line, quoted, with no quote
 *                 inside: its text is then raw[1 .. val_len - 1))
 * Nothing is written beyond the capacities; n_feat / n_qual always return the full counts. */
int cgbk_genbank_index(const char* block, int64_t n, int32_t* feat, int32_t feat_cap, int32_t* qual,
                       int32_t qual_cap, int32_t* n_feat, int32_t* n_qual);
/* (the index stops at the first line that starts with CONTIG, BASE COUNT, WGS or COMMENT in column 1:
 * the sections that may follow the table) */

/* ORIGIN block of a GenBank record -> sequence bytes (host only; replaces str(record.seq) of
 * SeqIO.read, cgb/chromid.py:32-35,57-60).  `from` = offset of the first line after the "ORIGIN"
 * line.  Every byte that is not one of " 0123456789\t\r\n" is kept, ASCII lower case folded to upper,
 * up to a line starting with "//" or the end of the text.  n_bases = bytes written to `out` (capacity
 * `cap`: CGBK_ERR_INVALID when it does not fit), end_pos = offset of the "//" line (or n). */
int cgbk_genbank_origin(const char* text, int64_t n, int64_t from, uint8_t* out, int64_t cap,
                        int64_t* n_bases, int64_t* end_pos);

/* ---- handles ---------------------------------------------------------------- */

/* logodds: host [m][4] float64, columns A,C,G,T (PSSMModel.pssm, pssm_model.py:47-50).
 * The reverse-complement table (pssm_model.py:52-55) is derived inside.  Synchronous. */
int cgbk_model_create(const double* logodds, int m, int device, cgbk_model** out);
void cgbk_model_destroy(cgbk_model* model);

/* n models at once (one PSSMModel per genome in the reference's go() loop,
 * cgb/__init__.py:673-686, pssm_model.py:31-35; or a JASPAR-scale motif set): logodds holds the
 * [m_i][4] tables back to back.  ONE device allocation and ONE upload ordered on `stream`
 * instead of n synchronous allocate + copy pairs.  cgbk_model_batch_get returns a borrowed
 * handle, valid until the batch is destroyed (cgbk_model_destroy on it is a no-op). */
int cgbk_model_batch_create(const double* logodds, const int* m, int n, int device, void* stream,
                            cgbk_model_batch** out);
cgbk_model* cgbk_model_batch_get(cgbk_model_batch* batch, int k);
int cgbk_model_batch_size(const cgbk_model_batch* batch);
void cgbk_model_batch_destroy(cgbk_model_batch* batch);

/* PSSMModel.score_self (cgb/pssm_model.py:84-86): score_seq of n_sites sequences of exactly m
 * bases each (sites: n_sites * m bytes back to back), on the HOST in the reference's
 * arithmetic -- the ~10^2 sites of a model are a few thousand table look-ups, not worth a
 * launch.  Outputs (each may be NULL): forward / minus-strand score as score_seq(both=False)
 * would return it for a single window (float32-rounded unless the ambiguity repair of
 * pssm_model.py:88-101 touched it) and the soft-max log2(2**f + 2**r). */
int cgbk_score_sites_host(const double* logodds, int m, const char* sites, int n_sites,
                          double* fwd, double* rc, double* both);

/* Sequence table of a packed buffer (host int64 arrays).  d_table: caller-owned device
 * int64[cgbk_seqset_table_ints(n_seq)] that must outlive the set (the library allocates
 * nothing on the per-genome path); the upload is ordered on `stream`. */
int64_t cgbk_seqset_table_ints(int n_seq);
int cgbk_seqset_create(const int64_t* seq_off, const int64_t* seq_len, int n_seq, int device,
                       int64_t* d_table, void* stream, cgbk_seqset** out);
void cgbk_seqset_destroy(cgbk_seqset* set);
int64_t cgbk_seqset_total_windows(const cgbk_seqset* set, int m);

/* ---- exact kernels (float64 accumulate in the reference's order) ------------- */

/* PSSMModel.score_seq (cgb/pssm_model.py:103-133) for n_regions windows-ranges.
 * Region k covers windows gpos = d_start[k] + i, i in [0, d_out_off[k+1]-d_out_off[k]);
 * d_out_off is the exclusive prefix sum of window counts (n_regions+1 entries).
 * Outputs (each may be NULL): float32 forward / minus-strand scores after the
 * ambiguity repair (what both=False returns) and the float64 soft-max
 * log2(2**f + 2**r) (both=True).  flags: CGBK_RC_*, CGBK_SINGLE_WINDOW_F64 (then
 * d_f64 / d_r64 receive the un-rounded pair of each 1-window region). */
int cgbk_score_windows(const cgbk_model* model, const cgbk_genome* genome,
                       const int64_t* d_start, const int64_t* d_out_off, int n_regions,
                       int64_t total_windows, int flags,
                       float* d_fwd32, float* d_rc32, double* d_both64,
                       double* d_f64, double* d_r64, void* stream);

/* TFBindingModel.binding_probability (cgb/binding_model.py:94-113) for a batch of
 * promoter regions (Gene.promoter_region, cgb/gene.py:111-127): region k scores
 * windows [d_start[k], d_start[k]+d_nwin[k]).  d_mu_std_m / d_mu_std_bg are
 * DEVICE double[2] {mean, std} (so the background pair can come straight from
 * cgbk_background_stats without a host round trip).  d_post[k] = posterior.
 * d_plane (optional, with `set`): the score plane cgbk_background_stats wrote for this
 * model and set; windows are then read from it and only the rare high-scoring ones
 * (mixture ratio x >= 1e-4) are re-scored exactly.  NULL: every window is scored in
 * float64 in the reference's order. */
int cgbk_posteriors(const cgbk_model* model, const cgbk_genome* genome,
                    const cgbk_seqset* set, const int32_t* d_plane,
                    const int64_t* d_start, const int32_t* d_nwin, int n_regions,
                    const double* d_mu_std_m, const double* d_mu_std_bg,
                    double p_motif, double alpha, double* d_post, void* stream);

/* ---- tiled scan kernels (sm_100a hot path) ---------------------------------- */

/* Bytes of device workspace cgbk_scan_hits / cgbk_background_stats need. */
int64_t cgbk_scan_workspace_bytes(const cgbk_seqset* set, int m, int64_t cand_cap);

/* Genome.identify_sites hit loop (cgb/genome.py:433-445) over every window of every
 * sequence of the set: writes all windows with float32 score >= threshold (compared
 * in float64) on either strand, unordered, to d_hits[0..hit_cap); d_counters is a
 * device int64[CGBK_CTR_COUNT] (zeroed inside).  flags: CGBK_RC_*.  Filter kernel on
 * 16-bit k-mer tables, then exact float64 re-scoring of every candidate, so the
 * result is bit-exact with the reference. */
int cgbk_scan_hits(cgbk_model* model, const cgbk_genome* genome, const cgbk_seqset* set,
                   double threshold, int flags,
                   cgbk_hit* d_hits, int64_t hit_cap, int64_t* d_counters,
                   void* d_work, int64_t work_bytes, int64_t cand_cap, void* stream);

/* Background distribution of Genome.build_PSSM_model (cgb/genome.py:215-226) taken
 * over ALL windows of the set instead of a 10 000-window sample: soft-max score of
 * every window -> histogram (uint64 d_hist[n_bins] over [hist_lo, hist_hi), clamped)
 * and moments (device double[CGBK_BG_COUNT], layout above).  With accumulate != 0
 * the raw n/sum/sumsq/hist are added to what the buffers hold (mean/std are
 * recomputed), for multi-call or multi-rank accumulation.  d_plane (optional,
 * plane_ints >= cgbk_score_plane_ints): receives every window's fixed-point soft-max
 * score in the scan's own tile order (one coalesced 128-byte store per position), for
 * cgbk_posteriors. */
int64_t cgbk_score_plane_ints(const cgbk_seqset* set, int m);
int cgbk_background_stats(cgbk_model* model, const cgbk_genome* genome, const cgbk_seqset* set,
                          double hist_lo, double hist_hi, int n_bins,
                          unsigned long long* d_hist, double* d_stats, int accumulate,
                          int32_t* d_plane, int64_t plane_ints,
                          void* d_work, int64_t work_bytes, void* stream);

/* The reference's own background population (cgb/genome.py:215-218): every m-mer of every
 * gene's upstream non-coding region, each scored as a sequence of its own (score_seq with
 * len(seq) == m), regions overlapping or repeated as the gene table has them.  Region k is the
 * windows [d_start[k], d_start[k] + d_nwin[k]) of the packed buffer -- the caller drops the
 * last window of each region as xrange(len(seq) - m) does.  Exact float64 arithmetic in the
 * reference's order; moments (and the optional histogram) as cgbk_background_stats.  The
 * reference then scores a random.sample of <= 10 000 of these m-mers; here all of them count. */
int64_t cgbk_region_stats_workspace_bytes(void);
int cgbk_background_stats_regions(const cgbk_model* model, const cgbk_genome* genome,
                                  const int64_t* d_start, const int32_t* d_nwin, int n_regions,
                                  double hist_lo, double hist_hi, int n_bins,
                                  unsigned long long* d_hist, double* d_stats, int accumulate,
                                  void* d_work, int64_t work_bytes, void* stream);

/* ---- motif batches ------------------------------------------------------------- */

/* n_models PSWMs over one packed buffer in ONE call and one pass per motif: for model k the
 * background record of all windows (d_stats[k][CGBK_BG_COUNT], and d_hist[k][n_bins] unless
 * d_hist is NULL: then the scan runs without a histogram) and -- when `thresholds` is given --
 * every site with float32 score >= thresholds[k] on either strand (Genome.identify_sites' hit
 * test, cgb/genome.py:433-445), found by the SAME pass: the stats scan's int32 fixed-point
 * strand scores are compared with the threshold minus their error bound (one flag bit per
 * window) and the flagged windows are re-scored in float64 in the reference's order (bit-exact
 * hit lists).
 *   win_cap      >= 0: only the first win_cap windows of each sequence are scored (sets of up to
 *                4 sequences).  A rank's shard of a longer sequence carries the halo of the longest
 *                motif and caps every motif at the shard's own windows; < 0: all windows
 *   hist_lo/hi   per model, or NULL for [floor(min score), ceil(max score) + 1)
 *   d_hits8      one compact list for the whole batch; model k's sites are the records
 *                [sum of CGBK_CTR_HITS of the models before k, + its own count), ordered by
 *                (position, + strand first) -- the order comes from prefix sums over the scan
 *                tiles, not from a sort, and is the same on every run
 *   d_counters   int64[n_models + 1][CGBK_CTR_COUNT]: per model candidates / hits / tiles with
 *                non-ACGT bytes; block n_models: [0] = records appended in total.  A total above
 *                hit_cap, or a model's candidate count above cand_cap, means records were
 *                dropped: run the models from the first affected one again with larger buffers
 *   cand_cap     windows of ONE model that may be flagged for the exact re-score (its sites plus
 *                the few windows within the fixed-point error of the threshold); 0 with
 *                thresholds == NULL
 *   d_work       >= cgbk_scan_batch_workspace_bytes(set, win_cap, cand_cap): moment partials,
 *                the scan's candidate flag plane (one bit per window), per-tile counts and
 *                16 bytes per candidate
 * flags: CGBK_RC_*.  No allocation, no synchronisation. */
int64_t cgbk_scan_batch_workspace_bytes(const cgbk_seqset* set, int64_t win_cap, int64_t cand_cap);
int cgbk_scan_batch(cgbk_model* const* models, int n_models, const cgbk_genome* genome,
                    const cgbk_seqset* set, int64_t win_cap, const double* thresholds, int flags,
                    const double* hist_lo, const double* hist_hi, int n_bins,
                    unsigned long long* d_hist, double* d_stats,
                    cgbk_hit8* d_hits8, int64_t hit_cap, int64_t* d_counters,
                    void* d_work, int64_t work_bytes, int64_t cand_cap, void* stream);

/* Sort n compact records in place by (position, + strand first); d_tmp >=
 * cgbk_sort_hits_workspace_bytes(n) bytes of device scratch.  (cgbk_scan_batch already delivers
 * ordered lists; this is for lists merged from several shards or scans.) */
int64_t cgbk_sort_hits_workspace_bytes(int64_t n);
int cgbk_sort_hits(cgbk_hit8* d_hits8, int64_t n, void* d_tmp, int64_t tmp_bytes, void* stream);

/* Split form of an ORDERED site list of one model for the trip to the host: 6 bytes per site
 * instead of 8 (every site of a large scan crosses PCIe; the reference keeps its sites as Python
 * objects, genome.py:433-445).
 *   local[k]       uint16: (offset of site k inside its 4 096-base block of the packed buffer) << 1 | minus
 *   score[k]       float32: the score, bit for bit
 *   block_start[b] b = 0 .. cgbk_split_sites_blocks(cap_bases): first site at or after base b * 4096
 * position of site k = (block << 12) + (local[k] >> 1): the 8-byte records can be rebuilt exactly. */
int64_t cgbk_split_sites_blocks(int64_t cap_bases);
int cgbk_split_sites(const cgbk_hit8* d_hits8, int64_t n, int64_t cap_bases, uint16_t* d_local, float* d_score,
                     int32_t* d_block_start, void* stream);
/* The lists of n_models models, back to back, in one launch: d_cum (DEVICE, int64 [n_models + 1]) = where
 * each list starts (d_cum[n_models] = n_total, each list below 2^31 sites); d_block_start is
 * [n_models][blocks + 1], relative to the start of each list. */
int cgbk_split_sites_batch(const cgbk_hit8* d_hits8, const int64_t* d_cum, int n_models, int64_t n_total,
                           int64_t cap_bases, uint16_t* d_local, float* d_score, int32_t* d_block_start,
                           void* stream);

/* Recompute CGBK_BG_MEAN / CGBK_BG_STD from n/sum/sumsq in d_stats (after an
 * all-reduce of the raw sums across ranks). */
int cgbk_background_finalize(double* d_stats, void* stream);
/* The same for n_records records laid out back to back (a motif batch after ONE all-reduce). */
int cgbk_background_finalize_batch(double* d_stats, int n_records, void* stream);

/* ---- host-buffer entry point -------------------------------------------------- */

/* One replicon, HOST buffers in and out: the hot-path part of the reference's per-genome
 * loop (cgb/__init__.py:686-695) -- Genome.build_PSSM_model's background fit
 * (cgb/genome.py:215-226, over all windows) then Genome.calculate_regulation_probabilities
 * (cgb/genome.py:330-333, cgb/gene.py:129-146, cgb/binding_model.py:94-113).
 *   h_ascii[n_bases]      the sequence bytes (pinned memory makes the copy asynchronous)
 *   h_start/h_end[n]      promoter slices sequence[start:end] (Python slice semantics,
 *                         cgb/gene.py:127, cgb/chromid.py:94); a slice shorter than the motif
 *                         is CGBK_ERR_INVALID (the reference raises there too)
 *   mu_m, std_m           motif score distribution (binding_model.py:85)
 *   d_scratch             device, >= cgbk_pipeline_scratch_bytes(n_bases, n, m, n_bins)
 *   h_staging             host, >= cgbk_pipeline_staging_bytes(n) (pinned: asynchronous upload)
 *   h_post[n]             posteriors; h_stats[CGBK_BG_COUNT]; h_hist[n_bins] (or NULL).  When
 *                         these three are pinned the kernels store the results straight into
 *                         them; otherwise they are copied back (one copy if the buffers are
 *                         laid out post | stats | hist)
 * The sequence goes up in one copy on `stream`, or -- from 16 MB on -- in four chunks on an
 * internal side stream while `stream` packs and scans the chunks that have arrived.  Returns
 * after one synchronisation of `stream`; the model's estimator is NOT touched (the caller passes
 * mu_m / std_m and reads mean / std of the background from h_stats). */
int64_t cgbk_pipeline_scratch_bytes(int64_t n_bases, int n_regions, int m, int n_bins);
int64_t cgbk_pipeline_staging_bytes(int n_regions);
int cgbk_regulation_pipeline(cgbk_model* model, const uint8_t* h_ascii, int64_t n_bases,
                             const int64_t* h_start, const int64_t* h_end, int n_regions,
                             double mu_m, double std_m, double p_motif, double alpha,
                             double hist_lo, double hist_hi, int n_bins,
                             void* d_scratch, int64_t scratch_bytes,
                             void* h_staging, int64_t staging_bytes,
                             double* h_post, double* h_stats, unsigned long long* h_hist,
                             void* stream);

/* The same with the sequence already PACKED on the host (cgbk_pack_host, or a packed genome
 * file): h_bases2[cap / 16], h_amb[cap / 32], h_lower[cap / 32] with cap =
 * cgbk_packed_capacity(n_bases); h_amb and / or h_lower may be NULL, meaning "no such byte in
 * this replicon" (the plane is cleared on the device instead of uploaded).  0.25 - 0.5 byte per
 * base crosses PCIe instead of 1 and no pack kernel runs; from 64 Mbp on the planes go up in
 * four chunks beside the scan like the ASCII form does.  Scratch as for the ASCII call. */
int cgbk_regulation_pipeline_packed(cgbk_model* model, const uint32_t* h_bases2,
                                    const uint32_t* h_amb, const uint32_t* h_lower, int64_t n_bases,
                                    const int64_t* h_start, const int64_t* h_end, int n_regions,
                                    double mu_m, double std_m, double p_motif, double alpha,
                                    double hist_lo, double hist_hi, int n_bins,
                                    void* d_scratch, int64_t scratch_bytes,
                                    void* h_staging, int64_t staging_bytes,
                                    double* h_post, double* h_stats, unsigned long long* h_hist,
                                    void* stream);

/* Either call with its arguments in one struct: a binding keeps the struct alive and only
 * rewrites the fields that change between calls (a ctypes call with 21 scalar arguments costs
 * several microseconds of marshalling; with one pointer it costs one).  h_ascii != NULL selects
 * the ASCII form, otherwise h_bases2 (+ optional h_amb / h_lower) the packed one. */
typedef struct {
    cgbk_model* model;
    const uint8_t* h_ascii;
    int64_t n_bases;
    const int64_t* h_start;
    const int64_t* h_end;
    int32_t n_regions;
    int32_t n_bins;
    double mu_m, std_m, p_motif, alpha, hist_lo, hist_hi;
    void* d_scratch;
    int64_t scratch_bytes;
    void* h_staging;
    int64_t staging_bytes;
    double* h_post;
    double* h_stats;
    unsigned long long* h_hist;
    void* stream;
    const uint32_t* h_bases2;      /* packed form (used when h_ascii is NULL) */
    const uint32_t* h_amb;
    const uint32_t* h_lower;
} cgbk_pipeline_args;
int cgbk_regulation_pipeline_args(const cgbk_pipeline_args* args);

/* Launch statistics of the calling thread's last tiled-scan call (for bench.py):
 * out[0] = kernels launched, out[1] = tiles, out[2] = bases per lane segment. */
int cgbk_last_launch_info(int64_t* out3);

/* Optional device timing of the dominant (tiled scan) kernel for bench.py: when
 * enabled, cgbk_scan_hits / cgbk_background_stats record CUDA events on `stream`
 * right before and after that kernel; cgbk_profile_last_ms synchronises on the stop
 * event and returns the elapsed milliseconds of the calling thread's last scan. */
int cgbk_profile_enable(int on);
int cgbk_profile_last_ms(float* ms);
/* on == 2: cgbk_scan_batch also brackets every model's scan kernel with its own event pair and,
 * when sites are searched, every model's site pass with another; cgbk_profile_batch_ms returns the
 * n first times of the calling thread's last cgbk_scan_batch (synchronises on the events): the
 * scan kernels of models 0 .. n_models - 1, then their site passes. */
int cgbk_profile_batch_ms(float* ms, int n);

#ifdef __cplusplus
}
#endif
#endif /* CGBK_H */
