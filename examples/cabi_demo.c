/* A non-Python host of libcgbk.so: plain C, the CUDA runtime for memory, nothing else.
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/cabi_demo.c \
 *       -Lcgb_b200/_lib -lcgbk -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/cgb_b200/_lib -o cabi_demo
 *
 * It builds a 4-position log-odds matrix by hand (the KAT of SURVEY.md §8c: sites = ["ACGT"],
 * pseudocount 1 -> 0.6780719 for a match, -0.3219281 otherwise), asks the library for IC and
 * Patser threshold (host float64, no GPU), and -- when a device is present -- packs a sequence,
 * lists the hits and runs the one-call regulation pipeline.  Exit code 0 = all checks passed;
 * `--host-only` skips the device part (what the CPU test runs).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>

#include "cgbk.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != CGBK_OK) {                                                            \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cgbk_last_error());      \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

static int near(double a, double b, double tol) { return fabs(a - b) <= tol; }

int main(int argc, char** argv) {
    const int host_only = argc > 1 && strcmp(argv[1], "--host-only") == 0;
    const int m = 4;
    double lo[16];
    for (int j = 0; j < m; ++j)
        for (int b = 0; b < 4; ++b) lo[j * 4 + b] = (b == j) ? log(0.4 / 0.25) / log(2.0) : log(0.2 / 0.25) / log(2.0);

    double thr = 0, ic = 0, mn = 0, mx = 0;
    CHECK(cgbk_patser_threshold(lo, m, 1000, &thr, &ic, &mn, &mx));
    printf("version %d  IC %.6f  threshold %.6f  min %.6f  max %.6f\n", cgbk_version(), ic, thr, mn, mx);
    if (!near(mx, 4 * 0.6780719051126377, 1e-12) || !near(mn, 4 * -0.3219280948873623, 1e-12)) return 2;
    if (cgbk_patser_threshold(lo, 0, 1000, &thr, NULL, NULL, NULL) != CGBK_ERR_INVALID) return 3;
    if (host_only) { puts("host-only checks passed"); return 0; }

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) { fprintf(stderr, "no CUDA device\n"); return 4; }

    /* a 4 096-base sequence: ACGT planted at known places in a run of T's, plus an N */
    enum { N = 4096 };
    unsigned char* seq = (unsigned char*)malloc(N);
    memset(seq, 'T', N);
    const int planted[3] = {100, 2000, 4092};
    for (int k = 0; k < 3; ++k) memcpy(seq + planted[k], "ACGT", 4);
    seq[3000] = 'N';

    const int64_t cap = (N + CGBK_SEQ_ALIGN - 1) / CGBK_SEQ_ALIGN * CGBK_SEQ_ALIGN;
    unsigned char* d_ascii; uint32_t *d_b2, *d_amb, *d_low;
    cudaMalloc((void**)&d_ascii, N); cudaMalloc((void**)&d_b2, cap / 4);
    cudaMalloc((void**)&d_amb, cap / 8); cudaMalloc((void**)&d_low, cap / 8);
    cudaMemcpy(d_ascii, seq, N, cudaMemcpyHostToDevice);
    CHECK(cgbk_pack_ascii(d_ascii, N, 0, d_b2, d_amb, d_low, cap, NULL));

    cgbk_model* M = NULL;
    CHECK(cgbk_model_create(lo, m, 0, &M));
    const int64_t off = 0, len = N;
    int64_t* d_table; cudaMalloc((void**)&d_table, 8 * cgbk_seqset_table_ints(1));
    cgbk_seqset* S = NULL;
    CHECK(cgbk_seqset_create(&off, &len, 1, 0, d_table, NULL, &S));

    cgbk_genome G; G.d_bases2 = d_b2; G.d_amb = d_amb; G.d_lower = d_low; G.cap_bases = cap;
    const int64_t hit_cap = 64, cand_cap = 4096;
    const int64_t wbytes = cgbk_scan_workspace_bytes(S, m, cand_cap);
    void* d_work; cudaMalloc(&d_work, wbytes);
    cgbk_hit* d_hits; cudaMalloc((void**)&d_hits, sizeof(cgbk_hit) * hit_cap);
    int64_t* d_ctr; cudaMalloc((void**)&d_ctr, 64);
    /* the forward score of ACGT is 2.712; its reverse complement is ACGT again */
    CHECK(cgbk_scan_hits(M, &G, S, 2.5, CGBK_RC_REVSEQ_ORDER, d_hits, hit_cap, d_ctr, d_work, wbytes, cand_cap, NULL));
    int64_t ctr[8]; cgbk_hit hits[64];
    cudaMemcpy(ctr, d_ctr, 64, cudaMemcpyDeviceToHost);
    cudaMemcpy(hits, d_hits, sizeof(hits), cudaMemcpyDeviceToHost);
    printf("hits %lld (dirty tiles %lld)\n", (long long)ctr[CGBK_CTR_HITS], (long long)ctr[CGBK_CTR_DIRTY_TILES]);
    if (ctr[CGBK_CTR_HITS] != 6) return 5;                        /* 3 sites x 2 strands */
    for (int i = 0; i < 6; ++i) {
        int ok = 0;
        for (int k = 0; k < 3; ++k) ok |= hits[i].gpos == planted[k];
        if (!ok || !near(hits[i].score, 2.7122876, 1e-6)) return 6;
    }

    /* the one-call host-buffer pipeline: background of all windows + posterior of 3 slices */
    const int64_t starts[3] = {0, 1900, -300}, ends[3] = {350, 2250, N};
    const int n_bins = 64;
    const int64_t sbytes = cgbk_pipeline_scratch_bytes(N, 3, m, n_bins);
    void* d_scratch; cudaMalloc(&d_scratch, sbytes);
    const int64_t stage_bytes = cgbk_pipeline_staging_bytes(3);
    void* h_stage = malloc(stage_bytes);
    double post[3], stats[CGBK_BG_COUNT]; unsigned long long hist[64];
    CHECK(cgbk_regulation_pipeline(M, seq, N, starts, ends, 3, 2.7122876, 0.5, 0.03, 1.0 / 350.0, -3.0, 5.0, n_bins,
                                   d_scratch, sbytes, h_stage, stage_bytes, post, stats, hist, NULL));
    unsigned long long total = 0;
    for (int i = 0; i < n_bins; ++i) total += hist[i];
    printf("background n %.0f mean %.4f std %.4f; posteriors %.4f %.4f %.4f\n", stats[CGBK_BG_N],
           stats[CGBK_BG_MEAN], stats[CGBK_BG_STD], post[0], post[1], post[2]);
    if (stats[CGBK_BG_N] != N - m + 1 || total != (unsigned long long)(N - m + 1)) return 7;
    for (int i = 0; i < 3; ++i)
        if (!(post[i] > 0.03 && post[i] <= 1.0)) return 8;         /* each slice holds a planted site */

    cgbk_seqset_destroy(S); cgbk_model_destroy(M);
    cudaFree(d_scratch); cudaFree(d_ctr); cudaFree(d_hits); cudaFree(d_work); cudaFree(d_table);
    cudaFree(d_low); cudaFree(d_amb); cudaFree(d_b2); cudaFree(d_ascii);
    free(h_stage); free(seq);
    puts("all checks passed");
    return 0;
}
