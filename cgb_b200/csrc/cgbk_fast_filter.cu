// MODE 0 of the tiled scan (cgbk_fast.cuh): the 16-bit hit filter of cgbk_scan_hits.
#include "cgbk_fast.cuh"

static cudaError_t dispatch_filter(int G, const uint32_t* gtab, GenomeView g, SeqTable seqs, FastWork w, int m,
                                   int L, int grid, StatsParams sp, int device, cudaStream_t st) {
#define CGBK_CASE(GG)                                                                                              \
    case GG:                                                                                                       \
        return L == 32 ? launch_one<GG, 0, CGBK_VAR_ONE>(gtab, g, seqs, w, m, L, grid, sp, device, st)             \
                       : launch_one<GG, 0, CGBK_VAR_UNIFIED>(gtab, g, seqs, w, m, L, grid, sp, device, st);
    switch (G) {
        CGBK_CASE(1) CGBK_CASE(2) CGBK_CASE(3) CGBK_CASE(4)
        CGBK_CASE(5) CGBK_CASE(6) CGBK_CASE(7) CGBK_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef CGBK_CASE
}

int fast_threads(int G, int mode) { return fast_threads_of(G, mode); }

cudaError_t launch_fast_filter(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, int L,
                               int grid, cudaStream_t st) {
    StatsParams sp = {};
    return dispatch_filter(mdl->G, mdl->d_coarse, g, seqs, w, mdl->m, L, grid, sp, mdl->device, st);
}
