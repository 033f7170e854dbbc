// MODE 4 of the tiled scan (cgbk_fast.cuh): MODE 2 (moments, + hit candidates) without the score plane.
#include "cgbk_fast.cuh"

int stats_variant(int L);

static cudaError_t dispatch_lean_moments(int G, const uint32_t* gtab, GenomeView g, SeqTable seqs, FastWork w, int m,
                                         int L, int grid, StatsParams sp, int device, cudaStream_t st) {
    const bool peel = stats_variant(L) == CGBK_VAR_PEEL;
#define CGBK_CASE(GG)                                                                                              \
    case GG:                                                                                                       \
        return peel ? launch_one<GG, 4, CGBK_VAR_PEEL>(gtab, g, seqs, w, m, L, grid, sp, device, st)               \
                    : launch_one<GG, 4, CGBK_VAR_UNIFIED>(gtab, g, seqs, w, m, L, grid, sp, device, st);
    switch (G) {
        CGBK_CASE(1) CGBK_CASE(2) CGBK_CASE(3) CGBK_CASE(4)
        CGBK_CASE(5) CGBK_CASE(6) CGBK_CASE(7) CGBK_CASE(8)
        default: return cudaErrorInvalidValue;
    }
#undef CGBK_CASE
}

cudaError_t launch_fast_moments_lean(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, int L, int grid,
                                     const StatsParams& sp, cudaStream_t st) {
    return dispatch_lean_moments(mdl->G, mdl->d_precise, g, seqs, w, mdl->m, L, grid, sp, mdl->device, st);
}
