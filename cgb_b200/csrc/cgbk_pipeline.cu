// Host-buffer entry point: one replicon's ASCII bytes + promoter slices in, posteriors out.
//
// This is the per-genome inner loop of the reference's go() for the hot path
// (cgb/__init__.py:686-695): Genome.build_PSSM_model's background fit (genome.py:215-226,
// over ALL windows) followed by Genome.calculate_regulation_probabilities
// (genome.py:330-333 -> gene.py:129-146 -> binding_model.py:94-113), behind ONE C-ABI call
// that takes HOST buffers.  Everything is enqueued on the caller's stream (H2D of the
// sequence, pack, stats scan, dirty tiles + moment reduction, H2D of the promoter
// descriptors, posteriors, D2H of the results) and the call returns after one stream
// synchronisation, so the host-side cost is one function call instead of a dozen Python /
// driver round trips.
#include <math.h>
#include <string.h>

#include <vector>

#include "cgbk_internal.cuh"

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

struct PipeLayout {
    int64_t cap_bases;
    int64_t off_ascii, off_planes, off_table, off_work, off_plane, off_hist, off_stats, off_msm, off_desc,
        off_post, total;
    int64_t work_bytes, plane_ints;
};

static PipeLayout pipe_layout(int64_t n_bases, int n_regions, int m, int n_bins) {
    PipeLayout p;
    int64_t padded = (n_bases + CGBK_SEQ_ALIGN - 1) / CGBK_SEQ_ALIGN * CGBK_SEQ_ALIGN;
    if (padded == 0) padded = CGBK_SEQ_ALIGN;
    p.cap_bases = padded;
    const int64_t nwin = n_bases - m + 1 > 0 ? n_bases - m + 1 : 0;
    const int64_t tiles32 = (nwin + 991) / 992;                 // most tiles: lane segment 32
    p.work_bytes = align256(64 + 3 * 8 * 1024 + 4 * tiles32 + 64 + 64);
    p.plane_ints = tiles32 * 2048;                              // (B + 1) * 1024 at B = 1 is the largest
    int64_t o = 0;
    p.off_ascii = o;  o += align256(n_bases);
    p.off_planes = o; o += align256(padded / 2);                // bases2 (cap/4) + amb + lower (cap/8 each)
    p.off_table = o;  o += align256(8 * 4);
    p.off_work = o;   o += p.work_bytes;
    p.off_plane = o;  o += align256(4 * p.plane_ints);
    p.off_hist = o;   o += align256(8 * (int64_t)(n_bins > 0 ? n_bins : 1));
    p.off_stats = o;  o += 256;
    p.off_msm = o;    o += 256;
    p.off_desc = o;   o += align256(12 * (int64_t)n_regions);
    p.off_post = o;   o += align256(8 * (int64_t)n_regions);
    p.total = o;
    return p;
}

extern "C" int64_t cgbk_pipeline_scratch_bytes(int64_t n_bases, int n_regions, int m, int n_bins) {
    if (n_bases < 0 || n_regions < 0 || m < 1) return 0;
    return pipe_layout(n_bases, n_regions, m, n_bins).total;
}

extern "C" int64_t cgbk_pipeline_staging_bytes(int n_regions) { return align256(12 * (int64_t)n_regions) + 256; }

extern "C" int cgbk_regulation_pipeline(cgbk_model* M, const uint8_t* h_ascii, int64_t n_bases,
                                        const int64_t* h_start, const int64_t* h_end, int n_regions,
                                        double mu_m, double std_m, double p_motif, double alpha,
                                        double hist_lo, double hist_hi, int n_bins,
                                        void* d_scratch, int64_t scratch_bytes,
                                        void* h_staging, int64_t staging_bytes,
                                        double* h_post, double* h_stats, unsigned long long* h_hist,
                                        void* stream) {
    if (!M || !h_ascii || n_bases < 1 || n_regions < 0 || !d_scratch || !h_stats ||
        (n_regions > 0 && (!h_start || !h_end || !h_post)))
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    if (n_bins < 0 || n_bins > CGBK_MAX_HIST_BINS) return cgbk_set_error(CGBK_ERR_INVALID, "bad n_bins");
    const int m = M->m;
    const PipeLayout P = pipe_layout(n_bases, n_regions, m, n_bins);
    if (scratch_bytes < P.total)
        return cgbk_set_error(CGBK_ERR_INVALID, "scratch of %lld bytes is smaller than the %lld required",
                              (long long)scratch_bytes, (long long)P.total);
    if (!h_staging || staging_bytes < cgbk_pipeline_staging_bytes(n_regions))
        return cgbk_set_error(CGBK_ERR_INVALID, "staging buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)d_scratch;

    // promoter slices -> (packed offset, window count): Chromid.subsequence's Python slice
    // (cgb/chromid.py:94) then len(seq) - m + 1 windows; len(seq) < m raises in the reference
    int64_t* st_start = (int64_t*)h_staging;
    int32_t* st_nwin = (int32_t*)((unsigned char*)h_staging + 8 * (int64_t)n_regions);
    for (int i = 0; i < n_regions; ++i) {
        int64_t s = h_start[i], e = h_end[i];
        s = s < 0 ? (s + n_bases > 0 ? s + n_bases : 0) : (s < n_bases ? s : n_bases);
        e = e < 0 ? (e + n_bases > 0 ? e + n_bases : 0) : (e < n_bases ? e : n_bases);
        if (e < s) e = s;
        const int64_t nw = e - s - m + 1;
        if (nw < 1)
            return cgbk_set_error(CGBK_ERR_INVALID, "region %d is shorter than the motif "
                                  "(negative dimensions are not allowed)", i);
        st_start[i] = s;           // sequence 0 starts at packed offset 0
        st_nwin[i] = (int32_t)nw;
    }
    double* st_msm = (double*)((unsigned char*)h_staging + cgbk_pipeline_staging_bytes(n_regions) - 256);
    st_msm[0] = mu_m; st_msm[1] = std_m;

    // 1. sequence to the device and into the 2-bit planes
    uint8_t* d_ascii = base + P.off_ascii;
    if (cudaMemcpyAsync(d_ascii, h_ascii, n_bases, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return cgbk_set_error(CGBK_ERR_CUDA, "H2D of the sequence failed");
    cgbk_genome G;
    uint32_t* planes = (uint32_t*)(base + P.off_planes);
    G.d_bases2 = planes;
    G.d_amb = planes + P.cap_bases / 16;
    G.d_lower = planes + P.cap_bases / 16 + P.cap_bases / 32;
    G.cap_bases = P.cap_bases;
    int rc = cgbk_pack_ascii(d_ascii, n_bases, 0, planes, planes + P.cap_bases / 16,
                             planes + P.cap_bases / 16 + P.cap_bases / 32, P.cap_bases, st);
    if (rc) return rc;

    // 2. background: stats scan (+ score plane) over all windows
    cgbk_seqset* S = nullptr;
    const int64_t zero = 0;
    rc = cgbk_seqset_create(&zero, &n_bases, 1, M->device, (int64_t*)(base + P.off_table), st, &S);
    if (rc) return rc;
    unsigned long long* d_hist = n_bins > 0 ? (unsigned long long*)(base + P.off_hist) : nullptr;
    double* d_stats = (double*)(base + P.off_stats);
    int32_t* d_plane = n_regions > 0 ? (int32_t*)(base + P.off_plane) : nullptr;
    rc = cgbk_background_stats(M, &G, S, hist_lo, hist_hi, n_bins, d_hist, d_stats, 0, d_plane, P.plane_ints,
                               base + P.off_work, P.work_bytes, st);
    if (rc) { cgbk_seqset_destroy(S); return rc; }

    // 3. posteriors of every promoter from the plane, estimator straight from d_stats
    if (n_regions > 0) {
        cudaError_t e1 = cudaMemcpyAsync(base + P.off_desc, h_staging, 12 * (int64_t)n_regions,
                                         cudaMemcpyHostToDevice, st);
        cudaError_t e2 = cudaMemcpyAsync(base + P.off_msm, st_msm, 16, cudaMemcpyHostToDevice, st);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            cgbk_seqset_destroy(S);
            return cgbk_set_error(CGBK_ERR_CUDA, "H2D of the promoter table failed");
        }
        rc = cgbk_posteriors(M, &G, S, d_plane, (const int64_t*)(base + P.off_desc),
                             (const int32_t*)(base + P.off_desc + 8 * (int64_t)n_regions), n_regions,
                             (const double*)(base + P.off_msm), d_stats + CGBK_BG_MEAN, p_motif, alpha,
                             (double*)(base + P.off_post), st);
        if (rc) { cgbk_seqset_destroy(S); return rc; }
        cudaMemcpyAsync(h_post, base + P.off_post, 8 * (int64_t)n_regions, cudaMemcpyDeviceToHost, st);
    }
    cudaMemcpyAsync(h_stats, d_stats, sizeof(double) * CGBK_BG_COUNT, cudaMemcpyDeviceToHost, st);
    if (h_hist && n_bins > 0)
        cudaMemcpyAsync(h_hist, d_hist, sizeof(unsigned long long) * n_bins, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cgbk_seqset_destroy(S);
    if (e != cudaSuccess) return cgbk_check_cuda(e, "regulation pipeline");
    return CGBK_OK;
}
