// Host-buffer entry point: one replicon's ASCII bytes + promoter slices in, posteriors out.
//
// This is the per-genome inner loop of the reference's go() for the hot path
// (cgb/__init__.py:686-695): Genome.build_PSSM_model's background fit (genome.py:215-226,
// over ALL windows) followed by Genome.calculate_regulation_probabilities
// (genome.py:330-333 -> gene.py:129-146 -> binding_model.py:94-113), behind ONE C-ABI call
// that takes HOST buffers.  The call enqueues the upload of the sequence (one copy, or chunks
// on a side stream for large replicons), the pack, the stats scan, the upload of the promoter
// descriptors (side stream), the posteriors and ONE download of the results, and returns after
// one stream synchronisation: the host-side cost is one function call instead of a dozen
// Python / driver round trips.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "cgbk_internal.cuh"

static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

// -DCGBK_PIPE_TRACE (tools/pipeline_trace.py builds that variant): device and host timestamps of
// the call's phases on stderr.  Compiled out of the shipped library.
#ifdef CGBK_PIPE_TRACE
#include <stdio.h>
#include <time.h>
struct PipeTrace {
    const char* name[24]; cudaEvent_t ev[24]; double host_us[24]; int n = 0;
    static double now_us() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
    void mark(const char* what, cudaStream_t st) {
        if (n >= 24) return;
        cudaEventCreate(&ev[n]); cudaEventRecord(ev[n], st); name[n] = what; host_us[n] = now_us(); ++n;
    }
    void dump() {
        const double end = now_us();
        for (int i = 0; i < n; ++i) {
            float ms = 0.f; cudaEventElapsedTime(&ms, ev[0], ev[i]);
            fprintf(stderr, "[pipe] %-12s device +%8.1f us   host enqueue +%8.1f us\n", name[i], ms * 1e3,
                    host_us[i] - host_us[0]);
            }
        fprintf(stderr, "[pipe] %-12s host +%8.1f us\n", "returned", end - host_us[0]);
        for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
    }
};
#define TRACE_DECL PipeTrace trace
#define TRACE(what) trace.mark(what, st)
#define TRACE_ON(what, stream) trace.mark(what, stream)
#define TRACE_DUMP() trace.dump()
#else
#define TRACE_DECL
#define TRACE(what)
#define TRACE_ON(what, stream)
#define TRACE_DUMP()
#endif

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
    p.work_bytes = align256(64 + 4 * 8 * 1024 + 4 * tiles32 + 64 + 64);   // = cgbk_scan_workspace_bytes(.., cand_cap 0)
    p.plane_ints = tiles32 * 2048;                              // (B + 1) * 1024 at B = 1 is the largest
    int64_t o = 0;
    p.off_ascii = o;  o += align256(n_bases);
    p.off_planes = o; o += align256(padded / 2);                // bases2 (cap/4) + amb + lower (cap/8 each)
    p.off_table = o;  o += align256(8 * cgbk_seqset_table_ints(1));
    p.off_plane = o;  o += align256(4 * p.plane_ints);
    p.off_msm = o;    o += 256;
    p.off_desc = o;   o += align256(12 * (int64_t)n_regions);
    // results, contiguous so that one D2H can return them: posteriors | stats | histogram; the
    // scan workspace (counters first) follows directly so that one memset clears record,
    // histogram and counters
    p.off_post = o;
    p.off_stats = p.off_post + 8 * (int64_t)n_regions;
    p.off_hist = p.off_stats + 8 * CGBK_BG_COUNT;
    p.off_work = p.off_hist + 8 * (int64_t)(n_bins > 0 ? n_bins : 1);
    o = p.off_work + p.work_bytes;
    o = align256(o);
    p.total = o;
    return p;
}

// side stream + events of the chunked upload: one set per (host thread, device), created on first
// use and released when the thread ends.  Per thread, not per process: two host threads that drive
// the same device through their own streams must not share the "scratch free" / "chunk arrived"
// events (a record by one would replace what the other is about to wait on).
#define PIPE_MAX_CHUNKS CGBK_MAX_TILE_RANGES
struct PipeStreams {
    cudaStream_t copy = nullptr;
    cudaEvent_t ev[PIPE_MAX_CHUNKS + 2] = {};   // chunk c arrived | scratch free | promoter table arrived
};
struct PipeStreamSet {
    PipeStreams* dev[64] = {};
    ~PipeStreamSet() {
        for (PipeStreams* p : dev) {
            if (!p) continue;
            // best effort: at process exit the context may already be gone
            for (cudaEvent_t e : p->ev) if (e) cudaEventDestroy(e);
            if (p->copy) cudaStreamDestroy(p->copy);
            delete p;
        }
        cudaGetLastError();
    }
};
static thread_local PipeStreamSet g_pipe;

static PipeStreams* pipe_streams(int device) {
    if (device < 0 || device >= 64) return nullptr;     // the caller has made `device` current
    if (g_pipe.dev[device]) return g_pipe.dev[device];
    PipeStreams* p = new PipeStreams;
    bool ok = cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < PIPE_MAX_CHUNKS + 2; ++i)
        ok = cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { delete p; return nullptr; }
    g_pipe.dev[device] = p;
    return p;
}

// device-visible alias of a pinned host pointer (false for pageable memory)
static bool device_alias(const void* h, void** d) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *d = at.devicePointer;
    return true;
}

// Chunk boundaries (multiples of CGBK_SEQ_ALIGN).  Measured on B200 / PCIe 5 (tools/h2d_probe.py,
// tools/pipeline_trace.py): cutting a pinned upload into 3-4 copies costs ~10 us, kernels
// running beside the DMA slow it by another ~5 %, and every chunk's scan pays the ~14 us fixed
// cost of a launch.  At 5 MB that cancels the overlap exactly (202 us either way); at 50 MB the
// chunked form wins ~5 % (1.17 vs 1.23 ms).  So small replicons go up in one piece and only
// large ones are pipelined, the last (exposed) chunk being the smallest.
static int plan_chunks(int64_t n, int64_t* cut, int bases_per_byte) {
    static const double frac3[3] = {0.50, 0.30, 0.20};
    static const double frac4[4] = {0.35, 0.30, 0.20, 0.15};
    int k = n >= (16ll << 20) * bases_per_byte ? 4 : 1;
    if (const char* e = getenv("CGBK_PIPE_CHUNKS")) {          // tuning knob: 1, 3 or 4
        const int want = atoi(e);
        if (want == 1 || ((want == 3 || want == 4) && n >= 4 * CGBK_SEQ_ALIGN)) k = want;
    }
    const double* f = k == 4 ? frac4 : frac3;
    cut[0] = 0;
    double acc = 0.0;
    for (int c = 1; c < k; ++c) {
        acc += f[c - 1];
        cut[c] = (int64_t)(acc * (double)n) / CGBK_SEQ_ALIGN * CGBK_SEQ_ALIGN;
    }
    cut[k] = n;
    return k;
}

extern "C" int64_t cgbk_pipeline_scratch_bytes(int64_t n_bases, int n_regions, int m, int n_bins) {
    if (n_bases < 0 || n_regions < 0 || m < 1) return 0;
    return pipe_layout(n_bases, n_regions, m, n_bins).total;
}

extern "C" int64_t cgbk_pipeline_staging_bytes(int n_regions) { return align256(12 * (int64_t)n_regions) + 256; }

// Where the sequence comes from: ASCII bytes (packed on the device), or the three planes already
// in the kernels' layout (cgbk_pack_host / a packed genome file): a quarter to a half of the bytes
// cross PCIe and no pack kernel runs.  A NULL amb / lower plane stands for "all zero".
struct PipeSource {
    const uint8_t* h_ascii;
    const uint32_t *h_bases2, *h_amb, *h_lower;
};

static int pipeline_impl(cgbk_model* M, const PipeSource& src, int64_t n_bases,
                         const int64_t* h_start, const int64_t* h_end, int n_regions,
                         double mu_m, double std_m, double p_motif, double alpha,
                         double hist_lo, double hist_hi, int n_bins,
                         void* d_scratch, int64_t scratch_bytes,
                         void* h_staging, int64_t staging_bytes,
                         double* h_post, double* h_stats, unsigned long long* h_hist,
                         void* stream) {
    const uint8_t* h_ascii = src.h_ascii;
    const bool packed = h_ascii == nullptr;
    if (!M || (!h_ascii && !src.h_bases2) || n_bases < 1 || n_regions < 0 || !d_scratch || !h_stats ||
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
    ON_DEVICE(M->device);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)d_scratch;
    TRACE_DECL;
    TRACE("enter");
    PipeStreams* ps = pipe_streams(M->device);
    if (!ps) return cgbk_set_error(CGBK_ERR_CUDA, "could not create the copy stream");
    cudaStream_t cs = ps->copy;
    uint8_t* d_ascii = base + P.off_ascii;
    int64_t cut[PIPE_MAX_CHUNKS + 1];
    // the chunking threshold is about bytes on the wire: packed input is 2-4x denser
    const int n_chunks = plan_chunks(n_bases, cut, packed ? 4 : 1);
    cgbk_genome G;
    uint32_t* planes = (uint32_t*)(base + P.off_planes);
    G.d_bases2 = planes;
    G.d_amb = planes + P.cap_bases / 16;
    G.d_lower = planes + P.cap_bases / 16 + P.cap_bases / 32;
    G.cap_bases = P.cap_bases;
    // bases [a, b) of the packed source -> device planes (b rounded up to the padded capacity for
    // the last piece: the host planes cover it, zero filled)
    auto upload_planes = [&](int64_t a, int64_t b, cudaStream_t s) -> cudaError_t {
        if (b >= n_bases) b = P.cap_bases;
        cudaError_t e = cudaMemcpyAsync(planes + a / 16, src.h_bases2 + a / 16, (b - a) / 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && src.h_amb)
            e = cudaMemcpyAsync((uint32_t*)G.d_amb + a / 32, src.h_amb + a / 32, (b - a) / 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && src.h_lower)
            e = cudaMemcpyAsync((uint32_t*)G.d_lower + a / 32, src.h_lower + a / 32, (b - a) / 8, cudaMemcpyHostToDevice, s);
        return e;
    };
    // whatever the caller's stream still has queued may use the scratch: the side stream starts
    // behind it
    if (cudaEventRecord(ps->ev[PIPE_MAX_CHUNKS], st) != cudaSuccess ||
        cudaStreamWaitEvent(cs, ps->ev[PIPE_MAX_CHUNKS], 0) != cudaSuccess)
        return cgbk_set_error(CGBK_ERR_CUDA, "stream hand-over failed");
    // one-piece upload: start the DMA now, prepare the promoter table while it runs
    const bool early = n_chunks == 1;
    if (early && !packed && cudaMemcpyAsync(d_ascii, h_ascii, n_bases, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return cgbk_set_error(CGBK_ERR_CUDA, "H2D of the sequence failed");
    if (early && packed && upload_planes(0, n_bases, st) != cudaSuccess)
        return cgbk_set_error(CGBK_ERR_CUDA, "H2D of the packed sequence failed");
    if (packed) {
        // absent planes are all zero (a replicon of upper-case ACGT only)
        cudaError_t e = cudaSuccess;
        if (!src.h_amb) e = cudaMemsetAsync((void*)G.d_amb, 0, P.cap_bases / 8, st);
        if (e == cudaSuccess && !src.h_lower) e = cudaMemsetAsync((void*)G.d_lower, 0, P.cap_bases / 8, st);
        if (e != cudaSuccess) { cudaStreamSynchronize(st); return cgbk_check_cuda(e, "clearing the absent planes"); }
    }

    // promoter slices -> (packed offset, window count): Chromid.subsequence's Python slice
    // (cgb/chromid.py:94) then len(seq) - m + 1 windows; len(seq) < m - 1 raises in the reference
    int64_t* st_start = (int64_t*)h_staging;
    int32_t* st_nwin = (int32_t*)((unsigned char*)h_staging + 8 * (int64_t)n_regions);
    for (int i = 0; i < n_regions; ++i) {
        int64_t s = h_start[i], e = h_end[i];
        s = s < 0 ? (s + n_bases > 0 ? s + n_bases : 0) : (s < n_bases ? s : n_bases);
        e = e < 0 ? (e + n_bases > 0 ? e + n_bases : 0) : (e < n_bases ? e : n_bases);
        if (e < s) e = s;
        const int64_t nw = e - s - m + 1;
        if (nw < 0) {      // len == m - 1 is an empty score list in the reference (posterior = prior)
            cudaStreamSynchronize(st);              // the upload may be in flight
            return cgbk_set_error(CGBK_ERR_INVALID, "region %d is more than one base shorter than the motif "
                                  "(negative dimensions are not allowed)", i);
        }
        st_start[i] = s;           // sequence 0 starts at packed offset 0
        st_nwin[i] = (int32_t)nw;
    }
    double* st_msm = (double*)((unsigned char*)h_staging + cgbk_pipeline_staging_bytes(n_regions) - 256);
    st_msm[0] = mu_m; st_msm[1] = std_m;

    // 1. Upload.  A small replicon goes up in ONE copy on the caller's stream (issued before the
    //    host-side preparation above, see `early`), the promoter table beside it on a side stream.
    //    A large one travels in chunks on the side stream while the caller's stream packs and
    //    scans the chunks that have arrived: only the last chunk's pack + scan is left after the
    //    last byte has crossed PCIe.
    TRACE("host prep");
    cudaError_t ce = cudaSuccess;
    for (int c = 0; c < n_chunks && !early && ce == cudaSuccess; ++c) {
        if (packed) ce = upload_planes(cut[c], cut[c + 1], cs);
        else ce = cudaMemcpyAsync(d_ascii + cut[c], h_ascii + cut[c], cut[c + 1] - cut[c], cudaMemcpyHostToDevice, cs);
        if (ce == cudaSuccess) ce = cudaEventRecord(ps->ev[c], cs);
        TRACE_ON("chunk copied", cs);
    }
    if (ce == cudaSuccess && n_regions > 0) {
        ce = cudaMemcpyAsync(base + P.off_desc, h_staging, 12 * (int64_t)n_regions, cudaMemcpyHostToDevice, cs);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(base + P.off_msm, st_msm, 16, cudaMemcpyHostToDevice, cs);
        if (ce == cudaSuccess) ce = cudaEventRecord(ps->ev[PIPE_MAX_CHUNKS + 1], cs);
    }
    if (ce != cudaSuccess) {
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(st);
        return cgbk_check_cuda(ce, "H2D of the sequence");
    }
    TRACE("h2d issued");

    // 2. background: stats scan (+ score plane) over all windows, chunk by chunk
    cgbk_seqset* S = nullptr;
    const int64_t zero = 0;
    int rc = cgbk_seqset_create(&zero, &n_bases, 1, M->device, (int64_t*)(base + P.off_table), st, &S);
    unsigned long long* d_hist = n_bins > 0 ? (unsigned long long*)(base + P.off_hist) : nullptr;
    double* d_stats = (double*)(base + P.off_stats);
    int32_t* d_plane = n_regions > 0 ? (int32_t*)(base + P.off_plane) : nullptr;
    BgScan run;
    if (!rc)
        rc = cgbk_bg_scan_begin(M, &G, S, hist_lo, hist_hi, n_bins, d_hist, d_stats, 0, d_plane, P.plane_ints,
                                base + P.off_work, P.work_bytes, st, &run);
    // Results straight into the caller's buffers when those are pinned (and therefore visible to
    // the device under unified addressing): the finalizing block of the scan stores the record
    // and the histogram, the posterior kernel its posteriors -- no D2H copy at all.  Pageable
    // buffers (or a sequence without a single window) take the copy below.
    const bool want_hist = h_hist && n_bins > 0;
    void *a_post = nullptr, *a_stats = nullptr, *a_hist = nullptr;
    const bool direct = !rc && S->tiling[CGBK_TILING_STATS].n_tiles > 0 && device_alias(h_stats, &a_stats) &&
                        (n_regions == 0 || device_alias(h_post, &a_post)) &&
                        (!want_hist || device_alias(h_hist, &a_hist));
    if (direct) {
        run.sp.host_stats = (double*)a_stats;
        run.sp.host_hist = (unsigned long long*)a_hist;
    }
    for (int c = 0; c < n_chunks && !rc; ++c) {
        if (!early && cudaStreamWaitEvent(st, ps->ev[c], 0) != cudaSuccess)
            rc = cgbk_set_error(CGBK_ERR_CUDA, "stream wait failed");
        if (!rc && !packed)
            rc = cgbk_pack_ascii(d_ascii + cut[c], cut[c + 1] - cut[c], cut[c], planes, planes + P.cap_bases / 16,
                                 planes + P.cap_bases / 16 + P.cap_bases / 32,
                                 P.cap_bases, st);
        if (!rc) rc = cgbk_bg_scan_range(&run, cut[c + 1], c == n_chunks - 1, st);
        TRACE("chunk scanned");
    }
    if (!rc) rc = cgbk_bg_scan_end(&run, st);
    if (rc) {
        // the side stream may still be writing into the caller's scratch: drain both
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(st);
        if (S) cgbk_seqset_destroy(S);
        return rc;
    }
    TRACE("stats");
    // 3. posteriors of every promoter from the plane, estimator straight from d_stats
    if (n_regions > 0) {
        if (cudaStreamWaitEvent(st, ps->ev[PIPE_MAX_CHUNKS + 1], 0) != cudaSuccess) {
            cudaStreamSynchronize(cs);
            cudaStreamSynchronize(st);
            cgbk_seqset_destroy(S);
            return cgbk_set_error(CGBK_ERR_CUDA, "stream wait failed");
        }
        rc = cgbk_posteriors(M, &G, S, d_plane, (const int64_t*)(base + P.off_desc),
                             (const int32_t*)(base + P.off_desc + 8 * (int64_t)n_regions), n_regions,
                             (const double*)(base + P.off_msm), d_stats + CGBK_BG_MEAN, p_motif, alpha,
                             direct ? (double*)a_post : (double*)(base + P.off_post), st);
        if (rc) {
            cudaStreamSynchronize(cs);
            cudaStreamSynchronize(st);
            cgbk_seqset_destroy(S);
            return rc;
        }
        TRACE("posteriors");
    }
    // 4. results back (unless the kernels stored them in place): one copy when the caller's
    //    three host buffers are laid out like ours
    const int64_t post_bytes = 8 * (int64_t)n_regions, stats_bytes = sizeof(double) * CGBK_BG_COUNT;
    if (direct) {
        // nothing to copy
    } else if (n_regions > 0 && (unsigned char*)h_stats == (unsigned char*)h_post + post_bytes &&
               (!want_hist || (unsigned char*)h_hist == (unsigned char*)h_stats + stats_bytes)) {
        ce = cudaMemcpyAsync(h_post, base + P.off_post, post_bytes + stats_bytes + (want_hist ? 8 * (int64_t)n_bins : 0),
                             cudaMemcpyDeviceToHost, st);
    } else {
        if (n_regions > 0) ce = cudaMemcpyAsync(h_post, base + P.off_post, post_bytes, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_stats, d_stats, stats_bytes, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && want_hist)
            ce = cudaMemcpyAsync(h_hist, d_hist, sizeof(unsigned long long) * n_bins, cudaMemcpyDeviceToHost, st);
    }
    TRACE("d2h");
    cudaError_t e = cudaStreamSynchronize(st);
    TRACE_DUMP();
    cgbk_seqset_destroy(S);
    if (ce != cudaSuccess) return cgbk_check_cuda(ce, "D2H of the results");
    if (e != cudaSuccess) return cgbk_check_cuda(e, "regulation pipeline");
    return CGBK_OK;
}

extern "C" int cgbk_regulation_pipeline(cgbk_model* M, const uint8_t* h_ascii, int64_t n_bases,
                                        const int64_t* h_start, const int64_t* h_end, int n_regions,
                                        double mu_m, double std_m, double p_motif, double alpha,
                                        double hist_lo, double hist_hi, int n_bins,
                                        void* d_scratch, int64_t scratch_bytes,
                                        void* h_staging, int64_t staging_bytes,
                                        double* h_post, double* h_stats, unsigned long long* h_hist,
                                        void* stream) {
    if (!h_ascii) return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    const PipeSource src = {h_ascii, nullptr, nullptr, nullptr};
    return pipeline_impl(M, src, n_bases, h_start, h_end, n_regions, mu_m, std_m, p_motif, alpha, hist_lo, hist_hi,
                         n_bins, d_scratch, scratch_bytes, h_staging, staging_bytes, h_post, h_stats, h_hist, stream);
}

extern "C" int cgbk_regulation_pipeline_packed(cgbk_model* M, const uint32_t* h_bases2, const uint32_t* h_amb,
                                               const uint32_t* h_lower, int64_t n_bases,
                                               const int64_t* h_start, const int64_t* h_end, int n_regions,
                                               double mu_m, double std_m, double p_motif, double alpha,
                                               double hist_lo, double hist_hi, int n_bins,
                                               void* d_scratch, int64_t scratch_bytes,
                                               void* h_staging, int64_t staging_bytes,
                                               double* h_post, double* h_stats, unsigned long long* h_hist,
                                               void* stream) {
    if (!h_bases2) return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    const PipeSource src = {nullptr, h_bases2, h_amb, h_lower};
    return pipeline_impl(M, src, n_bases, h_start, h_end, n_regions, mu_m, std_m, p_motif, alpha, hist_lo, hist_hi,
                         n_bins, d_scratch, scratch_bytes, h_staging, staging_bytes, h_post, h_stats, h_hist, stream);
}

extern "C" int cgbk_regulation_pipeline_args(const cgbk_pipeline_args* a) {
    if (!a) return cgbk_set_error(CGBK_ERR_INVALID, "NULL argument struct");
    const PipeSource src = {a->h_ascii, a->h_ascii ? nullptr : a->h_bases2, a->h_ascii ? nullptr : a->h_amb,
                            a->h_ascii ? nullptr : a->h_lower};
    return pipeline_impl(a->model, src, a->n_bases, a->h_start, a->h_end, a->n_regions, a->mu_m,
                         a->std_m, a->p_motif, a->alpha, a->hist_lo, a->hist_hi, a->n_bins,
                         a->d_scratch, a->scratch_bytes, a->h_staging, a->staging_bytes, a->h_post,
                         a->h_stats, a->h_hist, a->stream);
}
