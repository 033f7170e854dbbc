// Exact kernels: packing, per-window float64 scoring in the reference's order,
// candidate re-scoring, ambiguous ("dirty") tiles, posteriors, moment finalisation.
//
// These kernels carry the bit-exact semantics of the reference
// (cgb/pssm_model.py:88-133, cgb/binding_model.py:94-113, cgb/genome.py:433-445);
// the throughput kernels live in cgbk_fast.cu and hand their candidates here.
#include "cgbk_device.cuh"

// ---------------------------------------------------------------------------
// pack: ASCII -> 2-bit bases + ambiguity / lower-case bit planes
//   replaces "str(record.seq)" (cgb/chromid.py:57-60) as the sequence store.
//   One thread packs 32 bases: 32 B in, 8 B + 4 B + 4 B out (all coalesced).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pack4(uint32_t w, uint32_t& codes8, uint32_t& amb4, uint32_t& low4) {
    const uint32_t t = ((w >> 1) ^ (w >> 2)) & 0x03030303u;        // A0 C1 G2 T3 per byte
    codes8 = (t | (t >> 6) | (t >> 12) | (t >> 18)) & 0xffu;
    const uint32_t u = w & 0xdfdfdfdfu;                             // fold case
    const uint32_t valid = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) |
                           __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    const uint32_t gather = (1u << 24) | (1u << 17) | (1u << 10) | (1u << 3);
    amb4 = (((~valid) & 0x01010101u) * gather >> 24) & 0xfu;
    low4 = ((((w >> 5) & valid) & 0x01010101u) * gather >> 24) & 0xfu;
    // ambiguous bytes pack as base 0 so the planes are deterministic
    uint32_t keep = (valid & 0x01010101u) * gather >> 24;           // 1 bit per byte
    uint32_t keep2 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) keep2 |= ((keep >> k) & 1u) * (3u << (2 * k));
    codes8 &= keep2;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ ascii, long long n,
                                                   long long dst_off, uint32_t* __restrict__ b2,
                                                   uint32_t* __restrict__ amb,
                                                   uint32_t* __restrict__ lower, long long n_chunks) {
    const bool aligned = (reinterpret_cast<uintptr_t>(ascii) & 15) == 0;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_chunks;
         c += (long long)gridDim.x * blockDim.x) {
        const long long base = c * 32;
        uint32_t w[8];
        if (aligned && base + 32 <= n) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(ascii + base));
            const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(ascii + base) + 1);
            w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
            w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const long long i = base + k * 4 + b;
                    // past the end: pad with 'A' (packs to 0, not ambiguous)
                    const uint32_t ch = (i < n) ? (uint32_t)ascii[i] : 0x41u;
                    v |= ch << (8 * b);
                }
                w[k] = v;
            }
        }
        uint32_t lo = 0, hi = 0, am = 0, lw = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t c8, a4, l4;
            pack4(w[k], c8, a4, l4);
            if (k < 4) lo |= c8 << (8 * k); else hi |= c8 << (8 * (k - 4));
            am |= a4 << (4 * k);
            lw |= l4 << (4 * k);
        }
        const long long chunk = (dst_off >> 5) + c;
        reinterpret_cast<uint2*>(b2)[chunk] = make_uint2(lo, hi);
        amb[chunk] = am;
        lower[chunk] = lw;
    }
}

cudaError_t launch_pack(const uint8_t* d_ascii, long long n, long long dst_off, uint32_t* b2,
                        uint32_t* amb, uint32_t* lower, cudaStream_t st) {
    long long padded = ((n + CGBK_SEQ_ALIGN - 1) / CGBK_SEQ_ALIGN) * CGBK_SEQ_ALIGN;
    if (padded == 0) padded = CGBK_SEQ_ALIGN;      // an empty sequence still owns (and zeroes) one block
    const long long n_chunks = padded / 32;
    long long blocks = (n_chunks + 255) / 256;
    const long long cap = (long long)cgbk_current_sms() * 32;
    if (blocks > cap) blocks = cap;
    pack_kernel<<<(int)blocks, 256, 0, st>>>(d_ascii, n, dst_off, b2, amb, lower, n_chunks);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// score_windows: PSSMModel.score_seq (cgb/pssm_model.py:103-133), one thread per window
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_windows_kernel(
    const double* __restrict__ d_exact, int m, GenomeView g, const long long* __restrict__ d_start,
    const long long* __restrict__ d_out_off, int n_regions, long long total, int flags,
    float* __restrict__ f32, float* __restrict__ r32, double* __restrict__ both,
    double* __restrict__ f64, double* __restrict__ r64) {
    __shared__ ExactTables tab;
    stage_exact(&tab, d_exact);
    __syncthreads();
    const bool revseq = (flags & CGBK_RC_REVSEQ_ORDER) != 0;
    const bool single64 = (flags & CGBK_SINGLE_WINDOW_F64) != 0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_regions;     // d_out_off[lo] <= t < d_out_off[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(d_out_off + mid) <= t) lo = mid; else hi = mid;
        }
        const long long first = __ldg(d_out_off + lo);
        const long long nwin = __ldg(d_out_off + lo + 1) - first;
        const long long p = __ldg(d_start + lo) + (t - first);
        double f, r;
        const bool repaired = exact_window(g, &tab, m, p, revseq, f, r);
        const float ff = (float)f, rf = (float)r;      // Biopython stores float32
        double fs = (double)ff, rs = (double)rf;
        if (single64 && nwin == 1) {
            // len(seq)==m: scalars wrapped in Python lists, a repaired score stays float64
            if (repaired) { fs = f; rs = r; }
            if (f64) f64[lo] = fs;
            if (r64) r64[lo] = rs;
        }
        if (f32) f32[t] = ff;
        if (r32) r32[t] = rf;
        if (both) both[t] = softmax2_f64(fs, rs);
    }
}

cudaError_t launch_score_windows(const cgbk_model* mdl, GenomeView g, const long long* d_start,
                                 const long long* d_out_off, int n_regions, long long total,
                                 int flags, float* f32, float* r32, double* both, double* f64,
                                 double* r64, cudaStream_t st) {
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)mdl->sms * 16) blocks = (long long)mdl->sms * 16;
    score_windows_kernel<<<(int)blocks, 256, 0, st>>>(mdl->d_exact, mdl->m, g, d_start, d_out_off,
                                                     n_regions, total, flags, f32, r32, both, f64,
                                                     r64);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// posteriors: TFBindingModel.binding_probability (cgb/binding_model.py:94-113)
// one warp per promoter region, warp-level sum over window positions
//
//   log(lh_bg) - log(lh_m) = -log(1 - alpha) - log1p(x),
//   x = alpha/(1-alpha) * pdf_m/pdf_bg = c * exp((yb^2 - ym^2)/2),  c = alpha/(1-alpha) * std_bg/std_m
// The ratio form needs one exp + one log1p per window instead of the two pdfs and two logs
// of binding_model.py:106-112; it differs from the literal form only where a pdf underflows
// to 0 (|z| > 38).
//   Re-score path (no plane): every window is scored exactly; x below 1e-4 takes log1pf.
//   Plane path: the window's soft-max score comes from the stats scan as a fixed-point integer
//   (|error| ~ 1e-6), and log2 x is a quadratic in it (two DFMA).  Windows with x < 1e-2 --
//   all but the rare high-scoring ones -- take a float32 series for log1p (|error| < 3e-9 per
//   window, plus x * 1.4 * 1e-6 from the fixed-point score); the others are re-scored exactly.
// ---------------------------------------------------------------------------
#ifndef CGBK_POST_PF
#define CGBK_POST_PF 2      // plane loads in flight per lane (1, 2, 4, 8 measured: 10.1, 9.4, 9.7, 10.6 us)
#endif

__device__ __forceinline__ float ex2_approx_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool PLANE>
__global__ void __launch_bounds__(256) posterior_kernel(
    const double* __restrict__ d_exact, int m, GenomeView g, SeqTable seqs, PlaneView pv,
    const long long* __restrict__ d_start, const int* __restrict__ d_nwin, int n_regions,
    const double* __restrict__ ms_m, const double* __restrict__ ms_bg, double p_motif, double alpha,
    double nlog1m_alpha /* -log1p(-alpha) */, double* __restrict__ post) {
    // the re-score path stages the exact tables in shared memory; with a plane only the rare
    // high-scoring windows need them and they are read in place (device ExactTables, L1 resident)
    __shared__ ExactTables tab_s;
    __shared__ double kc[8];
    const ExactTables* tab = reinterpret_cast<const ExactTables*>(d_exact + CGBK_EXACT_TAB_OFFSET);
    if (!PLANE) {
        stage_exact(&tab_s, d_exact);
        tab = &tab_s;
    }
    if (threadIdx.x == 0) {
        // estimator constants, once per block
        const double mu_m = ms_m[0], std_m = ms_m[1], mu_bg = ms_bg[0], std_bg = ms_bg[1];
        const double im = 1.0 / std_m, ib = 1.0 / std_bg, c = alpha / (1 - alpha) * (std_bg / std_m);
        kc[0] = im; kc[1] = ib; kc[2] = c;
        // plane path: log2(x) as a quadratic in the fixed-point score sfix (s = sfix * inv_scale):
        //   log2 x = log2 c + log2(e) * ((s - mu_bg)^2 / (2 std_bg^2) - (s - mu_m)^2 / (2 std_m^2))
        const double log2e = 1.4426950408889634074;
        const double q2 = 0.5 * (ib * ib - im * im), q1 = -(mu_bg * ib * ib - mu_m * im * im);
        const double q0 = 0.5 * (mu_bg * mu_bg * ib * ib - mu_m * mu_m * im * im);
        kc[3] = log2e * q2 * pv.inv_scale * pv.inv_scale;
        kc[4] = log2e * q1 * pv.inv_scale;
        kc[5] = log2e * q0 + log2(c);
    }
    __syncthreads();
    const double mu_m = ms_m[0], mu_bg = ms_bg[0];
    const double inv_std_m = kc[0], inv_std_bg = kc[1], c = kc[2];
    const double A2 = kc[3], A1 = kc[4], A0 = kc[5];
    const float cf = (float)c;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int stride = 32 * pv.L - 32;
    const long long plane_tile = (long long)(pv.L / 32 + 1) * 1024;
    // one window, exactly: float64 strand sums in the reference's order, double exp / log1p
    auto exact_term = [&](long long gpos) {
        double f, r;
        exact_window(g, tab, m, gpos, false, f, r);
        const double se = softmax2_f32tail((float)f, (float)r);     // Biopython's float32 strand scores
        const double ym = (se - mu_m) * inv_std_m, yb = (se - mu_bg) * inv_std_bg;
        return log1p(c * exp(0.5 * (yb * yb - ym * ym)));
    };
    for (int reg = blockIdx.x * warps_per_block + (threadIdx.x >> 5); reg < n_regions;
         reg += gridDim.x * warps_per_block) {
        const long long start = __ldg(d_start + reg);
        const int nwin = __ldg(d_nwin + reg);
        double acc = 0.0;       // sum of log1p(x_i)
        if (PLANE) {
            const int sq = find_seq_by_offset(seqs, start);
            const long long x0 = start - seq_off_of(seqs, sq);
            long long tile = x0 < 0x7fffffffll ? (long long)((unsigned)x0 / (unsigned)stride) : x0 / stride;
            int wt = (int)(x0 - tile * stride) + lane;            // this lane's window, tile-local
            tile += seq_prefix_of(seqs, sq);
            if (wt >= stride) { wt -= stride; ++tile; }
            const int* tp = pv.plane + tile * plane_tile;
            int ln = wt / pv.L, within = wt - ln * pv.L;          // scan lane and position in its segment
            // Pass 1: every window from the plane on the float32 path; windows whose mixture
            // ratio x reaches 1e-2 are only marked (bit k = this lane's k-th window).  The plane
            // row of a window at position `within` of its scan lane's segment is within + D.
            unsigned long long marked = 0ull;
            constexpr int PF = CGBK_POST_PF;
            for (int base = 0, k0 = 0; base < nwin; base += 32 * PF, k0 += PF) {
                int sfix[PF];
#pragma unroll
                for (int j = 0; j < PF; ++j) {              // all loads first: independent L2 round trips
                    const int off = ((within + pv.D) << 5) + ln;
                    sfix[j] = base + lane + 32 * j < nwin ? __ldg(tp + off) : 0;
                    wt += 32; within += 32;
                    if (within >= pv.L) { within -= pv.L; ++ln; }
                    if (wt >= stride) { wt -= stride; tp += plane_tile; ln = 0; within = wt; }
                }
                float part = 0.0f;
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    const int w = base + lane + 32 * j;
                    const double sd = (double)sfix[j];
                    float t = (float)(fma(fma(A2, sd, A1), sd, A0));      // log2 x
                    if (w >= nwin) t = -200.0f;
                    if (t < -6.6438562f) {                               // x < 1e-2
                        // log1p(x) = x - x^2/2 + x^3/3 - ... : |error| < 3e-9 for x < 1e-2
                        const float xf = ex2_approx_f(t);                // 0 below 2^-126
                        part += xf * (1.0f - xf * (0.5f - xf * (1.0f / 3.0f)));
                    } else if (k0 + j < 64) {
                        marked |= 1ull << (k0 + j);
                    } else {
                        acc += exact_term(start + w);                    // regions longer than 2048 windows
                    }
                }
                acc += (double)part;
            }
            // pass 2: the marked windows, exactly
            while (marked) {
                const int kk = __ffsll((long long)marked) - 1;
                marked &= marked - 1;
                acc += exact_term(start + lane + 32 * kk);
            }
        } else {
            for (int w = lane; w < nwin; w += 32) {
                double f, r;
                exact_window(g, tab, m, start + w, false, f, r);
                const double s = softmax2_f32tail((float)f, (float)r);   // Biopython's float32 strand scores
                const double ym = (s - mu_m) * inv_std_m, yb = (s - mu_bg) * inv_std_bg;
                const double e = 0.5 * (yb * yb - ym * ym);
                const float ef = (float)e;
                double term = 0.0;                              // ef < -60: x < 1e-26
                if (!(ef < -60.0f)) {
                    const float xf = cf * expf(ef);
                    if (xf < 1e-4f) term = (double)log1pf(xf);
                    else term = log1p(c * exp(e));
                }
                acc += term;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            double res;
            if (nwin < 0) {
                res = nan("");      // the reference raises inside Biopython for len(seq) < m - 1
            } else {
                const double lh_ratio = exp((double)nwin * nlog1m_alpha - acc);
                res = 1 / (1 + lh_ratio * (1 - p_motif) / p_motif);
            }
            post[reg] = res;
        }
    }
}

cudaError_t launch_posteriors(const cgbk_model* mdl, GenomeView g, SeqTable seqs, PlaneView pv,
                              const long long* d_start, const int* d_nwin, int n_regions,
                              const double* d_ms_m, const double* d_ms_bg, double p_motif,
                              double alpha, double* d_post, cudaStream_t st) {
    if (n_regions <= 0) return cudaSuccess;
    // one warp per promoter, 4 warps per block: a few thousand promoters are all resident at once
    // (no partial second wave)
    int blocks = (n_regions + 3) / 4;
    if (blocks > mdl->sms * 16) blocks = mdl->sms * 16;
    if (pv.plane)
        posterior_kernel<true><<<blocks, 128, 0, st>>>(mdl->d_exact, mdl->m, g, seqs, pv, d_start, d_nwin,
                                                      n_regions, d_ms_m, d_ms_bg, p_motif, alpha,
                                                      -log1p(-alpha), d_post);
    else
        posterior_kernel<false><<<blocks, 128, 0, st>>>(mdl->d_exact, mdl->m, g, seqs, pv, d_start, d_nwin,
                                                       n_regions, d_ms_m, d_ms_bg, p_motif, alpha,
                                                       -log1p(-alpha), d_post);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// hits: exact re-scoring of filter candidates + exact scan of dirty tiles
//   genome.py:436,443: float32 score compared with the float64 threshold
// ---------------------------------------------------------------------------
// Warp-convergent append of up to two hits per lane (+ strand, - strand) with ONE global atomic
// per warp: with low-information motifs millions of hits would otherwise serialise on the counter.
__device__ __forceinline__ void emit_hits_warp(cgbk_hit* hits, long long cap, long long* counters,
                                               long long gpos, bool hf, float ff, bool hr, float rf) {
    const unsigned int mf = __ballot_sync(0xffffffffu, hf), mr = __ballot_sync(0xffffffffu, hr);
    const int nf = __popc(mf), tot = nf + __popc(mr);
    if (tot == 0) return;                                   // warp-uniform
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0)
        base = atomicAdd(reinterpret_cast<unsigned long long*>(counters + CGBK_CTR_HITS),
                         (unsigned long long)tot);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned int lt = (1u << lane) - 1u;
    if (hf) {
        const long long k = (long long)base + __popc(mf & lt);
        if (k < cap) { cgbk_hit h; h.gpos = gpos; h.strand = 1; h.score = ff; hits[k] = h; }
    }
    if (hr) {
        const long long k = (long long)base + nf + __popc(mr & lt);
        if (k < cap) { cgbk_hit h; h.gpos = gpos; h.strand = -1; h.score = rf; hits[k] = h; }
    }
}

__global__ void __launch_bounds__(128) rescore_kernel(const double* __restrict__ d_exact, int m,
                                                      GenomeView g, FastWork w, double thr,
                                                      int flags, cgbk_hit* hits, long long hit_cap) {
    __shared__ ExactTables tab;
    stage_exact(&tab, d_exact);
    __syncthreads();
    const bool revseq = (flags & CGBK_RC_REVSEQ_ORDER) != 0;
    long long n = w.counters[CGBK_CTR_CANDIDATES];
    if (n > w.cand_cap) n = w.cand_cap;
    // warp-uniform trip count so that the hits of a warp are appended with one atomic
    for (long long i0 = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); i0 < n;
         i0 += (long long)gridDim.x * blockDim.x) {
        const long long i = i0 + (threadIdx.x & 31);
        bool hf = false, hr = false;
        long long gpos = 0;
        float ff = 0.f, rf = 0.f;
        if (i < n) {
            const unsigned long long rec = w.cands[i];
            gpos = (long long)(rec & 0x3fffffffffffffffull);
            double f, r;
            exact_window(g, &tab, m, gpos, revseq, f, r);
            ff = (float)f; rf = (float)r;
            hf = (rec >> 63) && (double)ff >= thr;
            hr = ((rec >> 62) & 1ull) && (double)rf >= thr;
        }
        emit_hits_warp(hits, hit_cap, w.counters, gpos, hf, ff, hr, rf);
    }
}

// tile geometry shared with cgbk_fast.cu: a warp tile covers 32*L lookup positions and
// owns the windows [A, A + 32*L - 32) of its sequence.
__global__ void __launch_bounds__(128) dirty_hits_kernel(const double* __restrict__ d_exact, int m,
                                                         GenomeView g, SeqTable seqs, FastWork w,
                                                         int L, double thr, int flags,
                                                         cgbk_hit* hits, long long hit_cap) {
    __shared__ ExactTables tab;
    stage_exact(&tab, d_exact);
    __syncthreads();
    const bool revseq = (flags & CGBK_RC_REVSEQ_ORDER) != 0;
    const long long nd = w.counters[CGBK_CTR_DIRTY_TILES];
    for (long long d = blockIdx.x; d < nd; d += gridDim.x) {
        long long gbase, lo, hi;
        tile_range(seqs, w.dirty_tiles[d], L, m, gbase, lo, hi);
        for (long long x0 = lo + (threadIdx.x & ~31); x0 < hi; x0 += blockDim.x) {
            const long long x = x0 + (threadIdx.x & 31);
            bool hf = false, hr = false;
            float ff = 0.f, rf = 0.f;
            if (x < hi) {
                double f, r;
                exact_window(g, &tab, m, gbase + x, revseq, f, r);
                ff = (float)f; rf = (float)r;
                hf = (double)ff >= thr;
                hr = (double)rf >= thr;
            }
            emit_hits_warp(hits, hit_cap, w.counters, gbase + x, hf, ff, hr, rf);
        }
    }
}

cudaError_t launch_rescore(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w,
                           double thr, int flags, cgbk_hit* hits, long long hit_cap, int L,
                           cudaStream_t st) {
    rescore_kernel<<<mdl->sms * 2, 128, 0, st>>>(mdl->d_exact, mdl->m, g, w, thr, flags, hits, hit_cap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dirty_hits_kernel<<<mdl->sms * 2, 128, 0, st>>>(mdl->d_exact, mdl->m, g, seqs, w, L, thr, flags, hits,
                                                   hit_cap);
    return cudaGetLastError();
}

// Fused scan, site pass (cgbk_scan_batch): the windows flagged in the scan's flag plane are
// scored with the exact arithmetic and compared like rescore_kernel does (genome.py:436,443), and
// the sites come out ORDERED by (position, + strand first) without a sort:
//   1. exclusive prefix sum of the scan's per-tile flag counts        -> where a tile's candidates go
//   2. site_score_kernel, one warp per tile: the tile's flagged windows are listed in position
//      order (lanes own consecutive segments, rows and bits ascend inside a lane), dealt out to the
//      lanes 32 at a time (balanced however unevenly the flags fall) and scored; each candidate
//      leaves a (+ strand record, - strand record) pair, ~0 where that strand is no site
//   3. exclusive prefix sum of the per-tile site counts               -> where a tile's sites go
//   4. site_emit_kernel, one warp per tile: the pairs are compacted into the batch's hit list
//   5. site_finish_kernel: counters and the batch's append position
#define CGBK_SITE_QUEUE 256     // candidates listed per warp and round
#define CGBK_NO_SITE 0xffffffffffffffffull

// words of a warp's staged tile: 32 L + 32 bases (the last window's far end) rounded up, + 1 for
// the funnel shift's look-ahead
#define CGBK_SITE_B2_WORDS (2 * 128 + 4)
#define CGBK_SITE_AMB_WORDS (128 + 3)
#define CGBK_SITE_GROUP 8        // consecutive tiles a warp takes together
#define CGBK_SITE_DENSE 32       // a tile with at least this many candidates is scored on its own

// Per-warp scratch of the site pass.
struct SiteWarp {
    unsigned short queue[CGBK_SITE_QUEUE];
    uint32_t b2[CGBK_SITE_B2_WORDS];
    uint32_t amb[CGBK_SITE_AMB_WORDS];
    long long p0[CGBK_SITE_GROUP];       // first base of the group's tiles
    long long cand_off[CGBK_SITE_GROUP]; // where a tile's candidate pairs start
    int qstart[CGBK_SITE_GROUP];         // where a sparse tile's entries start in the queue
    int hits[CGBK_SITE_GROUP];
};

// A warp takes CGBK_SITE_GROUP consecutive tiles at a time.  Motif sets are mostly SPARSE in
// candidates (two thirds of the JASPAR-scale benchmark set have fewer than 32 per 4 096-window
// tile, 40 % fewer than 4): scoring tile by tile would run the 32-lane exact scorer for three or
// four candidates.  So the group's sparse tiles pool their candidates in one queue (entry = tile
// within the group << 12 | window within the tile; 8 x 31 entries at most) that is scored with
// all lanes busy, straight from the packed genome; a dense tile is scored on its own from a
// shared-memory copy of its bases (one coalesced load per tile instead of five scattered loads
// per candidate).  A candidate's place in the pair list depends only on its tile and its rank in
// the tile, so the order of the output is the same either way.
__global__ void __launch_bounds__(128) site_score_kernel(const double* __restrict__ d_exact, int m, GenomeView g,
                                                         SeqTable seqs, FastWork w, SitePass sp, int L, int D,
                                                         double thr, int flags) {
    __shared__ ExactTables tab;
    __shared__ SiteWarp scratch[4];
    stage_exact(&tab, d_exact);
    __syncthreads();
    const bool revseq = (flags & CGBK_RC_REVSEQ_ORDER) != 0;
    const int B = L >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long n_tiles = w.tile_end - w.tile_begin;
    const long long n_groups = (n_tiles + CGBK_SITE_GROUP - 1) / CGBK_SITE_GROUP;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    SiteWarp& sw = scratch[wib];
    unsigned short* q = sw.queue;
    const long long nw2 = g.cap_bases >> 4, nwa = g.cap_bases >> 5;
    const uint32_t maskm = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    const int n_b2 = 2 * L + 4, n_amb = L + 3;
    const int S = lane * L;

    // one scored candidate -> its pair of records; returns the number of sites (0..2)
    auto emit = [&](long long k, long long gpos, double f, double r) -> int {
        const float ff = (float)f, rf = (float)r;
        const bool hf = (double)ff >= thr, hr = (double)rf >= thr;
        if (k < sp.cand_cap)                  // both records in one 16-byte store
            *reinterpret_cast<ulonglong2*>(sp.pairs + 2 * k) =
                make_ulonglong2(hf ? cgbk_pack_hit8(gpos, 0, __float_as_uint(ff)) : CGBK_NO_SITE,
                                hr ? cgbk_pack_hit8(gpos, 1, __float_as_uint(rf)) : CGBK_NO_SITE);
        return (hf ? 1 : 0) + (hr ? 1 : 0);
    };

    for (long long grp = warp0; grp < n_groups; grp += nwarps) {
        const long long tile0 = w.tile_begin + grp * CGBK_SITE_GROUP;
        const int nt = (int)(w.tile_end - tile0 < CGBK_SITE_GROUP ? w.tile_end - tile0 : CGBK_SITE_GROUP);
        const int my_total = lane < nt ? __ldg(sp.tile_cand + tile0 + lane) : 0;
        if (!__any_sync(0xffffffffu, my_total > 0)) {
            if (lane < nt) sp.tile_hits[tile0 + lane] = 0;
            continue;
        }
        if (lane < CGBK_SITE_GROUP) {
            sw.hits[lane] = 0;
            sw.cand_off[lane] = lane < nt ? (long long)__ldg(sp.tile_cand_off + tile0 + lane) : 0;
        }
        __syncwarp();
        int qn = 0;                                         // pooled (sparse) queue entries, warp-uniform
        for (int t = 0; t < nt; ++t) {
            const int total = __shfl_sync(0xffffffffu, my_total, t);
            if (total == 0) continue;                       // warp-uniform
            const long long tile = tile0 + t;
            uint32_t words[5];
            int c = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                words[r] = r <= B ? __ldg(w.fplane + (tile * (B + 1) + r) * 32 + lane) : 0u;
                c += __popc(words[r]);
            }
            int off = c;                                    // inclusive scan over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, off, o);
                if (lane >= o) off += v;
            }
            off -= c;
            long long gbase, w_lo, w_hi;
            tile_range(seqs, (unsigned int)tile, L, m, gbase, w_lo, w_hi);
            const long long p0 = gbase + w_lo;              // a multiple of 32
            // tile-local window of flag bit p in row r of this lane
            auto window_of = [&](int r, int p) -> int {
                return r < B ? S + 32 * r + (31 - p) - D : S + L - D + (D - 1 - p);
            };
            if (total < CGBK_SITE_DENSE) {
                // sparse tile: pool its candidates
                int idx = qn + off;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    uint32_t word = words[r];
                    while (word) {
                        const int p = 31 - __clz(word);
                        word &= ~(1u << p);
                        q[idx++] = (unsigned short)((t << 12) | window_of(r, p));
                    }
                }
                if (lane == 0) { sw.qstart[t] = qn; sw.p0[t] = p0; }
                qn += total;
                continue;
            }
            // dense tile: bases and ambiguity bits staged once (coalesced), candidates in rounds
            __syncwarp();
            for (int i = lane; i < n_b2; i += 32) sw.b2[i] = ldw(g.bases2, (p0 >> 4) + i, nw2);
            for (int i = lane; i < n_amb; i += 32) sw.amb[i] = ldw(g.amb, (p0 >> 5) + i, nwa);
            // the pooled entries share the queue: dense rounds use its upper part
            unsigned short* dq = q + qn;
            const int room = CGBK_SITE_QUEUE - qn;          // >= 256 - 7 * 31
            // a dense tile's sites are compacted right here: the candidates of a round are in position
            // order across the lanes, so a ballot gives every record its place; they go, packed, to the
            // head of the tile's stretch of the pair buffer (8 bytes per SITE instead of 16 per
            // candidate written here and read back by site_emit_kernel, which then only copies)
            const long long base = 2 * sw.cand_off[t];
            const unsigned int lt = (1u << lane) - 1u;
            int n_hits = 0;                                 // warp-uniform
            int idx = off;                                  // rank (in the tile) of this lane's next candidate
            for (int chunk0 = 0; chunk0 < total; chunk0 += room) {
                // every flag bit is popped once: a lane resumes where the previous round stopped it
                // (its candidates are consecutive ranks, rows and bits in position order)
                const int limit = chunk0 + room;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    while (words[r] && idx < limit) {
                        const int p = 31 - __clz(words[r]);
                        words[r] &= ~(1u << p);
                        dq[idx - chunk0] = (unsigned short)window_of(r, p);
                        ++idx;
                    }
                }
                __syncwarp();
                const int n = total - chunk0 < room ? total - chunk0 : room;
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    bool hf = false, hr = false;
                    float ff = 0.0f, rf = 0.0f;
                    long long gpos = 0;
                    if (i < n) {
                        const int wt = dq[i];
                        gpos = p0 + wt;
                        const int wi = wt >> 4, sh = (wt & 15) * 2;
                        const uint32_t w0 = sw.b2[wi], w1 = sw.b2[wi + 1], w2 = sw.b2[wi + 2];
                        const uint64_t x = (uint64_t)__funnelshift_r(w0, w1, sh) |
                                           ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
                        const uint32_t amb = __funnelshift_r(sw.amb[wt >> 5], sw.amb[(wt >> 5) + 1], (uint32_t)(wt & 31)) & maskm;
                        double f, r;
                        exact_window_bits(g, &tab, m, gpos, x, amb, revseq, f, r);
                        ff = (float)f; rf = (float)r;
                        hf = (double)ff >= thr; hr = (double)rf >= thr;
                    }
                    const unsigned int mf = __ballot_sync(0xffffffffu, hf), mr = __ballot_sync(0xffffffffu, hr);
                    // records of the lanes before this one, then its own + strand, then its - strand
                    const long long k = base + n_hits + __popc(mf & lt) + __popc(mr & lt);
                    if (hf && k < 2 * sp.cand_cap) sp.pairs[k] = cgbk_pack_hit8(gpos, 0, __float_as_uint(ff));
                    if (hr && k + (hf ? 1 : 0) < 2 * sp.cand_cap)
                        sp.pairs[k + (hf ? 1 : 0)] = cgbk_pack_hit8(gpos, 1, __float_as_uint(rf));
                    n_hits += __popc(mf) + __popc(mr);
                }
                __syncwarp();
            }
            if (lane == 0) sw.hits[t] = n_hits;
        }
        __syncwarp();
        // the pooled candidates of the group's sparse tiles, all lanes busy
        for (int i = lane; i < qn; i += 32) {
            const int e = q[i], t = e >> 12, wt = e & 4095;
            const long long gpos = sw.p0[t] + wt;
            double f, r;
            exact_window(g, &tab, m, gpos, revseq, f, r);
            const int nh = emit(sw.cand_off[t] + (i - sw.qstart[t]), gpos, f, r);
            if (nh) atomicAdd(&sw.hits[t], nh);
        }
        __syncwarp();
        if (lane < nt) sp.tile_hits[tile0 + lane] = sw.hits[lane];
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) site_emit_kernel(FastWork w, SitePass sp) {
    const int lane = threadIdx.x & 31;
    const long long n_tiles = w.tile_end - w.tile_begin;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long cursor = *w.hit_cursor;
    for (long long ti = warp0; ti < n_tiles; ti += nwarps) {
        const long long tile = w.tile_begin + ti;
        const int n_cand = __ldg(sp.tile_cand + tile);
        if (n_cand == 0) continue;
        const long long in0 = 2ll * __ldg(sp.tile_cand_off + tile);
        long long out = cursor + __ldg(sp.tile_hit_off + tile);
        if (n_cand >= CGBK_SITE_DENSE) {
            // a dense tile's sites were compacted by the scoring warp: a straight copy
            // (eight independent loads in flight per lane, then the stores)
            long long n_site = __ldg(sp.tile_hits + tile);
            if (n_site > 2 * sp.cand_cap - in0) n_site = 2 * sp.cand_cap - in0;
            if (n_site > w.hit_cap - out) n_site = w.hit_cap - out;
            const unsigned long long* src = sp.pairs + in0;
            unsigned long long* dst = w.hits8 + out;
            for (long long i0 = lane; i0 < n_site; i0 += 256) {
                unsigned long long v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = i0 + 32 * u < n_site ? __ldcs(src + i0 + 32 * u) : 0ull;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (i0 + 32 * u < n_site) dst[i0 + 32 * u] = v[u];
            }
            continue;
        }
        const long long n_rec = 2ll * n_cand;
        for (long long i0 = 0; i0 < n_rec; i0 += 32) {
            const long long i = i0 + lane;
            unsigned long long rec = CGBK_NO_SITE;
            if (i < n_rec && in0 + i < 2 * sp.cand_cap) rec = sp.pairs[in0 + i];
            const bool keep = rec != CGBK_NO_SITE;
            const unsigned int mask = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const long long k = out + __popc(mask & ((1u << lane) - 1u));
                if (k < w.hit_cap) w.hits8[k] = rec;
            }
            out += __popc(mask);
        }
    }
}

__global__ void site_finish_kernel(FastWork w, SitePass sp) {
    if (threadIdx.x == 0) {
        const long long last = w.tile_end - 1;
        const long long n_cand = (long long)sp.tile_cand_off[last] + sp.tile_cand[last];
        const long long n_hits = (long long)sp.tile_hit_off[last] + sp.tile_hits[last];
        w.counters[CGBK_CTR_CANDIDATES] = n_cand;
        w.counters[CGBK_CTR_HITS] = n_hits;
        *w.hit_cursor += n_hits;
    }
}

cudaError_t launch_site_pass(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, const SitePass& sp,
                             int L, double thr, int flags, cudaStream_t st) {
    const long long n_tiles = w.tile_end - w.tile_begin;
    if (n_tiles <= 0) return cudaSuccess;
    cudaError_t e = cgbk_exclusive_sum(sp.tile_cand + w.tile_begin, sp.tile_cand_off + w.tile_begin, n_tiles,
                                       sp.scan_tmp, sp.scan_tmp_bytes, st);
    if (e != cudaSuccess) return e;
    long long blocks = (n_tiles + 3) / 4;
    if (blocks > mdl->sms * 16) blocks = mdl->sms * 16;
    // the score kernel's warps take CGBK_SITE_GROUP tiles at a time
    long long sblocks = ((n_tiles + CGBK_SITE_GROUP - 1) / CGBK_SITE_GROUP + 3) / 4;
    if (sblocks > mdl->sms * 16) sblocks = mdl->sms * 16;
    site_score_kernel<<<(int)sblocks, 128, 0, st>>>(mdl->d_exact, mdl->m, g, seqs, w, sp, L, 4 * (mdl->G - 1), thr, flags);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = cgbk_exclusive_sum(sp.tile_hits + w.tile_begin, sp.tile_hit_off + w.tile_begin, n_tiles, sp.scan_tmp,
                           sp.scan_tmp_bytes, st);
    if (e != cudaSuccess) return e;
    site_emit_kernel<<<(int)blocks, 128, 0, st>>>(w, sp);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    site_finish_kernel<<<1, 32, 0, st>>>(w, sp);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// background over a region table: Genome.build_PSSM_model (cgb/genome.py:215-226)
//   every m-mer of every gene's upstream non-coding region is scored on its own,
//   score_seq(seq) with len(seq) == m (pssm_model.py:117-127: one window, a repaired score stays
//   float64), and the scores go to np.mean / np.std.  One warp per region, windows strided over
//   the lanes; float64 partial sums per block in a fixed order (deterministic), histogram in
//   shared memory.  Regions may overlap or repeat: every listed window counts.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) region_stats_kernel(
    const double* __restrict__ d_exact, int m, GenomeView g, const long long* __restrict__ d_start,
    const int* __restrict__ d_nwin, int n_regions, double hist_lo, double inv_binw, int n_bins,
    unsigned long long* __restrict__ d_hist, double* __restrict__ partials) {
    __shared__ ExactTables tab;
    __shared__ double scratch[96];
    extern __shared__ unsigned int shist[];
    stage_exact(&tab, d_exact);
    if (d_hist)
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x) shist[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int reg = blockIdx.x * wpb + (threadIdx.x >> 5); reg < n_regions; reg += gridDim.x * wpb) {
        const long long start = __ldg(d_start + reg);
        const int nwin = __ldg(d_nwin + reg);
        for (int w = lane; w < nwin; w += 32) {
            double f, r;
            const bool repaired = exact_window(g, &tab, m, start + w, false, f, r);
            const double fs = repaired ? f : (double)(float)f, rs = repaired ? r : (double)(float)r;
            const double s = softmax2_f64(fs, rs);
            s1 += s; s2 += s * s; cnt += 1.0;
            if (d_hist) {
                int bin = (int)floor((s - hist_lo) * inv_binw);
                bin = bin < 0 ? 0 : (bin >= n_bins ? n_bins - 1 : bin);
                atomicAdd(shist + bin, 1u);
            }
        }
    }
    block_reduce3(s1, s2, cnt, scratch);
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x] = s1; partials[3 * blockIdx.x + 1] = s2; partials[3 * blockIdx.x + 2] = cnt;
    }
    if (d_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
            if (shist[i]) atomicAdd(d_hist + i, (unsigned long long)shist[i]);
    }
}

cudaError_t launch_region_stats(const cgbk_model* mdl, GenomeView g, const long long* d_start, const int* d_nwin,
                                int n_regions, double hist_lo, double hist_hi, int n_bins,
                                unsigned long long* d_hist, double* partials, int partial_cap, double* d_stats,
                                int accumulate, cudaStream_t st) {
    int blocks = (n_regions + 3) / 4;
    if (blocks > mdl->sms * 8) blocks = mdl->sms * 8;
    if (blocks > partial_cap) blocks = partial_cap;
    if (blocks > 0) {
        region_stats_kernel<<<blocks, 128, d_hist ? 4 * (size_t)n_bins : 0, st>>>(
            mdl->d_exact, mdl->m, g, d_start, d_nwin, n_regions, hist_lo, (double)n_bins / (hist_hi - hist_lo),
            n_bins, d_hist, partials);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return launch_bg_finalize(partials, blocks, d_stats, accumulate, st);
}

// ---------------------------------------------------------------------------
// background statistics: reduction of per-block partial moments (empty scans, and after the
// cross-rank all-reduce of the raw sums)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bg_finalize_kernel(const double* __restrict__ partials,
                                                          int n_partials, double* stats,
                                                          int accumulate) {
    __shared__ double scratch[96];
    finalize_stats(partials, n_partials, stats, accumulate, scratch);
}

// mean / std of many records at once (after the all-reduce of a motif batch): one block per record
__global__ void __launch_bounds__(32) bg_finalize_batch_kernel(double* stats) {
    if (threadIdx.x == 0) {
        double* s = stats + (size_t)blockIdx.x * CGBK_BG_COUNT;
        const double cnt = s[CGBK_BG_N], s1 = s[CGBK_BG_SUM], s2 = s[CGBK_BG_SUMSQ];
        const double mean = cnt > 0 ? s1 / cnt : nan("");
        double var = cnt > 0 ? s2 / cnt - mean * mean : nan("");
        if (var < 0) var = 0;
        s[CGBK_BG_MEAN] = mean;
        s[CGBK_BG_STD] = sqrt(var);
    }
}

cudaError_t launch_bg_finalize_batch(double* d_stats, int n_records, cudaStream_t st) {
    bg_finalize_batch_kernel<<<n_records, 32, 0, st>>>(d_stats);
    return cudaGetLastError();
}

cudaError_t launch_bg_finalize(const double* partials, int n_partials, double* d_stats,
                               int accumulate, cudaStream_t st) {
    bg_finalize_kernel<<<1, 256, 0, st>>>(partials, n_partials, d_stats, accumulate);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Split site records for the trip to the host: 6 bytes per site instead of 8.
// An ORDERED site list of one model (cgbk_hit8 records, what cgbk_scan_batch delivers) becomes
//   local[k]  uint16  (offset of the site inside its 4 096-base block of the packed buffer) << 1 | minus strand
//   score[k]  float32 the site's score, bit for bit
//   block_start[b], b = 0 .. n_blocks: index of the first site at or after base b * 4096 (n at the end),
//                     so block b holds the sites [block_start[b], block_start[b + 1])
// The position of site k is (its block << 12) + (local[k] >> 1): nothing is lost, the 8-byte records
// are rebuilt exactly (cgb_b200.motif_batch.SplitSites.records()).  Every site record of a large scan
// crosses PCIe, which is what bounds the end-to-end rate: 25 % fewer bytes.
// ---------------------------------------------------------------------------
#define CGBK_SITE_BLOCK_SHIFT 12

__global__ void __launch_bounds__(256) split_sites_kernel(const unsigned long long* __restrict__ hits8,
                                                          const long long* __restrict__ cum, int n_models,
                                                          long long n_single, unsigned short* __restrict__ local16,
                                                          float* __restrict__ score32, int* __restrict__ block_start,
                                                          long long n_blocks) {
    // lists of n_models models back to back: model i holds records [cum[i], cum[i + 1]) (cum == NULL:
    // one list of n_single records)
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nt = (long long)gridDim.x * blockDim.x;
    const long long n = cum ? __ldg(cum + n_models) : n_single;
    for (long long k = t0; k < n; k += nt) {
        const unsigned long long rec = __ldcs(hits8 + k);
        const unsigned long long gpos = rec >> 33;
        local16[k] = (unsigned short)(((gpos & ((1u << CGBK_SITE_BLOCK_SHIFT) - 1u)) << 1) | ((rec >> 32) & 1ull));
        score32[k] = __uint_as_float((unsigned int)rec);
    }
    // block_start[i][b] = lower bound of base b << 12 in model i's (ascending) positions, relative
    // to the start of its list
    const long long nb1 = n_blocks + 1;
    for (long long j = t0; j < nb1 * n_models; j += nt) {
        const int i = (int)(j / nb1);
        const long long b = j - (long long)i * nb1;
        const long long first = cum ? __ldg(cum + i) : 0, last = cum ? __ldg(cum + i + 1) : n_single;
        const unsigned long long want = (unsigned long long)b << CGBK_SITE_BLOCK_SHIFT;
        long long lo = first, hi = last;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if ((__ldg(hits8 + mid) >> 33) < want) lo = mid + 1; else hi = mid;
        }
        block_start[j] = (int)(lo - first);
    }
}

cudaError_t launch_split_sites(const unsigned long long* hits8, const long long* d_cum, int n_models, long long n,
                               unsigned short* local16, float* score32, int* block_start, long long n_blocks,
                               int sms, cudaStream_t st) {
    long long work = n > (n_blocks + 1) * n_models ? n : (n_blocks + 1) * n_models;
    long long blocks = (work + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    if (blocks < 1) blocks = 1;
    split_sites_kernel<<<(int)blocks, 256, 0, st>>>(hits8, d_cum, n_models, n, local16, score32, block_start, n_blocks);
    return cudaGetLastError();
}
