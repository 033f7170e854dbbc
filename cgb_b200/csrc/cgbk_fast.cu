// Tiled PSWM scan kernels for sm_100a (the throughput path).
//
// What is computed (reference: cgb/pssm_model.py:103-133 through Biopython's
// _pwm.c calculate()):  for every window i of a sequence,
//     f_i = sum_j pssm[s_{i+j}][j]          r_i = sum_j rev_comp_pssm[s_{i+j}][j]
//
// How (B200-first, not a translation of the CPU loop):
//  * the genome is a 2-bit stream in HBM; a warp tile covers 32*L consecutive
//    lookup positions, lane k owning positions [A + k*L, A + (k+1)*L);
//  * the motif is cut into G = ceil(m/4) groups of 4 positions.  For each
//    group a 256-entry table holds the partial sum of every 4-mer, for both
//    strands.  At stream position p ONE 4-mer index addresses all G groups at once
//    (entry = [group 0 .. group G-1]), fetched with LDS.128 / LDS.64 / LDS.32, and
//    group g's value is added to the accumulator of window p - 4g.  Accumulators
//    live in a 32-slot register ring whose indices are compile-time constants in
//    the fully unrolled 32-position body, so the window slides through registers;
//  * shared-memory tables are replicated per bank group (8 copies of each 16-byte
//    chunk, 16 of an 8-byte chunk, 32 of a 4-byte one), so the random 4-mer lookups
//    of a warp never bank-conflict: every chunk costs exactly 128 B/lane-group of
//    crossbar bandwidth, the LSU floor;
//  * windows that straddle two lanes are finished with warp shuffles: the first
//    D = 4(G-1) positions of a lane also accumulate the tail of the previous lane's
//    windows, and those partial sums are handed back with __shfl_down_sync;
//  * MODE 0 ("filter"): both strands ride in one 32-bit word as two biased 16-bit
//    deficits; a window whose score could reach the threshold has bit 15 / bit 31
//    set, so the per-window test is one LOP3.  Candidates go to cgbk_exact.cu for
//    float64 re-scoring in the reference's order (bit-exact hit lists);
//  * MODE 1 ("stats"): int32 fixed point per strand, soft-max log2(2^f + 2^r) on the
//    MUFU unit, per-block shared-memory histogram + float64 moments;
//  * tiles that contain a byte outside ACGTacgt are not scored here: their ids are
//    appended to a list the exact kernels walk (ambiguity repair, pssm_model.py:88-101).
#include "cgbk_device.cuh"

template <int NW>
struct TableLayout {
    static constexpr int N16 = NW / 4;
    static constexpr int HAS8 = (NW % 4) >= 2 ? 1 : 0;
    static constexpr int HAS4 = NW % 2;
    static constexpr int NCH = N16 + HAS8 + HAS4;
    static constexpr int BYTES = NCH * 32768;
};

template <int NW>
__device__ __forceinline__ void fill_tables(uint32_t* smem_words, const uint32_t* __restrict__ gtab) {
    using TL = TableLayout<NW>;
    for (int w = threadIdx.x; w < TL::NCH * 8192; w += blockDim.x) {
        const int chunk = w >> 13, rem = w & 8191, e = rem >> 5, slot = rem & 31;
        int k;
        if (chunk < TL::N16) k = chunk * 4 + (slot & 3);
        else if (TL::HAS8 && chunk == TL::N16) k = TL::N16 * 4 + (slot & 1);
        else k = NW - 1;
        smem_words[w] = __ldg(gtab + e * NW + k);
    }
}

// all NW words of table entry E (= 4-mer << 7); o16/o8/o4 = this lane's replica offsets
template <int NW>
__device__ __forceinline__ void lookup(const unsigned char* sb, uint32_t E, uint32_t o16, uint32_t o8,
                                       uint32_t o4, uint32_t (&v)[NW]) {
    using TL = TableLayout<NW>;
#pragma unroll
    for (int c = 0; c < TL::N16; ++c) {
        const uint4 q = *reinterpret_cast<const uint4*>(sb + c * 32768 + (E | o16));
        v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
    }
    if (TL::HAS8) {
        const uint2 q = *reinterpret_cast<const uint2*>(sb + TL::N16 * 32768 + (E | o8));
        v[4 * TL::N16] = q.x; v[4 * TL::N16 + 1] = q.y;
    }
    if (TL::HAS4) {
        v[NW - 1] = *reinterpret_cast<const uint32_t*>(sb + (TL::N16 + TL::HAS8) * 32768 + (E | o4));
    }
}

// (4-mer at body position t) << 7, from the body's two stream words + one lookahead word
__device__ __forceinline__ uint32_t kmer_addr(int t, uint32_t w0, uint32_t w1, uint32_t w2) {
    const uint32_t lo = t < 16 ? w0 : w1, hi = t < 16 ? w1 : w2;
    const int sh = t < 16 ? 2 * t : 2 * t - 32;
    uint32_t x;
    if (sh >= 7) x = __funnelshift_r(lo, hi, sh - 7);
    else x = lo << (7 - sh);
    return x & 0x7f80u;
}

struct StatsParams {
    float inv_scale_f;     // 2^-kexp
    double inv_scale_d;
    float hist_lo, inv_binw;
    int n_bins;
    unsigned long long* d_hist;
};

__device__ __forceinline__ uint2 ld_stream2(const uint32_t* __restrict__ b2, long long wi, long long nw) {
    if (wi + 1 < nw) return __ldg(reinterpret_cast<const uint2*>(b2 + wi));
    return make_uint2(wi < nw ? __ldg(b2 + wi) : 0u, 0u);
}

template <int G, int MODE>
struct FastCfg {
    static constexpr int THREADS = MODE == 0 ? 512 : (G <= 6 ? 384 : 256);
};

template <int G, int MODE>
__global__ void __launch_bounds__((FastCfg<G, MODE>::THREADS), 1)
fast_scan_kernel(const uint32_t* __restrict__ gtab, GenomeView g, SeqTable seqs, FastWork work, int m,
                 int B, StatsParams sp) {
    constexpr int NW = MODE == 0 ? G : 2 * G;
    constexpr int D = 4 * (G - 1);
    constexpr int DD = D > 0 ? D : 1;
    using TL = TableLayout<NW>;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned int* hist = reinterpret_cast<unsigned int*>(smem + TL::BYTES);
    __shared__ double red_scratch[96];

    fill_tables<NW>(reinterpret_cast<uint32_t*>(smem), gtab);
    if (MODE == 1)
        for (int i = threadIdx.x; i < sp.n_bins; i += blockDim.x) hist[i] = 0u;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t o16 = (lane & 7) << 4, o8 = (lane & 15) << 3, o4 = lane << 2;
    const int L = 32 * B;
    const long long stride = 32ll * L - 32;
    const long long nwords = g.cap_bases >> 4;
    double s1 = 0.0, s2 = 0.0;
    unsigned int cnt = 0;

    for (long long tile = (long long)blockIdx.x * wpb + warp; tile < seqs.n_tiles;
         tile += (long long)gridDim.x * wpb) {
        const int sq = find_seq(seqs, tile);
        const long long A = (tile - __ldg(seqs.tile_prefix + sq)) * stride;
        const long long nwin = __ldg(seqs.len + sq) - m + 1;
        const long long own_hi = A + stride < nwin ? A + stride : nwin;
        const long long gbase = __ldg(seqs.off + sq);
        const long long S = A + (long long)lane * L;           // local start of this lane's segment

        // tiles touching a non-ACGT byte go to the exact kernels
        uint32_t any_amb = 0;
        {
            const long long aw = (gbase + S) >> 5, naw = g.cap_bases >> 5;
            for (int b = 0; b < B; ++b) any_amb |= ldw(g.amb, aw + b, naw);
        }
        if (__any_sync(0xffffffffu, any_amb != 0)) {
            if (lane == 0) {
                const unsigned long long k = atomicAdd(
                    reinterpret_cast<unsigned long long*>(work.counters + CGBK_CTR_DIRTY_TILES), 1ull);
                work.dirty_tiles[k] = (unsigned int)tile;
            }
            continue;
        }

        uint32_t sa[32], sb[32], outa[DD], outb[DD];
#pragma unroll
        for (int i = 0; i < 32; ++i) { sa[i] = 0u; sb[i] = 0u; }
#pragma unroll
        for (int i = 0; i < DD; ++i) { outa[i] = 0u; outb[i] = 0u; }

        // one completed window: accumulators (a, b), local window index wl
        auto complete = [&](uint32_t a, uint32_t b, long long wl) {
            if (MODE == 0) {
                if (a & 0x80008000u) {
                    if (wl < own_hi) {
                        const unsigned long long k = atomicAdd(
                            reinterpret_cast<unsigned long long*>(work.counters + CGBK_CTR_CANDIDATES), 1ull);
                        if ((long long)k < work.cand_cap)
                            work.cands[k] = (unsigned long long)(gbase + wl) |
                                            ((unsigned long long)(a >> 31) << 63) |
                                            ((unsigned long long)((a >> 15) & 1u) << 62);
                    }
                }
            } else {
                const int fq = (int)a, rq = (int)b;
                const int mx = max(fq, rq), mn = min(fq, rq);
                const float dq = (float)(mx - mn);
                const float u = exp2f(-dq * sp.inv_scale_f);
                const float lg = __log2f(1.0f + u);
                const float sF = fmaf((float)mx, sp.inv_scale_f, lg);
                int bin = __float2int_rd((sF - sp.hist_lo) * sp.inv_binw);
                bin = min(max(bin, 0), sp.n_bins - 1);
                const double sD = fma((double)mx, sp.inv_scale_d, (double)lg);
                if (wl < own_hi) {
                    atomicAdd(hist + bin, 1u);
                    s1 += sD;
                    s2 = fma(sD, sD, s2);
                    cnt += 1u;
                }
            }
        };

        const long long wbase = (gbase + S) >> 4;
        uint2 cur = ld_stream2(g.bases2, wbase, nwords);
        for (int b = 0; b < B; ++b) {
            const uint2 nxt = ld_stream2(g.bases2, wbase + 2 * (b + 1), nwords);
            const uint32_t w0 = cur.x, w1 = cur.y, w2 = nxt.x;
            const long long Sb = S + 32ll * b;
            const bool first = (b == 0);
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const uint32_t E = kmer_addr(t, w0, w1, w2);
                uint32_t v[NW];
                lookup<NW>(smem, E, o16, o8, o4, v);
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const int si = (t - 4 * gi) & 31;
                    if (MODE == 0) {
                        sa[si] = (gi == 0) ? v[0] : sa[si] + v[gi];
                    } else {
                        sa[si] = (gi == 0) ? v[0] : sa[si] + v[2 * gi];
                        sb[si] = (gi == 0) ? v[1] : sb[si] + v[2 * gi + 1];
                    }
                }
                const int sc = (t - D) & 31;
                if (t < D) {
                    if (first) { outa[t] = sa[sc]; outb[t] = sb[sc]; }
                    else complete(sa[sc], sb[sc], Sb + t - D);
                } else {
                    complete(sa[sc], sb[sc], Sb + t - D);
                }
            }
            cur = nxt;
        }
        // windows straddling into the next lane: fetch the partial sums it accumulated
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const int si = (32 - D + j) & 31;
            const uint32_t ia = __shfl_down_sync(0xffffffffu, outa[j], 1);
            uint32_t ib = 0u;
            if (MODE == 1) ib = __shfl_down_sync(0xffffffffu, outb[j], 1);
            complete(sa[si] + ia, sb[si] + ib, S + L - D + j);
        }
    }

    if (MODE == 1) {
        __syncthreads();
        if (sp.d_hist)
            for (int i = threadIdx.x; i < sp.n_bins; i += blockDim.x)
                if (hist[i]) atomicAdd(sp.d_hist + i, (unsigned long long)hist[i]);
        double c = (double)cnt;
        block_reduce3(s1, s2, c, red_scratch);
        if (threadIdx.x == 0) {
            double* p = work.partials + 3ll * blockIdx.x;
            p[0] = s1; p[1] = s2; p[2] = c;
        }
    }
}

template <int G, int MODE>
static cudaError_t launch_one(const uint32_t* gtab, GenomeView g, SeqTable seqs, FastWork w, int m, int L,
                              int grid, StatsParams sp, cudaStream_t st) {
    constexpr int NW = MODE == 0 ? G : 2 * G;
    const int threads = FastCfg<G, MODE>::THREADS;
    const size_t smem = TableLayout<NW>::BYTES + (MODE == 1 ? (size_t)sp.n_bins * 4 : 0);
    static bool configured[2] = {false, false};   // per template instance
    auto kern = fast_scan_kernel<G, MODE>;
    if (!configured[0]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TableLayout<NW>::BYTES + CGBK_MAX_HIST_BINS * 4);
        if (e != cudaSuccess) return e;
        configured[0] = true;
    }
    kern<<<grid, threads, smem, st>>>(gtab, g, seqs, w, m, L / 32, sp);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t dispatch(int G, const uint32_t* gtab, GenomeView g, SeqTable seqs, FastWork w, int m,
                            int L, int grid, StatsParams sp, cudaStream_t st) {
    switch (G) {
        case 1: return launch_one<1, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 2: return launch_one<2, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 3: return launch_one<3, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 4: return launch_one<4, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 5: return launch_one<5, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 6: return launch_one<6, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 7: return launch_one<7, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        case 8: return launch_one<8, MODE>(gtab, g, seqs, w, m, L, grid, sp, st);
        default: return cudaErrorInvalidValue;
    }
}

int fast_threads(int G, int mode) {
    if (mode == 0) return 512;
    return G <= 6 ? 384 : 256;
}

cudaError_t launch_fast_filter(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, int L,
                               int grid, cudaStream_t st) {
    StatsParams sp = {};
    return dispatch<0>(mdl->G, mdl->d_coarse, g, seqs, w, mdl->m, L, grid, sp, st);
}

cudaError_t launch_fast_stats(const cgbk_model* mdl, GenomeView g, SeqTable seqs, FastWork w, int L,
                              int grid, double hist_lo, double inv_binw, int n_bins,
                              unsigned long long* d_hist, cudaStream_t st) {
    StatsParams sp;
    sp.inv_scale_d = ldexp(1.0, -mdl->kexp);
    sp.inv_scale_f = (float)sp.inv_scale_d;
    sp.hist_lo = (float)hist_lo;
    sp.inv_binw = (float)inv_binw;
    sp.n_bins = n_bins;
    sp.d_hist = d_hist;
    return dispatch<1>(mdl->G, mdl->d_precise, g, seqs, w, mdl->m, L, grid, sp, st);
}
