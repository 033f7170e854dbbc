// Tiled PSWM scan kernel for sm_100a (the throughput path), as a template that the three
// translation units cgbk_fast_filter.cu / cgbk_fast_stats.cu / cgbk_fast_moments.cu instantiate
// (one per MODE, so that they compile side by side).
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
//    of a warp never bank-conflict: every chunk costs exactly one 128-byte crossbar
//    pass per 8/16/32 lanes, the LSU floor;
//  * windows that straddle two lanes: the first D = 4(G-1) positions of a lane also
//    accumulate the tail of the previous lane's windows; those partial sums are
//    parked in a per-warp shared-memory strip and picked up by the previous lane
//    when its own segment is done (lane 31's tail belongs to the next tile, which
//    starts 32 positions early);
//  * the unrolled body is branch-free (window validity is a select, not a branch) so
//    that ptxas can hoist the next position's LDS above the current window's epilogue;
//  * MODE 0 ("filter"): both strands ride in one 32-bit word as two biased 16-bit
//    surpluses; a window whose score could reach the threshold has bit 15 / bit 31
//    set, so the per-window test is one LOP3.  Candidates go to cgbk_exact.cu for
//    float64 re-scoring in the reference's order (bit-exact hit lists);
//  * MODE 1 ("stats"): int32 fixed point (2^-kexp) per strand, soft-max
//    log2(2^f + 2^r) = max + log2(1 + 2^-|f-r|) on the MUFU unit, per-block
//    shared-memory histogram, and integer moments (sum d, sum d^2 of the score
//    re-centred and rounded to the finest unit the accumulators allow: order
//    independent, hence deterministic).  Optionally every window's fixed-point soft-max
//    score is written to a "score plane" laid out [tile][row][lane] so that each position
//    is one coalesced 128-byte store; cgbk_posteriors reads it back instead of re-scoring
//    the promoters;
//  * MODE 2 ("moments"): MODE 1 without the histogram (the reference's estimator only ever
//    uses mean and std, binding_model.py:85-87);
//  * MODE 3 / 4 ("lean" stats / moments): MODE 1 / 2 for callers that keep no score plane and
//    whose histogram range covers every possible score (the batched scan's default): no plane
//    store and no clamping of the bin index in the per-window epilogue -- a score that rounding
//    pushes a hair outside the range lands in a guard row on either side of the shared-memory
//    histogram, folded into the first / last bin at the flush (3 of ~38 instructions per window);
//  * MODE 1 / 2 with a threshold ("fused" scan): the fixed-point strand scores are exact to a
//    few units of 2^-kexp, so the same pass also finds the site-search candidates
//    (genome.py:433-445).  Per window TWO instructions: the sign of (threshold - 1 - max(f, r))
//    is shifted into a flag register (no branch, no atomic, the same cost whether one window
//    in a million or one in twenty is a candidate -- low-information motifs reach the latter);
//    the register goes to a "flag plane" [tile][row][lane] once per 32-position body, one
//    coalesced 128-byte store like the score plane's, and the tile's flag count to a per-tile
//    array.  cgbk_exact.cu walks the plane tile by tile, re-scores the flagged windows exactly
//    and writes the sites out IN ORDER (prefix sums over the tile counts: no sort).  One pass
//    over the sequence then yields the background record AND the hit candidates;
//  * VAR 2 ("peeled"): the first 32-position body of a lane segment is its own code copy
//    without the epilogues of the D windows that belong to the previous lane (they are
//    only parked), the following bodies a second copy without the parking: no work is
//    spent on windows that are not the lane's own.  VAR 0 keeps one body for both
//    (half the code: better when a warp runs a single small tile);
//  * tiles that contain a byte outside ACGTacgt are not scored on the table path: the
//    filter lists them for the exact kernels, the stats modes score them in place with the
//    reference's float64 arithmetic (ambiguity repair, pssm_model.py:88-101).
#pragma once

#include <mutex>
#include <type_traits>

#include "cgbk_device.cuh"

// Threads per CTA (one CTA per SM).  Measured on B200 at 1 Gbp / 5 Mbp, LexA m = 22, stats kernel:
// 256: 4.47e11 / 26.3 us, 320: 4.76e11 / 24.3, 384: 4.92e11 / 25.5, 448: 4.72e11 / 27.4,
// 512: 4.76e11 / 26.3 (128 registers, spills), 640: 4.08e11.
#ifndef CGBK_STATS_THREADS
#define CGBK_STATS_THREADS 384
#endif
#ifndef CGBK_STATS_THREADS_SMALLG      // G <= 2 (motifs of up to 8 bases): fewer registers, more warps
#define CGBK_STATS_THREADS_SMALLG 384
#endif
// per group count (tuning builds: -DCGBK_STATS_THREADS_G3=384 ...).  Up to G = 5 the kernel fits 128
// registers (no or a few bytes of spills): 512 threads measured 5-6 % faster than 384 on B200
// (250 Mbp, fused scan: G3 0.418 -> 0.394 ms, G4 0.420 -> 0.398, G5 0.474 -> 0.449); G = 6 needs
// its 168 registers (448 threads: 0.503 -> 0.550 ms)
#ifndef CGBK_STATS_THREADS_G1
#define CGBK_STATS_THREADS_G1 512
#endif
#ifndef CGBK_STATS_THREADS_G2
#define CGBK_STATS_THREADS_G2 512
#endif
#ifndef CGBK_STATS_THREADS_G3
#define CGBK_STATS_THREADS_G3 512
#endif
#ifndef CGBK_STATS_THREADS_G4
#define CGBK_STATS_THREADS_G4 512
#endif
#ifndef CGBK_STATS_THREADS_G5
#define CGBK_STATS_THREADS_G5 512
#endif
#ifndef CGBK_STATS_THREADS_G6
#define CGBK_STATS_THREADS_G6 CGBK_STATS_THREADS
#endif
#ifndef CGBK_STATS_THREADS_G7
#define CGBK_STATS_THREADS_G7 256
#endif
#ifndef CGBK_STATS_THREADS_G8
#define CGBK_STATS_THREADS_G8 256
#endif
#ifndef CGBK_STATS_EW
#define CGBK_STATS_EW 4
#endif
#ifndef CGBK_STATS_BATCH
#define CGBK_STATS_BATCH 1
#endif
#ifndef CGBK_FILTER_BATCH
#define CGBK_FILTER_BATCH 1
#endif
#ifndef CGBK_FILTER_THREADS
#define CGBK_FILTER_THREADS 640
#endif

template <int NW>
struct TableLayout {
    static constexpr int N16 = NW / 4;
    static constexpr int HAS8 = (NW % 4) >= 2 ? 1 : 0;
    static constexpr int HAS4 = NW % 2;
    static constexpr int NCH = N16 + HAS8 + HAS4;
    static constexpr int BYTES = NCH * 32768;
};

// Global tables arrive chunk-major (see chunked_layout() in cgbk_api.cu): all 16-byte
// chunks [c][e][4 words], then the 8-byte chunk [e][2], then the 4-byte chunk [e].
// Eight consecutive lanes write the eight 16-byte replicas of one chunk entry (one
// conflict-free 128-byte row); the global read of the entry is a broadcast.
template <int NW, int THREADS>
__device__ __forceinline__ void fill_tables(unsigned char* smem, const uint32_t* __restrict__ gtab) {
    using TL = TableLayout<NW>;
    constexpr int TOTAL = TL::NCH * 256 * 8;
    constexpr int ITER = (TOTAL + THREADS - 1) / THREADS;
    // all global loads first (they are independent), then the stores: one L2 round trip
    // instead of ITER serialized ones
    uint4 v[ITER];
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int i = threadIdx.x + k * THREADS;
        const int ce = i >> 3;
        const int c = ce >> 8, e = ce & 255;
        v[k] = make_uint4(0u, 0u, 0u, 0u);
        if (i < TOTAL) {
            if (c < TL::N16) {
                v[k] = __ldg(reinterpret_cast<const uint4*>(gtab) + c * 256 + e);
            } else if (TL::HAS8 && c == TL::N16) {
                const uint2 q = __ldg(reinterpret_cast<const uint2*>(gtab + TL::N16 * 1024) + e);
                v[k] = make_uint4(q.x, q.y, q.x, q.y);
            } else {
                const uint32_t q = __ldg(gtab + TL::N16 * 1024 + TL::HAS8 * 512 + e);
                v[k] = make_uint4(q, q, q, q);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int i = threadIdx.x + k * THREADS;
        if (i < TOTAL) *reinterpret_cast<uint4*>(smem + (i >> 3) * 128 + (i & 7) * 16) = v[k];
    }
}

// (x & 0x7f80) | o in ONE LOP3 (ptxas otherwise keeps the masked index as a common subexpression
// and spends a second instruction on every replica offset)
__device__ __forceinline__ uint32_t entry_addr(uint32_t x, uint32_t o) {
    uint32_t a;
    asm("lop3.b32 %0, %1, 0x7f80, %2, 0xEA;" : "=r"(a) : "r"(x), "r"(o));
    return a;
}

// all NW words of the table entry of stream word x (4-mer index at bits 7..14, anything around it);
// o16/o8/o4 = this lane's replica offsets
template <int NW>
__device__ __forceinline__ void lookup(const unsigned char* sb, uint32_t x, uint32_t o16, uint32_t o8,
                                       uint32_t o4, uint32_t (&v)[NW]) {
    using TL = TableLayout<NW>;
    if (TL::N16 > 0) {
        const uint32_t a = entry_addr(x, o16);
#pragma unroll
        for (int c = 0; c < TL::N16; ++c) {
            const uint4 q = *reinterpret_cast<const uint4*>(sb + c * 32768 + a);
            v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
        }
    }
    if (TL::HAS8) {
        const uint2 q = *reinterpret_cast<const uint2*>(sb + TL::N16 * 32768 + entry_addr(x, o8));
        v[4 * TL::N16] = q.x; v[4 * TL::N16 + 1] = q.y;
    }
    if (TL::HAS4) {
        v[NW - 1] = *reinterpret_cast<const uint32_t*>(sb + (TL::N16 + TL::HAS8) * 32768 + entry_addr(x, o4));
    }
}

// the stream shifted so that the 4-mer at body position t sits at bits 7..14 (NOT masked: lookup()
// masks and adds the replica offset in one instruction), from the body's two stream words + one
// lookahead word
__device__ __forceinline__ uint32_t kmer_addr(int t, uint32_t w0, uint32_t w1, uint32_t w2) {
    const uint32_t lo = t < 16 ? w0 : w1, hi = t < 16 ? w1 : w2;
    const int sh = t < 16 ? 2 * t : 2 * t - 32;
    if (sh >= 7) return __funnelshift_r(lo, hi, sh - 7);
    return lo << (7 - sh);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint2 ld_stream2(const uint32_t* __restrict__ b2, long long wi, long long nw) {
    if (wi + 1 < nw) return __ldg(reinterpret_cast<const uint2*>(b2 + wi));
    return make_uint2(wi < nw ? __ldg(b2 + wi) : 0u, 0u);
}

// Candidates are first parked in a small per-warp shared-memory buffer (one shared-memory
// atomic each) and moved to the global list by flush_candidates() at warp-convergent points
// with ONE global atomic per flush and coalesced stores: with low-information motifs ~1 % of
// the windows are candidates and a global atomic per candidate on a single counter would
// serialise the whole scan.  A full buffer falls back to the direct global append.
#define CGBK_CAND_STAGE 96      // entries per warp

__device__ __forceinline__ void park_candidate(unsigned long long rec, unsigned long long* stage,
                                               unsigned int* stage_cnt, long long* counters,
                                               unsigned long long* cands, long long cand_cap) {
    const unsigned int idx = atomicAdd(stage_cnt, 1u);
    if (idx < CGBK_CAND_STAGE) {
        stage[idx] = rec;
    } else {
        const unsigned long long slot =
            atomicAdd(reinterpret_cast<unsigned long long*>(counters + CGBK_CTR_CANDIDATES), 1ull);
        if ((long long)slot < cand_cap) cands[slot] = rec;
    }
}

// Rare path of the filter kernel, deliberately NOT inlined (it would otherwise be replicated
// at each of the 52 completion sites of the unrolled body and push the hot loop out of the
// instruction cache): examine 4 consecutive windows wt0..wt0+3 of this lane whose flag OR was
// non-zero and append those that are the tile's own to the candidate list.
static __device__ __noinline__ void emit_candidates4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, int wt0,
                                              int lane_first, int own_n, long long gpos0,
                                              unsigned long long* stage, unsigned int* stage_cnt,
                                              long long* counters, unsigned long long* cands,
                                              long long cand_cap) {
    const uint32_t a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int wt = wt0 + k;
        if ((a[k] & 0x80008000u) && wt >= lane_first && wt < own_n) {
            const unsigned long long rec = (unsigned long long)(gpos0 + wt) |
                                           ((unsigned long long)(a[k] >> 31) << 63) |
                                           ((unsigned long long)((a[k] >> 15) & 1u) << 62);
            park_candidate(rec, stage, stage_cnt, counters, cands, cand_cap);
        }
    }
}

// warp-convergent: move this warp's parked candidates to the global list
static __device__ __noinline__ void flush_candidates(unsigned long long* stage, unsigned int* stage_cnt, int lane,
                                                 long long* counters, unsigned long long* cands,
                                                 long long cand_cap) {
    __syncwarp();
    unsigned int n = *stage_cnt;
    if (n == 0) return;                                     // warp-uniform
    if (n > CGBK_CAND_STAGE) n = CGBK_CAND_STAGE;
    unsigned long long base = 0;
    if (lane == 0)
        base = atomicAdd(reinterpret_cast<unsigned long long*>(counters + CGBK_CTR_CANDIDATES),
                         (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (unsigned int i = lane; i < n; i += 32)
        if ((long long)(base + i) < cand_cap) cands[base + i] = stage[i];
    __syncwarp();
    if (lane == 0) *stage_cnt = 0u;
    __syncwarp();
}

// Stats scan, rare path: a tile that holds a byte outside ACGT is scored by the warp that drew
// it with the reference's exact arithmetic (PSSMModel._calculate repair included), straight from
// the device ExactTables (L1 resident, 5 KB).  Deliberately NOT inlined: the hot loop keeps its
// registers.  Moments go to the warp's three doubles in shared memory, the histogram to global.
// In a fused scan the windows that may be sites are flagged in the tile's rows of the flag plane
// like the table path does (conservatively: either strand within 1e-9 of the threshold; the
// re-score pass decides), and the tile's flag count is returned.
struct ExactFlagSink {
    unsigned int* frows;            // this tile's (B + 1) x 32 flag words (NULL: no hit search)
    double thr;
};

static __device__ __noinline__ int exact_tile_stats(const ExactTables* tab, const uint32_t* bases2, const uint32_t* amb,
                                              const uint32_t* lower, long long cap_bases, long long gbase,
                                              long long w_lo, long long w_hi, int m, long long tile, int L, int D,
                                              unsigned long long* d_hist, int n_bins, double hist_lo,
                                              double inv_binw, int* plane, double scale, double centre, double inv_q,
                                              unsigned long long* acc4, const ExactFlagSink* sink) {
    GenomeView g;
    g.bases2 = bases2; g.amb = amb; g.lower = lower; g.cap_bases = cap_bases;
    // moments in the scan's own integer units (score = centre + d * q, d rounded to nearest): exact
    // integer sums, so the totals do not depend on which warp or block scored the tile
    long long sd = 0, cnt = 0;
    unsigned long long sd2 = 0;               // |d| < 2^27 and at most 128 windows per lane: below 2^61
    const int lane = threadIdx.x & 31, B = L >> 5;
    int n_flag = 0;
    if (sink->frows) {
        for (int r = 0; r <= B; ++r) sink->frows[r * 32 + lane] = 0u;
        __syncwarp();
    }
    for (long long x = w_lo + lane; x < w_hi; x += 32) {
        double f, r;
        exact_window(g, tab, m, gbase + x, false, f, r);
        const double s = softmax2_f64((double)(float)f, (double)(float)r);
        const long long d = llrint((s - centre) * inv_q);
        sd += d; sd2 += (unsigned long long)(d * d); cnt += 1;
        if (d_hist) {
            int bin = (int)floor((s - hist_lo) * inv_binw);
            bin = bin < 0 ? 0 : (bin >= n_bins ? n_bins - 1 : bin);
            atomicAdd(d_hist + bin, 1ull);
        }
        if (plane) plane[plane_index(tile, (int)(x - w_lo), L, D)] = (int)llrint(s * scale);
        if (sink->frows && ((double)(float)f >= sink->thr - 1e-9 || (double)(float)r >= sink->thr - 1e-9)) {
            // the flag plane's addressing: row b < B holds bit 31 - t of the window completed at body
            // position t (within + D = 32 b + t), the tail row bit D - 1 - j
            const int wt = (int)(x - w_lo), ln = wt / L, within = wt - ln * L;
            int row, bit;
            if (within < L - D) { const int c = within + D; row = c >> 5; bit = 31 - (c & 31); }
            else { row = B; bit = D - 1 - (within - (L - D)); }
            atomicOr(sink->frows + row * 32 + ln, 1u << bit);
            ++n_flag;
        }
    }
    unsigned long long hi2 = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sd += __shfl_down_sync(0xffffffffu, sd, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        const unsigned long long lo_o = __shfl_down_sync(0xffffffffu, sd2, o), hi_o = __shfl_down_sync(0xffffffffu, hi2, o);
        sd2 += lo_o; hi2 += hi_o + (sd2 < lo_o ? 1ull : 0ull);
        n_flag += __shfl_down_sync(0xffffffffu, n_flag, o);
    }
    if (lane == 0) {
        // this warp's own slot: (n, sum d, sum d^2 low, sum d^2 high)
        acc4[0] += (unsigned long long)cnt; acc4[1] += (unsigned long long)sd;
        const unsigned long long before = acc4[2];
        acc4[2] = before + sd2; acc4[3] += hi2 + (acc4[2] < before ? 1ull : 0ull);
    }
    return n_flag;          // valid in lane 0
}

// block-wide sum of (n, sum d, 128-bit sum d^2) in integers: exact, hence the same whatever the
// order; result valid in thread 0.  scratch: 4 * 32 words of 8 bytes
__device__ __forceinline__ void block_reduce_moments(unsigned long long& n, unsigned long long& sd,
                                                     unsigned long long& lo, unsigned long long& hi,
                                                     unsigned long long* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    auto warp_sum = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n += __shfl_down_sync(0xffffffffu, n, o);
            sd += __shfl_down_sync(0xffffffffu, sd, o);
            const unsigned long long lo_o = __shfl_down_sync(0xffffffffu, lo, o), hi_o = __shfl_down_sync(0xffffffffu, hi, o);
            lo += lo_o; hi += hi_o + (lo < lo_o ? 1ull : 0ull);
        }
    };
    warp_sum();
    if (lane == 0) { scratch[warp] = n; scratch[32 + warp] = sd; scratch[64 + warp] = lo; scratch[96 + warp] = hi; }
    __syncthreads();
    if (warp == 0) {
        n = lane < nw ? scratch[lane] : 0ull; sd = lane < nw ? scratch[32 + lane] : 0ull;
        lo = lane < nw ? scratch[64 + lane] : 0ull; hi = lane < nw ? scratch[96 + lane] : 0ull;
        warp_sum();
    }
    __syncthreads();
}

// raw moments (sum s, sum s^2) of n scores s_i = centre + d_i * q from the exact integer sums
__device__ __forceinline__ void moments_from_integers(unsigned long long n, unsigned long long sd, unsigned long long lo,
                                                      unsigned long long hi, double centre, double q, double& cn,
                                                      double& s1, double& s2) {
    const double a = (double)(long long)sd, b2 = (double)hi * 18446744073709551616.0 + (double)lo;
    cn = (double)n;
    s1 = cn * centre + a * q;
    s2 = cn * centre * centre + 2.0 * centre * q * a + q * q * b2;
}

__host__ __device__ constexpr int fast_threads_of(int G, int MODE) {
    return MODE == 0 ? CGBK_FILTER_THREADS
         : G == 1 ? CGBK_STATS_THREADS_G1 : G == 2 ? CGBK_STATS_THREADS_G2 : G == 3 ? CGBK_STATS_THREADS_G3
         : G == 4 ? CGBK_STATS_THREADS_G4 : G == 5 ? CGBK_STATS_THREADS_G5 : G == 6 ? CGBK_STATS_THREADS_G6
         : G == 7 ? CGBK_STATS_THREADS_G7 : CGBK_STATS_THREADS_G8;
}

template <int G, int MODE>
struct FastCfg {
    static constexpr int THREADS = fast_threads_of(G, MODE);
    static constexpr int NW = MODE == 0 ? G : 2 * G;
    static constexpr int D = 4 * (G - 1);
    static constexpr int STASH_BYTES = ((THREADS / 32) * D * 33 * (MODE == 0 ? 4 : 8) + 15) / 16 * 16;
    // MODE 0: per-warp candidate staging [CGBK_CAND_STAGE] u64 + one counter word (padded to 16 B)
    static constexpr int CAND_BYTES = MODE == 0 ? (THREADS / 32) * (CGBK_CAND_STAGE * 8 + 16) : 0;
};

#define CGBK_VAR_UNIFIED 0   // one 32-position body for every part of a lane segment
#define CGBK_VAR_ONE 1       // lane segments of exactly 32 positions (filter only)
#define CGBK_VAR_PEEL 2      // first body and following bodies as separate code copies (stats modes)

template <int G, int MODE, int VAR>
__global__ void __launch_bounds__((FastCfg<G, MODE>::THREADS), 1)
fast_scan_kernel(const uint32_t* __restrict__ gtab, GenomeView g, SeqTable seqs, FastWork work, int m,
                 int B, StatsParams sp) {
    using Cfg = FastCfg<G, MODE>;
    constexpr int NW = Cfg::NW;
    constexpr int D = Cfg::D;
    constexpr bool ONE = VAR == CGBK_VAR_ONE;
    constexpr bool PEEL = VAR == CGBK_VAR_PEEL;
    constexpr bool HIST = MODE == 1 || MODE == 3;
    constexpr bool LEAN = MODE >= 3;       // no score plane, histogram guard rows instead of clamps
    constexpr bool GROUP_VALID = PEEL && MODE != 0;
    // completed windows are handled in groups: 4 for the filter's deferred flag test, CGBK_STATS_EW
    // for the stats epilogue (its dependent chains run side by side)
    constexpr int EW = MODE == 0 ? 4 : CGBK_STATS_EW;
    static_assert(MODE == 0 || CGBK_STATS_EW == 4, "the fused candidate test examines 4 windows at a time");
    // stream positions whose table look-ups are issued together before their accumulation
    constexpr int SB = MODE == 0 ? CGBK_FILTER_BATCH : CGBK_STATS_BATCH;
    using TL = TableLayout<NW>;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* stash_base = smem + TL::BYTES;
    // this warp's candidate staging buffer and its fill count (found again on demand: no pointer
    // is kept live through the hot loop), then (MODE 1) the histogram
    auto cand_ptr = [&]() -> unsigned char* {
        return smem + TL::BYTES + Cfg::STASH_BYTES + (threadIdx.x >> 5) * (CGBK_CAND_STAGE * 8 + 16);
    };
    unsigned int* hist_rows = reinterpret_cast<unsigned int*>(smem + TL::BYTES + Cfg::STASH_BYTES + Cfg::CAND_BYTES);
    __shared__ unsigned long long red_scratch[128];
    // stats modes: per warp, integer moments (n, sum d, sum d^2 as 128 bits) of the exactly scored tiles
    __shared__ unsigned long long exact_acc[32 * 4];
    __shared__ int is_last_block;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t o16 = (lane & 7) << 4, o8 = (lane & 15) << 3, o4 = lane << 2;
    if (ONE) B = 1;
    const int L = 32 * B;
    const int stride = 32 * L - 32;
    const long long nwords = g.cap_bases >> 4;
    const int S = lane * L;                                   // tile-local start of this lane's segment

    // tile -> (packed offset of its sequence, sequence length, first tile of the sequence)
    auto geometry = [&](long long tile, long long& gbase, long long& seq_len, long long& tile0) {
        if (seqs.n_seq <= CGBK_INLINE_SEQS) {
            int sq = 0;
#pragma unroll
            for (int i = 1; i < CGBK_INLINE_SEQS; ++i)
                if (i < seqs.n_seq && seqs.prefix_v[i] <= tile) sq = i;
            gbase = seqs.off_v[sq]; seq_len = seqs.len_v[sq]; tile0 = seqs.prefix_v[sq];
        } else {
            const int sq = find_seq(seqs, tile);
            gbase = __ldg(seqs.off + sq); seq_len = __ldg(seqs.len + sq); tile0 = __ldg(seqs.tile_prefix + sq);
        }
    };
    // The global loads that open a tile (ambiguity words + first stream words) are issued one
    // tile ahead: the first tile's before the table fill, the next tile's before the current
    // tile's tail section, so their DRAM latency never sits on a warp's critical path.
    uint32_t pre_amb = 0u;
    uint2 pre_cur = make_uint2(0u, 0u);
    auto preload = [&](long long tile) {
        if (tile >= work.tile_end) return;
        long long gbase, seq_len, tile0;
        geometry(tile, gbase, seq_len, tile0);
        const long long p0 = gbase + (tile - tile0) * stride + S;
        const long long aw = p0 >> 5, left = (g.cap_bases >> 5) - aw;
        const int n_ok = left < (long long)B ? (left < 0 ? 0 : (int)left) : B;     // B <= 4
        uint32_t a = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b < n_ok) a |= __ldg(g.amb + aw + b);
        pre_amb = a;
        pre_cur = ld_stream2(g.bases2, p0 >> 4, nwords);
    };
    // this launch covers tiles [work.tile_begin, work.tile_end); every warp's first tile is static
    long long tile = work.tile_begin + (long long)blockIdx.x * wpb + warp;
    preload(tile);

    fill_tables<NW, Cfg::THREADS>(smem, gtab);
    // LEAN: one guard row before bin 0 and one after the last bin
    unsigned int* hist = hist_rows + (LEAN ? sp.hist_words : 0);
    if (HIST)
        for (int i = threadIdx.x; i < (sp.n_bins + (LEAN ? 2 : 0)) * sp.hist_words; i += blockDim.x) hist_rows[i] = 0u;
    if (MODE == 0 && (threadIdx.x & 31) == 0) *reinterpret_cast<unsigned int*>(cand_ptr() + CGBK_CAND_STAGE * 8) = 0u;
    if (MODE != 0 && threadIdx.x < 128) exact_acc[threadIdx.x] = 0ull;
    __syncthreads();

    // per-warp strip for the partial sums handed to the previous lane: [D][33]
    uint32_t* stash32 = reinterpret_cast<uint32_t*>(stash_base) + warp * (D * 33) + lane;
    uint2* stash64 = reinterpret_cast<uint2*>(stash_base) + warp * (D * 33) + lane;
    const int last_bin = sp.n_bins - 1;

    int filler = 0;               // completions counted into the centre bin that are no real windows
    long long sum_d = 0;          // exact integer sum of re-centred scores (moment units)
    unsigned long long sum_d2 = 0ull, sum_d2_hi = 0ull;   // sum of their squares, 128 bits
    long long cnt = 0;

    // histogram row of bin b starts at b * hist_words words; with 32 words per row every lane
    // counts into its own bank (conflict free), with 1 word per row lanes share the counters
    const unsigned hist_saddr = (unsigned)__cvta_generic_to_shared(hist) + (sp.hist_words == 32 ? 4u * lane : 0u);
    const unsigned hist_row_bytes = 4u * (unsigned)sp.hist_words;

    // Tiles are handed out dynamically (one global atomic per warp tile, fetched one tile
    // ahead): with only a few tiles per warp a static split leaves most warps idle while the
    // unlucky ones run their extra tile.  All outputs are order independent.
    unsigned long long* tile_ctr = reinterpret_cast<unsigned long long*>(work.counters + work.ctr_slot);
    const long long first_dyn = work.tile_begin + (long long)gridDim.x * wpb;
    long long prefetched = 0;                                 // lane 0: id of the tile after the next one
    if (lane == 0) prefetched = first_dyn + (long long)atomicAdd(tile_ctr, 1ull);
    auto advance = [&]() -> long long {
        const long long t = __shfl_sync(0xffffffffu, prefetched, 0);   // fetched one tile ago
        if (lane == 0) prefetched = first_dyn + (long long)atomicAdd(tile_ctr, 1ull);
        return t;
    };
    auto flush_staged = [&]() {
        unsigned char* cb = cand_ptr();
        flush_candidates(reinterpret_cast<unsigned long long*>(cb),
                         reinterpret_cast<unsigned int*>(cb + CGBK_CAND_STAGE * 8), lane, work.counters,
                         work.cands, work.cand_cap);
    };
    while (tile < work.tile_end) {
        long long gbase, seq_len, tile0;
        geometry(tile, gbase, seq_len, tile0);
        const long long A = (tile - tile0) * stride;
        const long long rem = seq_len - m + 1 - A;            // valid windows from A on
        const int own_n = rem < stride ? (int)rem : stride;   // this tile owns windows [A, A + own_n)
        // GROUP_VALID (the peeled code of large inputs): the histogram / moment epilogue tests validity
        // once per group of EW windows, so the table path counts the tile's windows up to a multiple of
        // EW; the up to EW - 1 windows beyond it (only ever at the end of a sequence: a full tile's count
        // is a multiple of EW) are scored exactly after the tile.  The candidate flags keep the exact
        // count own_n.  The unified code (a warp runs a single small tile, launch-latency bound) tests
        // every window instead and never takes the exact detour.
        const int own_v = GROUP_VALID ? (own_n & ~(EW - 1)) : own_n;
        const long long wbase = (gbase + A + S) >> 4;

        uint2 cur = pre_cur;
        if (__any_sync(0xffffffffu, pre_amb != 0)) {
            // a tile touching a non-ACGT byte: exact arithmetic (filter: listed for the exact
            // hit kernel; stats: scored right here, out of line)
            if (MODE == 0) {
                if (lane == 0) {
                    const unsigned long long k = atomicAdd(
                        reinterpret_cast<unsigned long long*>(work.counters + CGBK_CTR_DIRTY_TILES), 1ull);
                    work.dirty_tiles[k] = (unsigned int)tile;
                }
            } else {
                if (lane == 0)
                    atomicAdd(reinterpret_cast<unsigned long long*>(work.counters + CGBK_CTR_DIRTY_TILES), 1ull);
                ExactFlagSink sink;
                sink.frows = work.fplane ? work.fplane + tile * (long long)((B + 1) * 32) : nullptr;
                sink.thr = sp.thr;
                const int nf = exact_tile_stats(static_cast<const ExactTables*>(sp.exact_tab), g.bases2, g.amb, g.lower,
                                                g.cap_bases, gbase, A, A + (own_n > 0 ? own_n : 0), m, tile, L, D,
                                                sp.d_hist, sp.n_bins, sp.hist_lo_d, sp.inv_binw_d, LEAN ? nullptr : sp.plane,
                                                (double)sp.scale, sp.c, sp.inv_q, exact_acc + 4 * warp, &sink);
                if (work.fplane && lane == 0) work.tile_cand[tile] = nf;
            }
            tile = advance();
            preload(tile);
            continue;
        }

        uint32_t sa[32], sb[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { sa[i] = 0u; sb[i] = 0u; }
        int acc_d = 0;                 // per-body integer sum (stats modes)
        long long acc_d2 = 0;          // per-tile integer sum of squares (stats modes)
        // score plane rows of this tile: [(B + 1) * 32][32] ints, row = body*32 + t (last: the tail)
        int* prow = (MODE != 0 && !LEAN && sp.plane) ? sp.plane + tile * (long long)((B + 1) * 1024) + lane : nullptr;

        // fused scan: candidate flags of the current body's completed windows (shifted in one by
        // one, newest at bit 0) and this lane's column of the tile's flag rows
        uint32_t cflags = 0u;
        int n_flag = 0;            // flagged windows of this lane in this tile
        unsigned int* fprow = (MODE != 0 && work.fplane) ? work.fplane + tile * (long long)((B + 1) * 32) + lane : nullptr;
        // bits [0, n) of a word, n in 0..32
        auto lowmask = [](int n) -> uint32_t { return (uint32_t)((1ull << n) - 1ull); };
        uint32_t flag_or = 0u, pend[EW], pendb[EW];            // the last (up to) EW completed windows
#pragma unroll
        for (int i = 0; i < EW; ++i) { pend[i] = 0u; pendb[i] = 0u; }
        bool staged = false;                                   // this lane parked candidates
        // one completed window: accumulators (a, b); its tile-local index is wbase_t + wrel with
        // wrel a compile-time constant, and it is one of the tile's own iff wrel < lim (lim =
        // own_n - wbase_t, one subtraction per body); plane row; `live` is false for the partial
        // sums of the previous lane's windows (unified body only)
        auto complete = [&](uint32_t a, uint32_t b, int wbase_t, int wrel, int lim, bool live, int row,
                            int q /* window % EW */) {
            const int wt = wbase_t + wrel;
            if (MODE == 0) {
                // Filter: only OR the two flag bits here (one LOP3).  Every 4th window the OR is
                // tested and, in the rare case it is set, the four windows are examined out of
                // line (emit_candidates4).  Partial sums of windows that are not this lane's
                // (first body, t < D) can never carry a flag bit: they lack group 0's offset.
                flag_or |= a & 0x80008000u;
                pend[q] = a;
                if (q == 3) {
                    if (flag_or) {
                        unsigned char* cb = cand_ptr();
                        emit_candidates4(pend[0], pend[1], pend[2], pend[3], wt - 3, S, own_n, (wbase << 4) - S,
                                         reinterpret_cast<unsigned long long*>(cb),
                                         reinterpret_cast<unsigned int*>(cb + CGBK_CAND_STAGE * 8),
                                         work.counters, work.cands, work.cand_cap);
                        staged = true;      // something may be parked: flush at the next convergent point
                    }
                    flag_or = 0u;
                }
            } else {
                // Stats: the soft-max / histogram / moment epilogue is one long dependent chain
                // (max, I2F, EX2, LG2, F2I, mul-hi, RED: ~170 cycles).  Four windows are
                // therefore collected and their chains run side by side, stage by stage.
                pend[q] = a;
                pendb[q] = b;
                if (q == EW - 1) {
                    int mx[EW], sfix[EW];
                    float u[EW];
#pragma unroll
                    for (int k = 0; k < EW; ++k) {
                        const int fq = (int)pend[k], rq = (int)pendb[k];
                        mx[k] = max(fq, rq);
#ifdef CGBK_MINMAX_DIFF
                        u[k] = (float)(mx[k] - min(fq, rq)) * sp.neg_inv_scale;
#else
                        // |f - r| as the absolute value of the converted difference: the |.| is a source
                        // modifier of the multiply (|f|, |r| < 2^30: the difference fits an int32)
                        u[k] = fabsf((float)(fq - rq)) * sp.neg_inv_scale;
#endif
                        // candidate <=> max(f, r) >= thr_fix <=> (thr_fix - 1 - max) < 0: its sign bit
                        // is shifted into the flag register (|both terms| < 2^30: no overflow)
                        cflags = __funnelshift_l((uint32_t)(sp.thr_m1 - mx[k]), cflags, 1);
                    }
#pragma unroll
                    for (int k = 0; k < EW; ++k) u[k] = ex2_approx(u[k]);                 // 2^-|f-r|
#pragma unroll
                    for (int k = 0; k < EW; ++k) u[k] = lg2_approx(1.0f + u[k]);
#pragma unroll
#ifdef CGBK_F2I_CORR
                    for (int k = 0; k < EW; ++k) sfix[k] = mx[k] + __float2int_rn(u[k] * sp.scale);   // fixed point
#else
                    // fixed point WITHOUT a conversion instruction (F2I shares the quarter-rate unit of
                    // EX2 / LG2, the most contended one here): u in [0, 1] plus a power of two `magic`
                    // whose ulp is the fixed-point unit (2^-kexp, at most 2^-23) rounds u to that unit in
                    // the mantissa; the bits above `magic`, scaled to 2^-kexp, are the correction
                    for (int k = 0; k < EW; ++k)
                        sfix[k] = mx[k] + (__float_as_int(u[k] + sp.magic) - sp.magic_bits) * sp.magic_mul;
#endif
#pragma unroll
                    for (int k = 0; k < EW; ++k) {
                        // (the table path's scores are off by -tab_bias; the plane holds true scores)
                        if (!LEAN && prow) prow[(row - (EW - 1) + k) * 32] = sfix[k] + sp.tab_bias;
                        // Validity is ONE select, not a branch: a window that is not the tile's own
                        // takes the moment centre as its score.  It then adds nothing to the
                        // moments and lands in the centre's histogram bin, from which the block's
                        // count of such windows is taken out again before the flush.  (Predicating
                        // the three side effects instead was tried: ptxas wraps a predicated
                        // shared-memory reduction in a convergence region, +4 instructions per window.)
                        // GROUP_VALID: one test per group of EW windows -- groups are aligned to multiples
                        // of EW in the tile and `lim` is too (own_v below), so a group is valid or invalid as
                        // a whole
                        const bool valid = live && (GROUP_VALID ? wrel < lim : wrel - (EW - 1) + k < lim);
                        const int sv = valid ? sfix[k] : sp.half;        // = c_fix - tab_bias
                        if (HIST) {
                            // bin = floor((sv - lo) * n_bins / range) as a signed high multiply (a value a
                            // hair below lo gives -1); LEAN: no clamps, -1 and n_bins are the guard rows
#ifdef CGBK_BIN_MULHI
                            int bin = __mulhi(sv - sp.lo_fix, (int)sp.bin_mul) >> sp.bin_shift;
#else
                            // (sv - lo) * M as ONE 64-bit multiply-add sv * M + (-lo * M): same 64-bit product,
                            // same high word, without the subtraction
                            // (tried: the shift as a high multiply by 2^(32 - shift) -- the multiply-add pipe is the
                            // busier one here, +3 %)
                            int bin = (int)(((long long)sv * (int)sp.bin_mul + sp.bin_c) >> 32) >> sp.bin_shift;
#endif
                            if (!LEAN) bin = min(max(bin, 0), last_bin);
                            // plain shared-memory reduction (atomicAdd(.., 1) would be wrapped in a
                            // warp-collective region by ptxas)
                            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_saddr + hist_row_bytes * (unsigned)bin),
                                         "r"(1u));
                        }
                        const int dq = sv >> sp.sh;       // the tables hold -c_fix + half already
                        acc_d += dq;
                        acc_d2 += (long long)dq * dq;
                    }
                }
            }
        };

        // one 32-position body.  FM 0: `first` known at run time (unified code), FM 1: the first
        // body of the lane segment, FM 2: a later body (peeled code copies).
        auto body = [&](auto fm_tag, int b, uint32_t w0, uint32_t w1, uint32_t w2) {
            constexpr int FM = decltype(fm_tag)::value;
            const int Sb = S + 32 * b;
            const int lim = own_v - Sb;
            const bool first = FM == 0 ? (ONE || b == 0) : (FM == 1);
#pragma unroll
            for (int t0 = 0; t0 < 32; t0 += SB) {
              // look up SB consecutive positions first (independent LDS in flight together) ...
              uint32_t vv[SB][NW];
#pragma unroll
              for (int u = 0; u < SB; ++u) lookup<NW>(smem, kmer_addr(t0 + u, w0, w1, w2), o16, o8, o4, vv[u]);
              // ... then accumulate and complete them in order
#pragma unroll
              for (int u = 0; u < SB; ++u) {
                const int t = t0 + u;
                const uint32_t (&v)[NW] = vv[u];
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
                    // first body: partial sums of the PREVIOUS lane's last D windows -> strip;
                    // later bodies: a window that started in the previous body completes
                    if (FM != 2 && first) {
                        if (MODE == 0) stash32[t * 33] = sa[sc];
                        else stash64[t * 33] = make_uint2(sa[sc], sb[sc]);
                    }
                    if (FM == 2) complete(sa[sc], sb[sc], Sb, t - D, lim, true, b * 32 + t, (t - D) & (EW - 1));
                    else if (FM == 0 && !ONE) complete(sa[sc], sb[sc], Sb, t - D, lim, !first, b * 32 + t, (t - D) & (EW - 1));
                } else {
                    complete(sa[sc], sb[sc], Sb, t - D, lim, true, b * 32 + t, (t - D) & (EW - 1));
                }
              }
            }
        };
        auto after_body = [&](int b) {
            if (MODE != 0) {
                sum_d += acc_d; acc_d = 0;
                if (fprow) {
                    // bit 31 - t <-> the window completed at body position t; keep the lane's own
                    // valid windows: t >= D in the first body (before: the previous lane's tail), and
                    // S + 32 b + t - D < own_n
                    const int t_lo = b == 0 ? D : 0;
                    int t_hi = own_n - (S + 32 * b) + D;
                    t_hi = t_hi < t_lo ? t_lo : (t_hi > 32 ? 32 : t_hi);
                    const uint32_t word = cflags & lowmask(32 - t_lo) & ~lowmask(32 - t_hi);
                    n_flag += __popc(word);
                    fprow[b * 32] = word;              // one coalesced 128-byte row per body
                }
                cflags = 0u;
            }
            if (MODE == 0 && __any_sync(0xffffffffu, staged)) {     // a vote, no memory access, when idle
                flush_staged();
                staged = false;
            }
        };

        if (PEEL) {
            uint2 nxt = ld_stream2(g.bases2, wbase + 2, nwords);
            body(std::integral_constant<int, 1>(), 0, cur.x, cur.y, nxt.x);
            cur = nxt;
            after_body(0);
            for (int b = 1; b < B; ++b) {
                nxt = ld_stream2(g.bases2, wbase + 2 * (b + 1), nwords);
                body(std::integral_constant<int, 2>(), b, cur.x, cur.y, nxt.x);
                cur = nxt;
                after_body(b);
            }
        } else {
            for (int b = 0; b < B; ++b) {
                const uint2 nxt = ld_stream2(g.bases2, wbase + 2 * (b + 1), nwords);
                body(std::integral_constant<int, 0>(), b, cur.x, cur.y, nxt.x);
                cur = nxt;
                after_body(b);
            }
        }
        const long long next_tile = advance();
        preload(next_tile);
        // windows straddling into the next lane: add the partial sums it parked for us
        if (D > 0) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const int si = (32 - D + j) & 31;
                uint32_t ia, ib = 0u;
                if (MODE == 0) {
                    ia = stash32[j * 33 + 1];
                } else {
                    const uint2 q = stash64[j * 33 + 1];
                    ia = q.x; ib = q.y;
                }
                complete(sa[si] + ia, sb[si] + ib, S + L, j - D, own_v - (S + L), true, B * 32 + j, (32 - D + j) & (EW - 1));
            }
            __syncwarp();
            if (MODE == 0 && __any_sync(0xffffffffu, staged)) {
                flush_staged();
                staged = false;
            }
        }
        if (MODE != 0) {
            if (fprow) {
                // the D straddling windows: bit D - 1 - j <-> window S + L - D + j (no bit when D == 0)
                int j_hi = own_n - (S + L - D);
                j_hi = j_hi < 0 ? 0 : (j_hi > D ? D : j_hi);
                const uint32_t word = cflags & lowmask(D) & ~lowmask(D - j_hi);
                n_flag += __popc(word);
                fprow[B * 32] = word;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) n_flag += __shfl_down_sync(0xffffffffu, n_flag, o);
                if (lane == 0) work.tile_cand[tile] = n_flag;
            }
            sum_d += acc_d;
            sum_d2 += (unsigned long long)acc_d2;
            sum_d2_hi += sum_d2 < (unsigned long long)acc_d2 ? 1ull : 0ull;
            const int mine = own_v - S;
            const int mine_c = mine < 0 ? 0 : (mine > L ? L : mine);
            cnt += mine_c;
            // completions of this tile that were not the lane's own windows (they sit in the
            // centre's histogram bin): L per table-scored tile, plus the D junk completions of
            // the unified first body
            filler += L + (PEEL ? 0 : D) - mine_c;
            if (own_v != own_n) {
                // the last windows of the sequence, [own_v, own_n): exact arithmetic, histogram and
                // moments only (their flags are set above, the score plane holds them already)
                ExactFlagSink none;
                none.frows = nullptr;
                none.thr = 0.0;
                exact_tile_stats(static_cast<const ExactTables*>(sp.exact_tab), g.bases2, g.amb, g.lower, g.cap_bases,
                                 gbase, A + own_v, A + own_n, m, tile, L, D, sp.d_hist, sp.n_bins, sp.hist_lo_d,
                                 sp.inv_binw_d, nullptr, (double)sp.scale, sp.c, sp.inv_q, exact_acc + 4 * warp, &none);
            }
        }
        tile = next_tile;
    }

    if (MODE != 0) {
        if (HIST) {
            // the filler completions leave the centre's bin again (every lane's own bank word when
            // the rows are 32 words wide, one shared word otherwise)
            const int cbin = min(max(__mulhi(sp.half - sp.lo_fix, (int)sp.bin_mul) >> sp.bin_shift, 0), last_bin);
            if (filler) atomicSub(hist + cbin * sp.hist_words + (sp.hist_words == 32 ? lane : 0), (unsigned)filler);
        }
        __syncthreads();
        if (HIST && sp.d_hist)
            for (int i = threadIdx.x; i < sp.n_bins; i += blockDim.x) {
                unsigned int v = 0;
                for (int l = 0; l < sp.hist_words; ++l)       // rotate the start: no bank conflict
                    v += hist[i * sp.hist_words + ((l + i) & (sp.hist_words - 1))];
                if (LEAN && (i == 0 || i == last_bin)) {       // the guard rows (both when n_bins == 1)
                    for (int l = 0; l < sp.hist_words; ++l) {
                        if (i == 0) v += hist_rows[l];
                        if (i == last_bin) v += hist[sp.n_bins * sp.hist_words + l];
                    }
                }
                if (v) atomicAdd(sp.d_hist + i, (unsigned long long)v);
            }
        // block totals in integers: n, sum d, sum d^2 (128 bits) -- the table path's lanes plus the
        // exactly scored tiles of every warp.  Exact sums: the record does not depend on how the
        // tiles fell to the warps and blocks.
        unsigned long long bn = (unsigned long long)cnt, bsd = (unsigned long long)sum_d, blo = sum_d2, bhi = sum_d2_hi;
        if (lane == 0) {
            const unsigned long long* e = exact_acc + 4 * warp;
            bn += e[0]; bsd += e[1];
            blo += e[2]; bhi += e[3] + (blo < e[2] ? 1ull : 0ull);
        }
        block_reduce_moments(bn, bsd, blo, bhi, red_scratch);
        if (threadIdx.x == 0) {
            unsigned long long* p = work.partials + 4ll * (work.partial_base + blockIdx.x);
            p[0] = bn; p[1] = bsd; p[2] = blo; p[3] = bhi;
            __threadfence();
            const unsigned long long done =
                atomicAdd(reinterpret_cast<unsigned long long*>(work.counters + CGBK_CTR_INTERNAL_DONE), 1ull);
            is_last_block = sp.finalize && done + 1 == (unsigned long long)(work.partial_base + gridDim.x);
        }
        __syncthreads();
        if (is_last_block) {
            // Every block of this and of the earlier tile ranges has published its partial and
            // flushed its histogram.  All global loads of the finalisation are issued together
            // (partials, the record being accumulated into, the histogram for the host copy):
            // one L2 round trip instead of three dependent ones.
            __threadfence();
            const int n_part = work.partial_base + (int)gridDim.x;
            unsigned long long tn = 0ull, tsd = 0ull, tlo = 0ull, thi = 0ull;
            for (int i = threadIdx.x; i < n_part; i += blockDim.x) {
                const unsigned long long* p = work.partials + 4ll * i;
                const unsigned long long l = __ldcg(p + 2);
                tn += __ldcg(p); tsd += __ldcg(p + 1);
                tlo += l; thi += __ldcg(p + 3) + (tlo < l ? 1ull : 0ull);
            }
            double prev = 0.0;                                   // n, sum, sumsq accumulated so far
            if (threadIdx.x < 3) prev = __ldcg(sp.d_stats + threadIdx.x);
            unsigned long long hcopy = 0ull;                     // n_bins <= blockDim.x in the common case
            const bool copy_hist = HIST && sp.host_stats && sp.host_hist && sp.d_hist;
            if (copy_hist && (int)threadIdx.x < sp.n_bins) hcopy = __ldcg(sp.d_hist + threadIdx.x);
            block_reduce_moments(tn, tsd, tlo, thi, red_scratch);
            double s1 = 0.0, s2 = 0.0, cn = 0.0;
            if (threadIdx.x == 0) moments_from_integers(tn, tsd, tlo, thi, sp.c, sp.q, cn, s1, s2);
            // thread 0 holds the totals; threads 0..2 hold the previous n / sum / sumsq
            const double prev_n = __shfl_sync(0xffffffffu, prev, 0), prev_s1 = __shfl_sync(0xffffffffu, prev, 1),
                         prev_s2 = __shfl_sync(0xffffffffu, prev, 2);
            if (threadIdx.x == 0) {
                cn += prev_n; s1 += prev_s1; s2 += prev_s2;
                const double mean = cn > 0 ? s1 / cn : nan("");
                double var = cn > 0 ? s2 / cn - mean * mean : nan("");
                if (var < 0) var = 0;
                const double sd = sqrt(var);
                double* st = sp.d_stats;
                st[CGBK_BG_N] = cn; st[CGBK_BG_SUM] = s1; st[CGBK_BG_SUMSQ] = s2;
                st[CGBK_BG_MEAN] = mean; st[CGBK_BG_STD] = sd;
                if (sp.host_stats) {
                    double* hs = sp.host_stats;
                    hs[CGBK_BG_N] = cn; hs[CGBK_BG_SUM] = s1; hs[CGBK_BG_SUMSQ] = s2;
                    hs[CGBK_BG_MEAN] = mean; hs[CGBK_BG_STD] = sd;
                    for (int k = CGBK_BG_STD + 1; k < CGBK_BG_COUNT; ++k) hs[k] = 0.0;
                }
            }
            if (copy_hist) {
                if ((int)threadIdx.x < sp.n_bins) sp.host_hist[threadIdx.x] = hcopy;
                for (int i = threadIdx.x + blockDim.x; i < sp.n_bins; i += blockDim.x)
                    sp.host_hist[i] = __ldcg(sp.d_hist + i);
            }
        }
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of the function: one flag
// per (template instance, device), set under a lock (a host may drive several GPUs from one
// process, from several threads).
#define CGBK_MAX_DEVICES 64

template <int G, int MODE, int VAR>
static cudaError_t launch_one(const uint32_t* gtab, GenomeView g, SeqTable seqs, FastWork w, int m, int L,
                              int grid, StatsParams sp, int device, cudaStream_t st) {
    using Cfg = FastCfg<G, MODE>;
    const size_t fixed = TableLayout<Cfg::NW>::BYTES + Cfg::STASH_BYTES + Cfg::CAND_BYTES;
    // histogram: (n_bins + 1) rows of 1 word (<= 4096 bins) or of 32 words (<= 256 bins, conflict free)
    constexpr bool HIST = MODE == 1 || MODE == 3;
    const size_t smem_max = fixed + (HIST ? (size_t)(256 + 2) * 128 : 0);
    const size_t smem = fixed + (HIST ? ((size_t)sp.n_bins + 2) * 4 * sp.hist_words : 0);
    static std::mutex mu;
    static bool configured[CGBK_MAX_DEVICES] = {};   // per template instance
    auto kern = fast_scan_kernel<G, MODE, VAR>;
    if (device < 0 || device >= CGBK_MAX_DEVICES) return cudaErrorInvalidDevice;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[device]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
            if (e != cudaSuccess) return e;
            configured[device] = true;
        }
    }
    kern<<<grid, Cfg::THREADS, smem, st>>>(gtab, g, seqs, w, m, L / 32, sp);
    return cudaGetLastError();
}
