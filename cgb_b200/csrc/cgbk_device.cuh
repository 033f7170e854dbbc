// Device helpers shared by the exact and the tiled kernels.
#pragma once

#include "cgbk_internal.cuh"

__device__ __forceinline__ uint32_t ldw(const uint32_t* __restrict__ p, long long i, long long n) {
    return (i >= 0 && i < n) ? __ldg(p + i) : 0u;
}

// 32 consecutive bases starting at base offset p (2 bits each, base p in bits 0..1)
__device__ __forceinline__ uint64_t load_bases32(const GenomeView& g, long long p) {
    const long long nw = g.cap_bases >> 4;
    const long long wi = p >> 4;
    const int sh = (int)(p & 15) * 2;
    const uint32_t w0 = ldw(g.bases2, wi, nw), w1 = ldw(g.bases2, wi + 1, nw);
    const uint32_t w2 = ldw(g.bases2, wi + 2, nw);
    // funnel shifts: sh in [0, 30]
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

// 32 consecutive mask bits starting at base offset p
__device__ __forceinline__ uint32_t load_bits32(const uint32_t* __restrict__ plane, long long cap_bases,
                                                long long p) {
    const long long nw = cap_bases >> 5;
    const long long wi = p >> 5;
    const uint32_t a0 = ldw(plane, wi, nw), a1 = ldw(plane, wi + 1, nw);
    return __funnelshift_r(a0, a1, (uint32_t)(p & 31));
}

// Exact tables staged in shared memory, the two strands interleaved so that one 16-byte
// load serves both: fr[j][b] = (pssm[b][j], rev_comp_pssm[b][j]); mean[j] likewise.
struct ExactTables {
    double2 fr[CGBK_MAXM * 4];
    double2 mean[CGBK_MAXM];
};

// d_exact layout (global): fwd[128] rc[128] mean_f[32] mean_r[32]
__device__ __forceinline__ void stage_exact(ExactTables* s, const double* __restrict__ d_exact) {
    for (int i = threadIdx.x; i < CGBK_MAXM * 4; i += blockDim.x)
        s->fr[i] = make_double2(d_exact[i], d_exact[CGBK_MAXM * 4 + i]);
    for (int i = threadIdx.x; i < CGBK_MAXM; i += blockDim.x)
        s->mean[i] = make_double2(d_exact[CGBK_MAXM * 8 + i], d_exact[CGBK_MAXM * 9 + i]);
}

// The reference's arithmetic for one window starting at base offset p.
//   f: Biopython _pwm.c calculate(): double accumulate, j ascending (pssm_model.py:114)
//   r: CGBK_RC_MATRIX_ORDER: rev_comp_pssm.calculate(seq) (pssm_model.py:115)
//      CGBK_RC_REVSEQ_ORDER: pssm.calculate(reverse_complement(seq)) read back through
//      reversed() (genome.py:440-445): same addends, accumulated from the window's far end
//   windows holding a byte outside ACGTacgt: PSSMModel._calculate (pssm_model.py:88-101),
//   where lower-case letters are KeyErrors too and add the column mean.
// Returns true when the window went through the ambiguity repair.
// x = the window's bases (2 bits each from bit 0), amb = its ambiguity bits (already masked to m).
__device__ __forceinline__ bool exact_window_bits(const GenomeView& g, const ExactTables* t, int m, long long p,
                                                  uint64_t x, uint32_t amb, bool revseq, double& f, double& r) {
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    const uint32_t maskm = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    double fs = 0.0, rs = 0.0;
    if (amb == 0) {
        // Four positions per iteration, the window's 2-bit word shifted along by one byte: the
        // look-up offsets come from immediate shifts of that byte ((code << 4) & 0x30 into the
        // 64-byte row of a position).  The last block may run past the motif: rows j >= m hold
        // +0.0 (the model block is zero filled), and s + 0.0 == s exactly, so the sums are the
        // reference's left-to-right sums over j < m, bit for bit.
        const char* tb = reinterpret_cast<const char*>(t->fr);
        const int nblk = (m + 3) >> 2;
        if (!revseq) {
            uint32_t xw = xl;
            for (int b = 0; b < nblk; ++b) {
                if (b == 4) xw = xh;
                const char* q = tb + b * 256;
                const double2 v0 = *reinterpret_cast<const double2*>(q + ((xw << 4) & 0x30));
                const double2 v1 = *reinterpret_cast<const double2*>(q + 64 + ((xw << 2) & 0x30));
                const double2 v2 = *reinterpret_cast<const double2*>(q + 128 + (xw & 0x30));
                const double2 v3 = *reinterpret_cast<const double2*>(q + 192 + ((xw >> 2) & 0x30));
                fs += v0.x; rs += v0.y;
                fs += v1.x; rs += v1.y;
                fs += v2.x; rs += v2.y;
                fs += v3.x; rs += v3.y;
                xw >>= 8;
            }
        } else {
            // y_j = 3 - x_{m-1-j}: the reverse complement as a base string of its own (bit reversal
            // mirrors the positions and swaps the two bits of every base; swap them back, complement,
            // right-align), so that both strands walk j upwards with one shift-and-mask per look-up.
            // Same addends in the same order as fs += fr[j][x_j].x; rs += fr[j][3 - x_{m-1-j}].x.
            uint64_t y = __brevll(x);
            y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
            y = (~y) >> (64 - 2 * m);
            uint32_t xw = xl, yw = (uint32_t)y;
            for (int b = 0; b < nblk; ++b) {
                if (b == 4) { xw = xh; yw = (uint32_t)(y >> 32); }
                const char* q = tb + b * 256;
                fs += *reinterpret_cast<const double*>(q + ((xw << 4) & 0x30));
                rs += *reinterpret_cast<const double*>(q + ((yw << 4) & 0x30));
                fs += *reinterpret_cast<const double*>(q + 64 + ((xw << 2) & 0x30));
                rs += *reinterpret_cast<const double*>(q + 64 + ((yw << 2) & 0x30));
                fs += *reinterpret_cast<const double*>(q + 128 + (xw & 0x30));
                rs += *reinterpret_cast<const double*>(q + 128 + (yw & 0x30));
                fs += *reinterpret_cast<const double*>(q + 192 + ((xw >> 2) & 0x30));
                rs += *reinterpret_cast<const double*>(q + 192 + ((yw >> 2) & 0x30));
                xw >>= 8; yw >>= 8;
            }
        }
        f = fs; r = rs;
        return false;
    }
    const uint32_t unk = amb | (load_bits32(g.lower, g.cap_bases, p) & maskm);
    for (int j = 0; j < m; ++j) {
        const int b = (int)((x >> (2 * j)) & 3);
        const bool u = (unk >> j) & 1u;
        fs += u ? t->mean[j].x : t->fr[j * 4 + b].x;
        if (!revseq) {
            rs += u ? t->mean[j].y : t->fr[j * 4 + b].y;
        } else {
            const int q = m - 1 - j;
            const int bq = (int)((x >> (2 * q)) & 3);
            rs += ((unk >> q) & 1u) ? t->mean[j].x : t->fr[j * 4 + (3 - bq)].x;
        }
    }
    f = fs; r = rs;
    return true;
}

__device__ __forceinline__ bool exact_window(const GenomeView& g, const ExactTables* t, int m,
                                             long long p, bool revseq, double& f, double& r) {
    const uint64_t x = load_bases32(g, p);
    const uint32_t maskm = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    const uint32_t amb = load_bits32(g.amb, g.cap_bases, p) & maskm;
    return exact_window_bits(g, t, m, p, x, amb, revseq, f, r);
}

// pssm_model.py:129-131 with the pinned numpy 1.x: pow(2, s) in double, math.log(x, 2)
__device__ __forceinline__ double softmax2_f64(double s, double r) {
    return log(exp2(s) + exp2(r)) / 0.6931471805599453094;
}

// the same soft-max for float32 strand scores with a float32 tail:
// max + log2(1 + 2^-(max - min)); |error| ~ 1e-7 (used where 1e-7 is ample: posteriors)
__device__ __forceinline__ double softmax2_f32tail(float ff, float rf) {
    const float mx = fmaxf(ff, rf), mn = fminf(ff, rf);
    return (double)mx + (double)log2f(1.0f + exp2f(mn - mx));
}

// Warp-convergent append of up to two hits per lane (+ strand, - strand) to the compact hit
// list: ONE atomic on the batch's append position and one on the model's hit counter per warp.
__device__ __forceinline__ void emit_hits8_warp(unsigned long long* hits8, long long cap, long long* cursor,
                                                long long* counters, long long gpos, bool hf, float ff,
                                                bool hr, float rf) {
    const unsigned int mf = __ballot_sync(0xffffffffu, hf), mr = __ballot_sync(0xffffffffu, hr);
    const int nf = __popc(mf), tot = nf + __popc(mr);
    if (tot == 0) return;                                   // warp-uniform
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(reinterpret_cast<unsigned long long*>(cursor), (unsigned long long)tot);
        atomicAdd(reinterpret_cast<unsigned long long*>(counters + CGBK_CTR_HITS), (unsigned long long)tot);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned int lt = (1u << lane) - 1u;
    if (hf) {
        const long long k = (long long)base + __popc(mf & lt);
        if (k < cap) hits8[k] = cgbk_pack_hit8(gpos, 0, __float_as_uint(ff));
    }
    if (hr) {
        const long long k = (long long)base + nf + __popc(mr & lt);
        if (k < cap) hits8[k] = cgbk_pack_hit8(gpos, 1, __float_as_uint(rf));
    }
}

// Sequence table accessors: tables of up to CGBK_INLINE_SEQS sequences travel by value in the
// kernel parameters (constant bank), larger ones live in device memory.
__device__ __forceinline__ long long seq_off_of(const SeqTable& s, int k) {
    return s.n_seq <= CGBK_INLINE_SEQS ? s.off_v[k] : __ldg(s.off + k);
}
__device__ __forceinline__ long long seq_len_of(const SeqTable& s, int k) {
    return s.n_seq <= CGBK_INLINE_SEQS ? s.len_v[k] : __ldg(s.len + k);
}
__device__ __forceinline__ long long seq_prefix_of(const SeqTable& s, int k) {
    return s.n_seq <= CGBK_INLINE_SEQS ? s.prefix_v[k] : __ldg(s.tile_prefix + k);
}

// sequence of a tile: tile_prefix[k] <= tile < tile_prefix[k + 1]
__device__ __forceinline__ int find_seq(const SeqTable& s, long long tile) {
    if (s.n_seq <= CGBK_INLINE_SEQS) {
        int k = 0;
#pragma unroll
        for (int i = 1; i < CGBK_INLINE_SEQS; ++i)
            if (i < s.n_seq && s.prefix_v[i] <= tile) k = i;
        return k;
    }
    int lo = 0, hi = s.n_seq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(s.tile_prefix + mid) <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// sequence holding packed offset gpos: off[k] <= gpos < off[k + 1]
__device__ __forceinline__ int find_seq_by_offset(const SeqTable& s, long long gpos) {
    if (s.n_seq <= CGBK_INLINE_SEQS) {
        int k = 0;
#pragma unroll
        for (int i = 1; i < CGBK_INLINE_SEQS; ++i)
            if (i < s.n_seq && s.off_v[i] <= gpos) k = i;
        return k;
    }
    int lo = 0, hi = s.n_seq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(s.off + mid) <= gpos) lo = mid; else hi = mid;
    }
    return lo;
}

// deterministic block reduction of (a, b, c) doubles; result valid in thread 0
__device__ __forceinline__ void block_reduce3(double& a, double& b, double& c, double* scratch /*3*32*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { scratch[warp] = a; scratch[32 + warp] = b; scratch[64 + warp] = c; }
    __syncthreads();
    if (warp == 0) {
        a = lane < nw ? scratch[lane] : 0.0;
        b = lane < nw ? scratch[32 + lane] : 0.0;
        c = lane < nw ? scratch[64 + lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
    }
    __syncthreads();
}

// windows [w_lo, w_hi) owned by a tile and the packed offset of its sequence
__device__ __forceinline__ void tile_range(const SeqTable& seqs, unsigned int tile, int L, int m,
                                           long long& gbase, long long& w_lo, long long& w_hi) {
    const int s = find_seq(seqs, tile);
    const long long local = (long long)tile - seq_prefix_of(seqs, s);
    const long long stride = 32ll * L - 32;
    const long long nwin = seq_len_of(seqs, s) - m + 1;
    gbase = seq_off_of(seqs, s);
    w_lo = local * stride;
    w_hi = w_lo + stride < nwin ? w_lo + stride : nwin;
}

// block-wide: per-block partial moments (sum, sumsq, n) -> stats record (n, sum, sumsq, mean, std)
__device__ __forceinline__ void finalize_stats(const double* partials, int n_partials,
                                               double* stats, int accumulate, double* scratch) {
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) {
        s1 += partials[3 * i]; s2 += partials[3 * i + 1]; cnt += partials[3 * i + 2];
    }
    block_reduce3(s1, s2, cnt, scratch);
    if (threadIdx.x == 0) {
        if (accumulate) {
            cnt += stats[CGBK_BG_N]; s1 += stats[CGBK_BG_SUM]; s2 += stats[CGBK_BG_SUMSQ];
        }
        stats[CGBK_BG_N] = cnt; stats[CGBK_BG_SUM] = s1; stats[CGBK_BG_SUMSQ] = s2;
        const double mean = cnt > 0 ? s1 / cnt : nan("");
        double var = cnt > 0 ? s2 / cnt - mean * mean : nan("");
        if (var < 0) var = 0;
        stats[CGBK_BG_MEAN] = mean;
        stats[CGBK_BG_STD] = sqrt(var);
    }
}
