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
    uint64_t lo = (uint64_t)w0 | ((uint64_t)w1 << 32);
    if (sh == 0) return lo;
    const uint32_t w2 = ldw(g.bases2, wi + 2, nw);
    return (lo >> sh) | ((uint64_t)w2 << (64 - sh));
}

// 32 consecutive mask bits starting at base offset p
__device__ __forceinline__ uint32_t load_bits32(const uint32_t* __restrict__ plane, long long cap_bases,
                                                long long p) {
    const long long nw = cap_bases >> 5;
    const long long wi = p >> 5;
    const uint32_t a0 = ldw(plane, wi, nw), a1 = ldw(plane, wi + 1, nw);
    return __funnelshift_r(a0, a1, (uint32_t)(p & 31));
}

// Exact tables staged in shared memory: fwd[m][4], rc[m][4], mean_f[m], mean_r[m]
struct ExactTables {
    double fwd[CGBK_MAXM * 4];
    double rc[CGBK_MAXM * 4];
    double mean_f[CGBK_MAXM];
    double mean_r[CGBK_MAXM];
};

__device__ __forceinline__ void stage_exact(ExactTables* s, const double* __restrict__ d_exact) {
    double* dst = reinterpret_cast<double*>(s);
    for (int i = threadIdx.x; i < CGBK_EXACT_DOUBLES; i += blockDim.x) dst[i] = d_exact[i];
}

// The reference's arithmetic for one window starting at base offset p.
//   f: Biopython _pwm.c calculate(): double accumulate, j ascending (pssm_model.py:114)
//   r: CGBK_RC_MATRIX_ORDER: rev_comp_pssm.calculate(seq) (pssm_model.py:115)
//      CGBK_RC_REVSEQ_ORDER: pssm.calculate(reverse_complement(seq)) read back through
//      reversed() (genome.py:440-445): same addends, accumulated from the window's far end
//   windows holding a byte outside ACGTacgt: PSSMModel._calculate (pssm_model.py:88-101),
//   where lower-case letters are KeyErrors too and add the column mean.
// Returns true when the window went through the ambiguity repair.
__device__ __forceinline__ bool exact_window(const GenomeView& g, const ExactTables* t, int m,
                                             long long p, bool revseq, double& f, double& r) {
    const uint64_t x = load_bases32(g, p);
    const uint32_t maskm = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    const uint32_t amb = load_bits32(g.amb, g.cap_bases, p) & maskm;
    double fs = 0.0, rs = 0.0;
    if (amb == 0) {
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const int b = (int)((x >> (2 * j)) & 3);
            fs += t->fwd[j * 4 + b];
        }
        if (!revseq) {
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
                const int b = (int)((x >> (2 * j)) & 3);
                rs += t->rc[j * 4 + b];
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
                const int b = (int)((x >> (2 * (m - 1 - j))) & 3);
                rs += t->fwd[j * 4 + (3 - b)];
            }
        }
        f = fs; r = rs;
        return false;
    }
    const uint32_t unk = amb | (load_bits32(g.lower, g.cap_bases, p) & maskm);
    for (int j = 0; j < m; ++j) {
        const int b = (int)((x >> (2 * j)) & 3);
        fs += ((unk >> j) & 1u) ? t->mean_f[j] : t->fwd[j * 4 + b];
    }
    if (!revseq) {
        for (int j = 0; j < m; ++j) {
            const int b = (int)((x >> (2 * j)) & 3);
            rs += ((unk >> j) & 1u) ? t->mean_r[j] : t->rc[j * 4 + b];
        }
    } else {
        for (int j = 0; j < m; ++j) {
            const int q = m - 1 - j;
            const int b = (int)((x >> (2 * q)) & 3);
            rs += ((unk >> q) & 1u) ? t->mean_f[j] : t->fwd[j * 4 + (3 - b)];
        }
    }
    f = fs; r = rs;
    return true;
}

// pssm_model.py:129-131 with the pinned numpy 1.x: pow(2, s) in double, math.log(x, 2)
__device__ __forceinline__ double softmax2_f64(double s, double r) {
    return log(exp2(s) + exp2(r)) / 0.6931471805599453094;
}

// sequence of tile -> (sequence index, local tile)
__device__ __forceinline__ int find_seq(const SeqTable& s, long long tile) {
    int lo = 0, hi = s.n_seq;   // tile_prefix[lo] <= tile < tile_prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(s.tile_prefix + mid) <= tile) lo = mid; else hi = mid;
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
}
