// Host-side packer: ASCII bases -> the three planes of the device layout (cgbk_device.cuh), on the
// CPU.  It replaces "str(record.seq)" (cgb/chromid.py:57-60) as the sequence store one step
// earlier than pack_kernel does: a replicon that is kept (or stored on disk) in packed form crosses
// PCIe as 0.25 - 0.5 byte per base instead of 1, and the pipeline runs no pack kernel
// (cgbk_regulation_pipeline_packed).  Bit-identical to pack_kernel by construction and by test.
//
//   bases2  2 bits/base (A0 C1 G2 T3), base p of a word at bits 2(p % 16); non-ACGT -> 0
//   amb     1 bit/base, byte outside ACGTacgt;   lower  1 bit/base, lower-case acgt
//
// 32 bases per iteration.  AVX2 when the CPU has it (checked at run time; the generic 64-bit
// SWAR loop otherwise); large inputs are split over host threads.
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "cgbk_internal.cuh"

namespace {

// spread the 32 bits of x to the even bit positions of a 64-bit word
inline uint64_t spread32(uint64_t x) {
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// high bit of every byte that equals c (exact, no carries between bytes)
inline uint64_t eq_bytes(uint64_t u, uint8_t c) {
    const uint64_t z = u ^ (0x0101010101010101ull * c);
    return ~(((z & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | z) & 0x8080808080808080ull;
}

// the high bits of the 8 bytes gathered into bits 0..7
inline uint32_t gather_high(uint64_t v) { return (uint32_t)(((v >> 7) * 0x0102040810204080ull) >> 56); }

struct Block32 {
    uint64_t codes;      // 32 x 2 bits
    uint32_t amb, lower;
};

inline Block32 pack32_generic(const uint8_t* p) {
    Block32 r = {0, 0, 0};
    for (int k = 0; k < 4; ++k) {
        uint64_t w;
        memcpy(&w, p + 8 * k, 8);
        const uint64_t u = w & 0xdfdfdfdfdfdfdfdfull;
        const uint64_t valid = eq_bytes(u, 'A') | eq_bytes(u, 'C') | eq_bytes(u, 'G') | eq_bytes(u, 'T');
        const uint64_t t = ((w >> 1) ^ (w >> 2)) & 0x0303030303030303ull & ((valid >> 7) * 3);
        uint64_t x = (t | (t >> 6)) & 0x000f000f000f000full;
        x = (x | (x >> 12)) & 0x000000ff000000ffull;
        x = (x | (x >> 24)) & 0xffffull;
        r.codes |= x << (16 * k);
        r.amb |= (gather_high(~valid & 0x8080808080808080ull)) << (8 * k);
        r.lower |= (gather_high((w << 2) & valid)) << (8 * k);
    }
    return r;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline Block32 pack32_avx2(const uint8_t* p) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const __m256i u = _mm256_and_si256(v, _mm256_set1_epi8((char)0xdf));
    const __m256i valid = _mm256_or_si256(
        _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C'))),
        _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'))));
    // code = ((c >> 1) ^ (c >> 2)) & 3, zero where not valid; only bits 1..2 of each byte are used,
    // so 16-bit shifts (which leak bits across bytes at the top) are fine
    const __m256i t = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), valid);
    const uint32_t bit0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(t, 7));
    const uint32_t bit1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(t, 6));
    Block32 r;
    r.codes = spread32(bit0) | (spread32(bit1) << 1);
    r.amb = ~(uint32_t)_mm256_movemask_epi8(valid);
    r.lower = (uint32_t)_mm256_movemask_epi8(_mm256_and_si256(_mm256_slli_epi16(v, 2), valid));
    return r;
}

__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t* a, int64_t blocks, uint64_t* b2, uint32_t* amb,
                                                     uint32_t* lower, int64_t* n_amb, int64_t* n_lower) {
    int64_t ca = 0, cl = 0;
    for (int64_t i = 0; i < blocks; ++i) {
        const Block32 r = pack32_avx2(a + 32 * i);
        b2[i] = r.codes; amb[i] = r.amb; lower[i] = r.lower;
        ca += __builtin_popcount(r.amb); cl += __builtin_popcount(r.lower);
    }
    *n_amb = ca; *n_lower = cl;
}
#endif

void pack_range_generic(const uint8_t* a, int64_t blocks, uint64_t* b2, uint32_t* amb, uint32_t* lower,
                        int64_t* n_amb, int64_t* n_lower) {
    int64_t ca = 0, cl = 0;
    for (int64_t i = 0; i < blocks; ++i) {
        const Block32 r = pack32_generic(a + 32 * i);
        b2[i] = r.codes; amb[i] = r.amb; lower[i] = r.lower;
        ca += __builtin_popcount(r.amb); cl += __builtin_popcount(r.lower);
    }
    *n_amb = ca; *n_lower = cl;
}

void pack_range(const uint8_t* a, int64_t blocks, uint64_t* b2, uint32_t* amb, uint32_t* lower, int64_t* n_amb,
                int64_t* n_lower) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) {
        pack_range_avx2(a, blocks, b2, amb, lower, n_amb, n_lower);
        return;
    }
#endif
    pack_range_generic(a, blocks, b2, amb, lower, n_amb, n_lower);
}

}  // namespace

extern "C" int64_t cgbk_packed_capacity(int64_t n_bases) {
    int64_t cap = (n_bases + CGBK_SEQ_ALIGN - 1) / CGBK_SEQ_ALIGN * CGBK_SEQ_ALIGN;
    return cap > 0 ? cap : CGBK_SEQ_ALIGN;
}

extern "C" int cgbk_pack_host(const uint8_t* ascii, int64_t n_bases, uint32_t* bases2, uint32_t* amb, uint32_t* lower,
                              int64_t cap_bases, int n_threads, int64_t* n_amb, int64_t* n_lower) {
    if (n_bases < 0 || (n_bases > 0 && !ascii) || !bases2 || !amb || !lower)
        return cgbk_set_error(CGBK_ERR_INVALID, "NULL or empty argument");
    if (cap_bases < cgbk_packed_capacity(n_bases) || cap_bases % CGBK_SEQ_ALIGN)
        return cgbk_set_error(CGBK_ERR_INVALID, "cap_bases must be a multiple of %d that covers the sequence",
                              CGBK_SEQ_ALIGN);
    const int64_t full = n_bases / 32;                       // whole 32-base blocks
    int64_t tot_a = 0, tot_l = 0;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 1 || full < (1 << 17)) nt = 1;                  // below 4 Mbp a thread start costs more than it saves
    uint64_t* b2 = reinterpret_cast<uint64_t*>(bases2);
    if (nt == 1) {
        pack_range(ascii, full, b2, amb, lower, &tot_a, &tot_l);
    } else {
        std::vector<std::thread> th;
        std::vector<int64_t> ca(nt, 0), cl(nt, 0);
        for (int t = 0; t < nt; ++t) {
            const int64_t lo = full * t / nt, hi = full * (t + 1) / nt;
            th.emplace_back([=, &ca, &cl]() { pack_range(ascii + 32 * lo, hi - lo, b2 + lo, amb + lo, lower + lo, &ca[t], &cl[t]); });
        }
        for (auto& x : th) x.join();
        for (int t = 0; t < nt; ++t) { tot_a += ca[t]; tot_l += cl[t]; }
    }
    // the ragged tail: pad with 'A' (packs to zero bits in every plane), then zero words up to cap
    int64_t blk = full;
    if (n_bases % 32) {
        uint8_t tail[32];
        memset(tail, 'A', 32);
        memcpy(tail, ascii + 32 * full, (size_t)(n_bases % 32));
        int64_t a = 0, l = 0;
        pack_range_generic(tail, 1, b2 + full, amb + full, lower + full, &a, &l);
        tot_a += a; tot_l += l;
        ++blk;
    }
    const int64_t blocks_cap = cap_bases / 32;
    if (blk < blocks_cap) {
        memset(b2 + blk, 0, (size_t)(blocks_cap - blk) * 8);
        memset(amb + blk, 0, (size_t)(blocks_cap - blk) * 4);
        memset(lower + blk, 0, (size_t)(blocks_cap - blk) * 4);
    }
    if (n_amb) *n_amb = tot_a;
    if (n_lower) *n_lower = tot_l;
    return CGBK_OK;
}
