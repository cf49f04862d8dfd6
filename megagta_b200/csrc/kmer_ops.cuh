// kmer_ops.cuh -- multi-word 2-bit sequence primitives on registers (sm_100a).
//
// Replaces the reference's rolling GenericKmer (megahit_kmer.h:42-174: init / ShiftAppend /
// ShiftPreappend / ReverseComplement / cmp) and CopySubstring[RC] (packed_reads.h:44-176).  The
// reference rolls one k-mer per read sequentially; here every base position is an independent
// thread that cuts its window out of shared-memory-staged read words with funnel shifts, so all
// arrays are register-resident (every index below is a compile-time constant after unrolling).
//
// Layout everywhere: chars MSB-first, 16 chars per u32 (definitions.h:40-42), zero padded.
//
// The functions are __host__ __device__ so tests/cpu/logic_test.cpp can run the very same item
// logic on the CPU against the oracle (no GPU in the build container); the product only ever
// calls them from kernels.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MGTA_HD __host__ __device__ __forceinline__
#else
#define MGTA_HD inline
#endif

namespace mgta {

MGTA_HD uint32_t brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// high 32 bits of (hi:lo) << s, 0 <= s < 32
MGTA_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return s ? ((hi << s) | (lo >> (32 - s))) : hi;
#endif
}

// reverse-complement of one word of 16 chars (bit_operation.h:40-46 does it with 4 swaps + ~)
MGTA_HD uint32_t rc_word(uint32_t x) {
    uint32_t t = brev32(x);                                     // also reverses the 2 bits inside each char
    t = ((t >> 1) & 0x55555555u) | ((t & 0x55555555u) << 1);    // restore bit order inside chars
    return ~t;
}

// mask keeping the first `kept` (<=0 .. >=16) chars of a word
MGTA_HD uint32_t head_mask(int kept) {
    return kept <= 0 ? 0u : (kept >= 16 ? 0xFFFFFFFFu : (0xFFFFFFFFu << (32 - 2 * kept)));
}

// X[0..W) = n chars starting at char offset q of `words` (reads words[q/16 .. q/16 + W])
template <int W>
MGTA_HD void load_chars(const uint32_t *words, uint32_t q, int n, uint32_t (&X)[W]) {
    const uint32_t idx = q >> 4, bs = (q & 15) * 2;
    uint32_t cur = words[idx];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        uint32_t nxt = words[idx + i + 1];
        X[i] = funnel_l(nxt, cur, bs) & head_mask(n - 16 * i);
        cur = nxt;
    }
}

MGTA_HD int char_at(const uint32_t *words, uint32_t q) {
    return (words[q >> 4] >> ((15 - (q & 15)) * 2)) & 3;
}

// Y = reverse complement of the n chars in X (both MSB-aligned, zero padded); needs 16W - n < 32
template <int W>
MGTA_HD void revcomp(const uint32_t (&X)[W], int n, uint32_t (&Y)[W]) {
    uint32_t T[W + 2];
#pragma unroll
    for (int i = 0; i < W; ++i) T[i] = rc_word(X[W - 1 - i]);
    T[W] = 0;
    T[W + 1] = 0;
    const int pad = 16 * W - n;  // leading chars of T that came from X's zero padding
    const bool ws = pad >= 16;
    const uint32_t bs = (pad & 15) * 2;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        uint32_t hi = ws ? T[i + 1] : T[i];
        uint32_t lo = ws ? T[i + 2] : T[i + 1];
        Y[i] = funnel_l(lo, hi, bs);
    }
}

// lexicographic compare of zero-padded arrays (megahit_kmer.h:115-128)
template <int W>
MGTA_HD int cmp_words(const uint32_t (&A)[W], const uint32_t (&B)[W]) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
        if (A[i] != B[i]) return A[i] < B[i] ? -1 : 1;
    }
    return 0;
}

// Y = chars [c, c+n) of X, MSB-aligned, zero padded (0 <= c < 16)
template <int W>
MGTA_HD void sub_chars(const uint32_t (&X)[W], int c, int n, uint32_t (&Y)[W]) {
    const uint32_t bs = 2 * c;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        uint32_t lo = (i + 1 < W) ? X[(i + 1 < W) ? i + 1 : i] : 0u;
        Y[i] = funnel_l(lo, X[i], bs) & head_mask(n - 16 * i);
    }
}

}  // namespace mgta
